// Network container + layer factory with the reference's JSON schema, validation rules and pass order
// (NeuralNetwork.cpp:37-190, LayerFactory.cu:43-88).
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "Layers.hpp"

struct LayerFactory {
    static layers::Layer *createLayer(bl_ctx *ctx, const std::string &layerType, const helpers::JsonValue &layerChild,
                                      const helpers::JsonValue *weightsSection, int parallelSequences, int maxSeqLength,
                                      layers::Layer *precedingLayer = nullptr);
};

class NeuralNetwork {
public:
    NeuralNetwork(bl_ctx *ctx, helpers::JsonDocument &jsonDoc, int parallelSequences, int maxSeqLength,
                  int inputSizeOverride = -1, int outputSizeOverride = -1);
    ~NeuralNetwork();

    const std::vector<std::shared_ptr<layers::Layer>> &layers() const { return m_layers; }
    layers::InputLayer &inputLayer();
    layers::TrainableLayer &outputLayer();
    layers::PostOutputLayer &postOutputLayer();

    void loadSequences(const data_sets::DataSetFraction &fraction);
    void computeForwardPass();
    void computeBackwardPass();
    real_t calculateError();

    void exportLayers(helpers::JsonDocument &jsonDoc) const;
    void exportWeights(helpers::JsonDocument &jsonDoc) const;
    std::vector<std::vector<std::vector<real_t>>> getOutputs();

    // data parallelism (no reference counterpart): when a communicator is attached, every trainable
    // layer's weightUpdates() is handed to bl_allreduce_sum_f32 as soon as that layer's backward is enqueued;
    // joinGradients() completes the reductions before the update (schedules: see csrc/comm.cu).
    void setCommunicator(bl_comm *comm) { m_comm = comm; }
    bl_comm *communicator() const { return m_comm; }
    // a rank whose shard of the fraction is empty still has to take part in the reduction, with zero gradients
    void contributeZeroGradients();
    void joinGradients();
    bl_ctx *ctx() const { return m_ctx; }

private:
    const layers::Layer *lastReducedLayer() const;
    void reduceGradient(layers::Layer *layer, const layers::Layer *last);
    bl_ctx *m_ctx;
    bl_comm *m_comm;
    std::vector<std::shared_ptr<layers::Layer>> m_layers;
};
