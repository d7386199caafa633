#include "NeuralNetwork.hpp"
#include <stdexcept>

using device::check;

layers::Layer *LayerFactory::createLayer(bl_ctx *ctx, const std::string &layerType, const helpers::JsonValue &layerChild,
                                         const helpers::JsonValue *weightsSection, int parallelSequences, int maxSeqLength,
                                         layers::Layer *precedingLayer)
{
    using namespace layers;
    if (layerType == "input")
        return new InputLayer(ctx, layerChild, parallelSequences, maxSeqLength);
    if (!precedingLayer)
        throw std::runtime_error("Layer type '" + layerType + "' needs a preceding layer");
    if (layerType == "feedforward_tanh")
        return new FeedForwardLayer(BL_ACT_TANH, layerChild, weightsSection, *precedingLayer);
    if (layerType == "feedforward_logistic")
        return new FeedForwardLayer(BL_ACT_LOGISTIC, layerChild, weightsSection, *precedingLayer);
    if (layerType == "feedforward_identity")
        return new FeedForwardLayer(BL_ACT_IDENTITY, layerChild, weightsSection, *precedingLayer);
    if (layerType == "softmax")
        return new SoftmaxLayer(layerChild, weightsSection, *precedingLayer);
    if (layerType == "lstm")
        return new LstmLayer(layerChild, weightsSection, *precedingLayer, false);
    if (layerType == "blstm")
        return new LstmLayer(layerChild, weightsSection, *precedingLayer, true);
    if (layerType == "sse")
        return new SsePostOutputLayer(layerChild, *precedingLayer);
    if (layerType == "ce")
        return new CePostOutputLayer(layerChild, *precedingLayer);
    if (layerType == "multiclass_classification")
        return new MulticlassClassificationLayer(layerChild, *precedingLayer);
    if (layerType == "weightedsse")
        return new WeightedSsePostOutputLayer(layerChild, *precedingLayer);
    if (layerType == "rmse")
        return new RmsePostOutputLayer(layerChild, *precedingLayer);
    if (layerType == "wf")          // "sse_mask" never reaches the reference's constructor either (LayerFactory.cu:66, 79)
        return new SseMaskPostOutputLayer(layerChild, *precedingLayer);
    if (layerType == "binary_classification")
        return new BinaryClassificationLayer(layerChild, *precedingLayer);
    throw std::runtime_error(std::string("Unknown layer type '") + layerType + "'");
}

NeuralNetwork::NeuralNetwork(bl_ctx *ctx, helpers::JsonDocument &jsonDoc, int parallelSequences, int maxSeqLength,
                             int inputSizeOverride, int /*outputSizeOverride*/)
    : m_ctx(ctx), m_comm(nullptr)
{
    try {
        if (!jsonDoc.IsObject() || !jsonDoc.HasMember("layers"))
            throw std::runtime_error("Missing section 'layers'");
        helpers::JsonValue &layersSection = jsonDoc.member("layers");
        if (!layersSection.IsArray())
            throw std::runtime_error("Section 'layers' is not an array");

        const helpers::JsonValue *weightsSection = nullptr;
        if (jsonDoc.HasMember("weights")) {
            if (!jsonDoc["weights"].IsObject())
                throw std::runtime_error("Section 'weights' is not an object");
            weightsSection = &jsonDoc["weights"];
        }

        for (size_t i = 0; i < layersSection.Size(); ++i) {
            helpers::JsonValue &layerChild = layersSection.at(i);
            if (!layerChild.IsObject())
                throw std::runtime_error("A layer section in the 'layers' array is not an object");
            if (!layerChild.HasMember("type"))
                throw std::runtime_error("Missing value 'type' in layer description");
            const std::string layerType = layerChild["type"].GetString();
            if (inputSizeOverride > 0 && layerType == "input")
                layerChild.member("size").SetInt(inputSizeOverride);          // NeuralNetwork.cpp:71-73
            try {
                layers::Layer *layer = LayerFactory::createLayer(ctx, layerType, layerChild, weightsSection, parallelSequences,
                                                                 maxSeqLength, m_layers.empty() ? nullptr : m_layers.back().get());
                m_layers.push_back(std::shared_ptr<layers::Layer>(layer));
            } catch (const std::exception &e) {
                throw std::runtime_error(std::string("Could not create layer: ") + e.what());
            }
        }

        if (m_layers.size() < 3)
            throw std::runtime_error("Not enough layers defined");
        if (!dynamic_cast<layers::InputLayer *>(m_layers.front().get()))
            throw std::runtime_error("The first layer is not an input layer");
        for (size_t i = 1; i < m_layers.size(); ++i)
            if (dynamic_cast<layers::InputLayer *>(m_layers[i].get()))
                throw std::runtime_error("Multiple input layers defined");
        if (!dynamic_cast<layers::PostOutputLayer *>(m_layers.back().get()))
            throw std::runtime_error("The last layer is not a post output layer");
        for (size_t i = 0; i + 1 < m_layers.size(); ++i)
            if (dynamic_cast<layers::PostOutputLayer *>(m_layers[i].get()))
                throw std::runtime_error("Multiple post output layers defined");
        for (size_t i = 0; i < m_layers.size(); ++i)
            for (size_t j = 0; j < m_layers.size(); ++j)
                if (i != j && m_layers[i]->name() == m_layers[j]->name())
                    throw std::runtime_error(std::string("Different layers have the same name '") + m_layers[i]->name() + "'");
    } catch (const std::exception &e) {
        throw std::runtime_error(std::string("Invalid network file: ") + e.what());
    }
}

NeuralNetwork::~NeuralNetwork()
{
    // layers hold references to their predecessors: destroy back to front
    while (!m_layers.empty()) m_layers.pop_back();
}

layers::InputLayer &NeuralNetwork::inputLayer() { return static_cast<layers::InputLayer &>(*m_layers.front()); }
layers::TrainableLayer &NeuralNetwork::outputLayer() { return static_cast<layers::TrainableLayer &>(*m_layers[m_layers.size() - 2]); }
layers::PostOutputLayer &NeuralNetwork::postOutputLayer() { return static_cast<layers::PostOutputLayer &>(*m_layers.back()); }

void NeuralNetwork::loadSequences(const data_sets::DataSetFraction &fraction)
{
    for (auto &layer : m_layers) layer->loadSequences(fraction);
    // the copies out of the fraction's pinned buffers are asynchronous: leave a ticket so that the fraction outlives them
    unsigned long long ticket = 0;
    check(m_ctx, bl_upload_mark(m_ctx, &ticket));
    fraction.noteUpload(m_ctx, ticket);
}

void NeuralNetwork::computeForwardPass()
{
    for (auto &layer : m_layers) layer->computeForwardPass();
}

// The trainable layer whose backward pass runs last (the first hidden layer): its gradient has nothing left to overlap with.  EVERY
// rank must route the same layer through the same communicator -- the ranks of an all-reduce meet inside one communicator -- so the
// backward pass and contributeZeroGradients() (a rank whose shard of a fraction is empty) share this rule.  (They did not at first:
// C5 at 4 / 8 GPUs, where truncation leaves a ragged last fraction, hung in its first 8-GPU run.)
const layers::Layer *NeuralNetwork::lastReducedLayer() const
{
    for (auto &layer : m_layers) {
        layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer.get());
        if (tl && !tl->weightUpdates().empty()) return layer.get();
    }
    return nullptr;
}

void NeuralNetwork::reduceGradient(layers::Layer *layer, const layers::Layer *last)
{
    layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer);
    if (!m_comm || !tl || tl->weightUpdates().empty()) return;
    if (layer == last) check(m_ctx, bl_allreduce_sum_f32_last(m_comm, tl->weightUpdates().data(), tl->weightUpdates().size()));
    else check(m_ctx, bl_allreduce_sum_f32(m_comm, tl->weightUpdates().data(), tl->weightUpdates().size()));
}

void NeuralNetwork::computeBackwardPass()
{
    const layers::Layer *last = m_comm ? lastReducedLayer() : nullptr;
    for (auto it = m_layers.rbegin(); it != m_layers.rend(); ++it) {
        (*it)->computeBackwardPass();
        reduceGradient(it->get(), last);          // as soon as this layer's backward pass is enqueued
    }
}

void NeuralNetwork::contributeZeroGradients()
{
    const layers::Layer *last = m_comm ? lastReducedLayer() : nullptr;
    for (auto it = m_layers.rbegin(); it != m_layers.rend(); ++it) {
        layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(it->get());
        if (!tl || tl->weightUpdates().empty()) continue;
        check(m_ctx, bl_memset(m_ctx, tl->weightUpdates().data(), 0, tl->weightUpdates().size() * sizeof(real_t)));
        reduceGradient(it->get(), last);
    }
}

void NeuralNetwork::joinGradients()
{
    if (m_comm) check(m_ctx, bl_comm_join(m_comm));
}

real_t NeuralNetwork::calculateError()
{
    return postOutputLayer().calculateError();
}

void NeuralNetwork::exportLayers(helpers::JsonDocument &jsonDoc) const
{
    if (!jsonDoc.IsObject())
        throw std::runtime_error("JSON document root must be an object");
    helpers::JsonValue layersArray = helpers::JsonValue::makeArray();
    for (size_t i = 0; i < m_layers.size(); ++i) m_layers[i]->exportLayer(layersArray);
    jsonDoc.member("layers") = layersArray;
}

void NeuralNetwork::exportWeights(helpers::JsonDocument &jsonDoc) const
{
    if (!jsonDoc.IsObject())
        throw std::runtime_error("JSON document root must be an object");
    helpers::JsonValue weightsObject = helpers::JsonValue::makeObject();
    for (const auto &layer : m_layers) {
        layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer.get());
        if (tl) tl->exportWeights(weightsObject);
    }
    jsonDoc.member("weights") = weightsObject;
}

// per-sequence outputs of the output layer, selected by pattern type (NeuralNetwork.cpp:237-262)
std::vector<std::vector<std::vector<real_t>>> NeuralNetwork::getOutputs()
{
    layers::TrainableLayer &ol = outputLayer();
    const std::vector<real_t> all = ol.outputsToHost();
    const std::vector<char> &pat = ol.hostPatTypes();
    std::vector<std::vector<std::vector<real_t>>> outputs;
    for (int patIdx = 0; patIdx < (int)pat.size(); ++patIdx) {
        switch (pat[patIdx]) {
        case PATTYPE_FIRST:
            outputs.resize(outputs.size() + 1);
            // fall through
        case PATTYPE_NORMAL:
        case PATTYPE_LAST: {
            const int psIdx = patIdx % ol.parallelSequences();
            outputs[psIdx].push_back(std::vector<real_t>(all.begin() + (size_t)patIdx * ol.size(), all.begin() + (size_t)(patIdx + 1) * ol.size()));
            break;
        }
        default:
            break;
        }
    }
    return outputs;
}
