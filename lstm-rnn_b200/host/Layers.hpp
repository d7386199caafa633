// Host-side layer classes with the reference's names, virtual interface and semantics
// (layers/Layer.hpp:83-178, TrainableLayer.hpp:84-155, LstmLayer.hpp:141-248, FeedForwardLayer.hpp, SoftmaxLayer.hpp,
// PostOutputLayer.hpp:80, Sse/Ce/MulticlassClassification layers).  Where the reference's layers call helpers::Matrix
// (cuBLAS) and Thrust functors, these call the C ABI of include/blstm_b200.h.  Device tensors keep the reference's
// pattern-major layout with rows padded to 16 bytes (ld() floats per pattern).
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "DataSetFraction.hpp"
#include "Json.hpp"
#include "Types.hpp"

namespace layers {

class Layer {
public:
    typedef device::real_vector    real_vector;
    typedef device::pattype_vector pattype_vector;

    // padRows: pad each pattern row to 16 bytes (layers written by kernels); false for the layers filled by H2D copies
    // (input layer, dense targets), which then take ONE contiguous copy per fraction instead of a row-strided one
    Layer(bl_ctx *ctx, const helpers::JsonValue &layerChild, int parallelSequences, int maxSeqLength, bool createOutputs = true,
          bool padRows = true);
    virtual ~Layer();

    const std::string &name() const { return m_name; }
    int size() const { return m_size; }
    int ld() const { return m_ld; }                       // floats per pattern row in outputs()/outputErrors()
    int parallelSequences() const { return m_parallelSequences; }
    int maxSeqLength() const { return m_maxSeqLength; }
    int curMaxSeqLength() const { return m_curMaxSeqLength; }
    int curMinSeqLength() const { return m_curMinSeqLength; }
    int curNumSeqs() const { return m_curNumSeqs; }
    int curPatterns() const { return m_curMaxSeqLength * m_parallelSequences; }
    const pattype_vector &patTypes() const { return m_patTypes; }
    const std::vector<char> &hostPatTypes() const { return m_hostPatTypes; }
    real_vector &outputs() { return m_outputs; }
    real_vector &outputErrors() { return m_outputErrors; }
    bl_ctx *ctx() const { return m_ctx; }

    virtual const std::string &type() const = 0;
    virtual void loadSequences(const data_sets::DataSetFraction &fraction);
    virtual void computeForwardPass() = 0;
    virtual void computeBackwardPass() = 0;
    virtual void exportLayer(helpers::JsonValue &layersArray) const;

    // packed host copies [curPatterns][size] (row padding stripped)
    std::vector<real_t> outputsToHost();
    std::vector<real_t> outputErrorsToHost();

protected:
    real_vector &_outputs() { return m_outputs; }
    std::vector<real_t> rowsToHost(const real_vector &v);

private:
    bl_ctx *m_ctx;
    std::string m_name;
    int m_size, m_ld;
    int m_parallelSequences, m_maxSeqLength;
    int m_curMaxSeqLength, m_curMinSeqLength, m_curNumSeqs;
    pattype_vector m_patTypes;
    std::vector<char> m_hostPatTypes;
    real_vector m_outputs, m_outputErrors;
};

class InputLayer : public Layer {
public:
    InputLayer(bl_ctx *ctx, const helpers::JsonValue &layerChild, int parallelSequences, int maxSeqLength);
    virtual const std::string &type() const;
    virtual void loadSequences(const data_sets::DataSetFraction &fraction);
    virtual void computeForwardPass() {}
    virtual void computeBackwardPass() {}
};

class TrainableLayer : public Layer {
public:
    TrainableLayer(const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection,
                   int inputWeightsPerBlock, int internalWeightsPerBlock, Layer &precedingLayer);
    Layer &precedingLayer() { return m_precedingLayer; }
    const Layer &precedingLayer() const { return m_precedingLayer; }
    real_t bias() const { return m_bias; }
    real_t learningRate() const { return m_learningRate; }
    real_vector &weights() { return m_weights; }
    const real_vector &weights() const { return m_weights; }
    real_vector &weightUpdates() { return m_weightUpdates; }
    const real_vector &weightUpdates() const { return m_weightUpdates; }
    void setWeights(const real_t *hostWeights, size_t n);
    void injectWeightNoise(real_t sigma);
    virtual void exportWeights(helpers::JsonValue &weightsObject) const;
    virtual void exportLayer(helpers::JsonValue &layersArray) const;

protected:
    real_vector &_weightUpdates() { return m_weightUpdates; }
    bool precedingIsTrainable() const { return m_precedingTrainable; }

private:
    Layer &m_precedingLayer;
    bool m_precedingTrainable;
    const int m_inputWeightsPerBlock, m_internalWeightsPerBlock;
    const real_t m_bias, m_learningRate;
    real_vector m_weights, m_weightUpdates;
};

class FeedForwardLayer : public TrainableLayer {
public:
    FeedForwardLayer(int act, const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual void computeForwardPass();
    virtual void computeBackwardPass();
protected:
    int m_act;
};

class SoftmaxLayer : public FeedForwardLayer {
public:
    SoftmaxLayer(const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual void computeForwardPass();
    virtual void computeBackwardPass();
};

class LstmLayer : public TrainableLayer {
public:
    LstmLayer(const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection, Layer &precedingLayer, bool bidirectional = false);
    virtual ~LstmLayer();
    virtual const std::string &type() const;
    bool isBidirectional() const { return m_isBidirectional; }
    // reference accessors (LstmLayer.hpp:169-232): [curPatterns][size] host copies; "Not implemented" for blstm
    std::vector<real_t> cellStates()       { return internal(0); }
    std::vector<real_t> cellStateErrors()  { return internal(1); }
    std::vector<real_t> netInputActs()     { return internal(2); }
    std::vector<real_t> inputGateActs()    { return internal(3); }
    std::vector<real_t> forgetGateActs()   { return internal(4); }
    std::vector<real_t> outputGateActs()   { return internal(5); }
    std::vector<real_t> netInputDeltas()   { return internal(6); }
    std::vector<real_t> inputGateDeltas()  { return internal(7); }
    std::vector<real_t> forgetGateDeltas() { return internal(8); }
    std::vector<real_t> outputGateDeltas() { return internal(9); }
    // per-direction variant (extension: lets tests diff blstm internals too)
    std::vector<real_t> internalOfDirection(int dir, int which);
    void planInfo(int *out8) const;
    bl_lstm_plan *plan() const { return m_plan; }
    virtual void computeForwardPass();
    virtual void computeBackwardPass();
private:
    std::vector<real_t> internal(int which);
    const bool m_isBidirectional;
    bl_lstm_plan *m_plan;
};

class PostOutputLayer : public Layer {
public:
    PostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer, int requiredSize, bool createOutputs = true);
    virtual void loadSequences(const data_sets::DataSetFraction &fraction);
    virtual real_t calculateError() = 0;
protected:
    real_vector &_targets() { return this->outputs(); }
    real_vector &_actualOutputs() { return m_precedingLayer.outputs(); }
    real_vector &_outputErrors() { return m_precedingLayer.outputErrors(); }
    Layer &preceding() { return m_precedingLayer; }
    device::Vector<float> m_devScalar;      // objective value on the device
private:
    Layer &m_precedingLayer;
};

class SsePostOutputLayer : public PostOutputLayer {
public:
    SsePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual real_t calculateError();
    virtual void computeForwardPass() {}
    virtual void computeBackwardPass();
};

class CePostOutputLayer : public PostOutputLayer {
public:
    CePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual real_t calculateError();
    virtual void computeForwardPass() {}
    virtual void computeBackwardPass();
};

// weightedsse / wf: twice as wide as the output layer, targets are (target, weight | filter input) pairs
// (layers/WeightedSsePostOutputLayer.cu, layers/SseMaskPostOutputLayer.cu)
class WeightedSsePostOutputLayer : public PostOutputLayer {
public:
    WeightedSsePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual real_t calculateError();
    virtual void computeForwardPass() {}
    virtual void computeBackwardPass();
};

class SseMaskPostOutputLayer : public PostOutputLayer {
public:
    SseMaskPostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual real_t calculateError();
    virtual void computeForwardPass() {}
    virtual void computeBackwardPass();
};

// layers/RmsePostOutputLayer.cu: the per-pattern RMSEs are computed in the forward pass and reused by both other calls
class RmsePostOutputLayer : public PostOutputLayer {
public:
    RmsePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual real_t calculateError();
    virtual void computeForwardPass();
    virtual void computeBackwardPass();
private:
    real_vector m_rmses;
};

// layers/BinaryClassificationLayer.cu: one logistic output, targets = the target classes copied as reals
class BinaryClassificationLayer : public PostOutputLayer {
public:
    BinaryClassificationLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual void loadSequences(const data_sets::DataSetFraction &fraction);
    virtual real_t calculateError();
    int countCorrectClassifications();
    virtual void computeForwardPass() {}
    virtual void computeBackwardPass();
private:
    void evaluate();
    device::Vector<int> m_devCorrect;
    std::vector<real_t> m_hostTargets;
    bool m_evaluated;
    real_t m_error; int m_correct;
};

class MulticlassClassificationLayer : public PostOutputLayer {
public:
    MulticlassClassificationLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer);
    virtual const std::string &type() const;
    virtual void loadSequences(const data_sets::DataSetFraction &fraction);
    virtual real_t calculateError();
    int countCorrectClassifications();
    virtual void computeForwardPass() {}
    virtual void computeBackwardPass();
private:
    void evaluate();
    device::int_vector m_patTargetClasses;
    device::Vector<int> m_devCorrect;
    bool m_evaluated;
    real_t m_error; int m_correct;
};

} // namespace layers

// process-wide options the layers read (the reference's Configuration singleton: only what the hot path uses,
// Configuration.hpp:300-342, TrainableLayer.cu:103-125)
struct Configuration {
    unsigned randomSeed = 0;
    bool     weightsUniform = true;
    real_t   weightsUniformMin = -0.1f, weightsUniformMax = 0.1f;
    real_t   weightsNormalMean = 0.0f, weightsNormalSigma = 0.1f;
    static Configuration &instance();
};
