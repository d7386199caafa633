#include "Layers.hpp"
#include <cstring>
#include <random>

using device::check;

Configuration &Configuration::instance()
{
    static Configuration c;
    return c;
}

namespace layers {

// ------------------------------------------------------------------ Layer (layers/Layer.cpp:41-157)
Layer::Layer(bl_ctx *ctx, const helpers::JsonValue &layerChild, int parallelSequences, int maxSeqLength, bool createOutputs, bool padRows)
    : m_ctx(ctx)
    , m_name(layerChild.HasMember("name") ? layerChild["name"].GetString() : "")
    , m_size(layerChild.HasMember("size") ? layerChild["size"].GetInt() : 0)
    , m_ld(padRows ? device::paddedLd(layerChild.HasMember("size") ? layerChild["size"].GetInt() : 0)
                   : (layerChild.HasMember("size") ? layerChild["size"].GetInt() : 0))
    , m_parallelSequences(parallelSequences)
    , m_maxSeqLength(maxSeqLength)
    , m_curMaxSeqLength(0)
    , m_curMinSeqLength(0)
    , m_curNumSeqs(0)
{
    if (!layerChild.HasMember("name"))
        throw std::runtime_error("Missing value 'name' in layer description");
    if (m_name.empty())
        throw std::runtime_error("Empty layer name in layer description");
    if (!layerChild.HasMember("size"))
        throw std::runtime_error(std::string("Missing value 'size' in layer '") + m_name + "'");

    const size_t slots = (size_t)m_parallelSequences * m_maxSeqLength;
    if (createOutputs) {
        m_outputs.allocate(ctx, slots * m_ld);
        m_outputErrors.allocate(ctx, slots * m_ld);
    }
    m_patTypes.allocate(ctx, slots);
}

Layer::~Layer() {}

void Layer::loadSequences(const data_sets::DataSetFraction &fraction)
{
    if (fraction.parallelSequences() != m_parallelSequences)
        throw std::runtime_error("Fraction was packed for a different number of parallel sequences");
    if (fraction.maxSeqLength() > m_maxSeqLength)
        throw std::runtime_error("Fraction is longer than the network's maximum sequence length");
    m_curMaxSeqLength = fraction.maxSeqLength();
    m_curMinSeqLength = fraction.minSeqLength();
    m_curNumSeqs      = fraction.numSequences();
    const size_t n = (size_t)curPatterns();
    m_hostPatTypes.assign(fraction.patTypes().data(), fraction.patTypes().data() + n);
    m_patTypes.fromHost(fraction.patTypes().data(), n);
}

void Layer::exportLayer(helpers::JsonValue &layersArray) const
{
    if (!layersArray.IsArray())
        throw std::runtime_error("The JSON value is not an array");
    helpers::JsonValue o = helpers::JsonValue::makeObject();
    o.member("name") = helpers::JsonValue::makeString(name());
    o.member("type") = helpers::JsonValue::makeString(type());
    o.member("size") = helpers::JsonValue::makeNumber(size(), true);
    layersArray.PushBack(o);
}

std::vector<real_t> Layer::rowsToHost(const real_vector &v)
{
    const size_t n = (size_t)curPatterns();
    std::vector<real_t> out(n * m_size);
    if (n && m_size) {
        if (v.size() < n * m_ld) throw std::runtime_error("layer has no such buffer");
        check(m_ctx, bl_memcpy2d_d2h(m_ctx, out.data(), (size_t)m_size * sizeof(real_t), v.data(), (size_t)m_ld * sizeof(real_t),
                                     (size_t)m_size * sizeof(real_t), n));
        check(m_ctx, bl_sync(m_ctx));
    }
    return out;
}

std::vector<real_t> Layer::outputsToHost() { return rowsToHost(m_outputs); }
std::vector<real_t> Layer::outputErrorsToHost() { return rowsToHost(m_outputErrors); }

// ------------------------------------------------------------------ InputLayer (layers/InputLayer.cpp:31-60)
InputLayer::InputLayer(bl_ctx *ctx, const helpers::JsonValue &layerChild, int parallelSequences, int maxSeqLength)
    : Layer(ctx, layerChild, parallelSequences, maxSeqLength, true, /*padRows=*/false)
{
}

const std::string &InputLayer::type() const { static const std::string s("input"); return s; }

void InputLayer::loadSequences(const data_sets::DataSetFraction &fraction)
{
    if (fraction.inputPatternSize() != this->size())
        throw std::runtime_error(std::string("Input layer size of ") + std::to_string(this->size())
                                 + " != data input pattern size of " + std::to_string(fraction.inputPatternSize()));
    Layer::loadSequences(fraction);
    _outputs().fromHost(fraction.inputs().data(), (size_t)curPatterns() * size());      // ld() == size(): one contiguous copy
}

// ------------------------------------------------------------------ TrainableLayer (layers/TrainableLayer.cu:50-248)
TrainableLayer::TrainableLayer(const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection,
                               int inputWeightsPerBlock, int internalWeightsPerBlock, Layer &precedingLayer)
    : Layer(precedingLayer.ctx(), layerChild, precedingLayer.parallelSequences(), precedingLayer.maxSeqLength())
    , m_precedingLayer(precedingLayer)
    , m_precedingTrainable(dynamic_cast<TrainableLayer *>(&precedingLayer) != nullptr)
    , m_inputWeightsPerBlock(inputWeightsPerBlock)
    , m_internalWeightsPerBlock(internalWeightsPerBlock)
    , m_bias(layerChild.HasMember("bias") ? (real_t)layerChild["bias"].GetDouble() : 0)
    , m_learningRate(layerChild.HasMember("learningRate") ? (real_t)layerChild["learningRate"].GetDouble() : -1)
{
    if (!layerChild.HasMember("bias"))
        throw std::runtime_error(std::string("Missing value 'bias' in layer '") + this->name() + "'");

    std::vector<real_t> weights;
    const size_t P = (size_t)m_precedingLayer.size(), sz = (size_t)this->size();
    if (weightsSection && weightsSection->HasMember(this->name())) {
        const helpers::JsonValue &wc = (*weightsSection)[this->name()];
        if (!wc.IsObject())
            throw std::runtime_error(std::string("Weights section for layer '") + this->name() + "' is not an object");
        if (!wc.HasMember("input") || !wc["input"].IsArray())
            throw std::runtime_error(std::string("Missing array 'weights/") + this->name() + "/input'");
        if (!wc.HasMember("bias") || !wc["bias"].IsArray())
            throw std::runtime_error(std::string("Missing array 'weights/") + this->name() + "/bias'");
        if (!wc.HasMember("internal") || !wc["internal"].IsArray())
            throw std::runtime_error(std::string("Missing array 'weights/") + this->name() + "/internal'");
        const helpers::JsonValue &in = wc["input"], &bi = wc["bias"], &it = wc["internal"];
        if (in.Size() != sz * inputWeightsPerBlock * P)
            throw std::runtime_error(std::string("Invalid number of input weights for layer '") + this->name() + "'");
        if (bi.Size() != sz * inputWeightsPerBlock)
            throw std::runtime_error(std::string("Invalid number of bias weights for layer '") + this->name() + "'");
        if (it.Size() != sz * internalWeightsPerBlock)
            throw std::runtime_error(std::string("Invalid number of internal weights for layer '") + this->name() + "'");
        weights.reserve(in.Size() + bi.Size() + it.Size());
        for (size_t i = 0; i < in.Size(); ++i) weights.push_back((real_t)in.at(i).GetDouble());
        for (size_t i = 0; i < bi.Size(); ++i) weights.push_back((real_t)bi.at(i).GetDouble());
        for (size_t i = 0; i < it.Size(); ++i) weights.push_back((real_t)it.at(i).GetDouble());
    } else {
        // random init; one generator shared by all layers, seeded once (TrainableLayer.cu:108-112).  The stream is
        // std::mt19937 + our own scaling, not Boost's distributions: parity runs always pass explicit weights.
        weights.resize(sz * (inputWeightsPerBlock * (P + 1) + internalWeightsPerBlock));
        const Configuration &config = Configuration::instance();
        static std::mt19937 *gen = nullptr;
        if (!gen) { gen = new std::mt19937; gen->seed(config.randomSeed); }
        if (config.weightsUniform) {
            std::uniform_real_distribution<real_t> dist(0, config.weightsUniformMax - config.weightsUniformMin);
            for (size_t i = 0; i < weights.size(); ++i) weights[i] = dist(*gen) + config.weightsUniformMin;
        } else {
            std::normal_distribution<real_t> dist(config.weightsNormalMean, config.weightsNormalSigma);
            for (size_t i = 0; i < weights.size(); ++i) weights[i] = dist(*gen);
        }
    }
    m_weights.allocate(ctx(), weights.size(), false);
    m_weightUpdates.allocate(ctx(), weights.size(), true);
    if (!weights.empty()) {
        m_weights.fromHost(weights.data(), weights.size());
        check(ctx(), bl_sync(ctx()));                  // `weights` is a stack temporary
    }
}

void TrainableLayer::setWeights(const real_t *hostWeights, size_t n)
{
    if (n != m_weights.size()) throw std::runtime_error("setWeights: wrong number of weights for layer '" + name() + "'");
    m_weights.fromHost(hostWeights, n);
    check(ctx(), bl_sync(ctx()));
}

void TrainableLayer::injectWeightNoise(real_t sigma)
{
    static std::mt19937 *gen = nullptr;
    if (!gen) { gen = new std::mt19937; gen->seed(Configuration::instance().randomSeed); }
    std::normal_distribution<real_t> dist(0.0f, sigma);
    std::vector<real_t> w = m_weights.toHost();
    for (size_t i = 0; i < w.size(); ++i) w[i] += dist(*gen);
    setWeights(w.data(), w.size());
}

void TrainableLayer::exportWeights(helpers::JsonValue &weightsObject) const
{
    if (!weightsObject.IsObject())
        throw std::runtime_error("The JSON value is not an object");
    if (m_weights.empty())
        return;
    const std::vector<real_t> w = m_weights.toHost();
    const size_t nIn = (size_t)size() * m_inputWeightsPerBlock * m_precedingLayer.size();
    const size_t nBi = (size_t)size() * m_inputWeightsPerBlock;
    const size_t nIt = (size_t)size() * m_internalWeightsPerBlock;
    helpers::JsonValue in = helpers::JsonValue::makeArray(), bi = helpers::JsonValue::makeArray(), it = helpers::JsonValue::makeArray();
    in.Reserve(nIn); bi.Reserve(nBi); it.Reserve(nIt);
    for (size_t i = 0; i < nIn; ++i) in.PushBack(helpers::JsonValue::makeNumber(w[i]));
    for (size_t i = 0; i < nBi; ++i) bi.PushBack(helpers::JsonValue::makeNumber(w[nIn + i]));
    for (size_t i = 0; i < nIt; ++i) it.PushBack(helpers::JsonValue::makeNumber(w[nIn + nBi + i]));
    helpers::JsonValue sec = helpers::JsonValue::makeObject();
    sec.member("input") = in; sec.member("bias") = bi; sec.member("internal") = it;
    weightsObject.member(name()) = sec;
}

void TrainableLayer::exportLayer(helpers::JsonValue &layersArray) const
{
    Layer::exportLayer(layersArray);
    layersArray.at(layersArray.Size() - 1).member("bias") = helpers::JsonValue::makeNumber(m_bias);
}

// ------------------------------------------------------------------ FeedForwardLayer (layers/FeedForwardLayer.cu:107-224)
FeedForwardLayer::FeedForwardLayer(int act, const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection, Layer &precedingLayer)
    : TrainableLayer(layerChild, weightsSection, 1, 0, precedingLayer), m_act(act)
{
}

const std::string &FeedForwardLayer::type() const
{
    static const std::string t("feedforward_tanh"), l("feedforward_logistic"), i("feedforward_identity");
    return m_act == BL_ACT_TANH ? t : m_act == BL_ACT_LOGISTIC ? l : i;
}

void FeedForwardLayer::computeForwardPass()
{
    Layer &pl = precedingLayer();
    check(ctx(), bl_ff_forward(ctx(), m_act, pl.size(), size(), curPatterns(), bias(), weights().data(),
                               pl.outputs().data(), pl.ld(), _outputs().data(), ld()));
}

void FeedForwardLayer::computeBackwardPass()
{
    Layer &pl = precedingLayer();
    check(ctx(), bl_ff_backward(ctx(), m_act, pl.size(), size(), curPatterns(), bias(), weights().data(),
                                pl.outputs().data(), pl.ld(), outputs().data(), ld(), outputErrors().data(), ld(),
                                precedingIsTrainable() ? pl.outputErrors().data() : nullptr, pl.ld(), _weightUpdates().data()));
}

// ------------------------------------------------------------------ SoftmaxLayer (layers/SoftmaxLayer.cu:225-353)
SoftmaxLayer::SoftmaxLayer(const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection, Layer &precedingLayer)
    : FeedForwardLayer(BL_ACT_IDENTITY, layerChild, weightsSection, precedingLayer)
{
}

const std::string &SoftmaxLayer::type() const { static const std::string s("softmax"); return s; }

void SoftmaxLayer::computeForwardPass()
{
    FeedForwardLayer::computeForwardPass();
    check(ctx(), bl_softmax_forward(ctx(), size(), curPatterns(), patTypes().data(), _outputs().data(), ld()));
}

void SoftmaxLayer::computeBackwardPass()
{
    check(ctx(), bl_softmax_backward(ctx(), size(), curPatterns(), patTypes().data(), outputs().data(), ld(), outputErrors().data(), ld()));
    FeedForwardLayer::computeBackwardPass();
}

// ------------------------------------------------------------------ LstmLayer (layers/LstmLayer.cu:520-1051)
LstmLayer::LstmLayer(const helpers::JsonValue &layerChild, const helpers::JsonValue *weightsSection, Layer &precedingLayer, bool bidirectional)
    : TrainableLayer(layerChild, weightsSection, 4, (bidirectional ? 2 : 4) * helpers::safeJsonGetInt(layerChild, "size") + 3, precedingLayer)
    , m_isBidirectional(bidirectional)
    , m_plan(nullptr)
{
    if (m_isBidirectional && this->size() % 2 != 0)
        throw std::runtime_error("Cannot create a bidirectional layer with an odd layer size");
    check(ctx(), bl_lstm_plan_create(ctx(), precedingLayer.size(), size(), bidirectional ? 1 : 0, parallelSequences(), maxSeqLength(), bias(), &m_plan));
}

LstmLayer::~LstmLayer() { bl_lstm_plan_destroy(m_plan); }

const std::string &LstmLayer::type() const
{
    static const std::string su("lstm"), sb("blstm");
    return m_isBidirectional ? sb : su;
}

void LstmLayer::computeForwardPass()
{
    Layer &pl = precedingLayer();
    check(ctx(), bl_lstm_forward(m_plan, weights().data(), pl.outputs().data(), pl.ld(), patTypes().data(),
                                 curMaxSeqLength(), curMinSeqLength(), _outputs().data(), ld()));
}

void LstmLayer::computeBackwardPass()
{
    Layer &pl = precedingLayer();
    check(ctx(), bl_lstm_backward(m_plan, weights().data(), pl.outputs().data(), pl.ld(), outputs().data(), ld(),
                                  outputErrors().data(), ld(), patTypes().data(), curMaxSeqLength(), curMinSeqLength(),
                                  precedingIsTrainable() ? pl.outputErrors().data() : nullptr, pl.ld(), _weightUpdates().data()));
}

std::vector<real_t> LstmLayer::internalOfDirection(int dir, int which)
{
    const int H = size() / (m_isBidirectional ? 2 : 1);
    const size_t n = (size_t)curPatterns() * H;
    device::real_vector tmp(ctx(), n, false);
    check(ctx(), bl_lstm_get_internal(m_plan, dir, which, curMaxSeqLength(), tmp.data()));
    return tmp.toHost();
}

std::vector<real_t> LstmLayer::internal(int which)
{
    if (m_isBidirectional)
        throw std::runtime_error("Not implemented");             // LstmLayer.cu:646-734
    return internalOfDirection(0, which);
}

void LstmLayer::planInfo(int *out8) const { bl_lstm_plan_info(m_plan, out8); }

// ------------------------------------------------------------------ PostOutputLayer (layers/PostOutputLayer.cpp:49-79)
PostOutputLayer::PostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer, int requiredSize, bool createOutputs)
    : Layer(precedingLayer.ctx(), layerChild, precedingLayer.parallelSequences(), precedingLayer.maxSeqLength(), createOutputs, /*padRows=*/false)
    , m_devScalar(precedingLayer.ctx(), 1)
    , m_precedingLayer(precedingLayer)
{
    if (this->size() != requiredSize)
        throw std::runtime_error("Size mismatch: " + std::to_string(this->size()) + " vs. " + std::to_string(requiredSize));
}

void PostOutputLayer::loadSequences(const data_sets::DataSetFraction &fraction)
{
    if (fraction.outputPatternSize() != this->size())
        throw std::runtime_error(std::string("Output layer size of ") + std::to_string(this->size())
                                 + " != data target pattern size of " + std::to_string(fraction.outputPatternSize()));
    Layer::loadSequences(fraction);
    if (!this->_outputs().empty() && !fraction.outputs().empty())
        this->_outputs().fromHost(fraction.outputs().data(), (size_t)curPatterns() * size());   // ld() == size()
}

// ------------------------------------------------------------------ SSE / CE (layers/SsePostOutputLayer.cu, CePostOutputLayer.cu)
SsePostOutputLayer::SsePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer)
    : PostOutputLayer(layerChild, precedingLayer, precedingLayer.size()) {}
const std::string &SsePostOutputLayer::type() const { static const std::string s("sse"); return s; }

real_t SsePostOutputLayer::calculateError()
{
    check(ctx(), bl_sse_error(ctx(), size(), curPatterns(), patTypes().data(), _targets().data(), ld(),
                              _actualOutputs().data(), preceding().ld(), m_devScalar.data()));
    real_t e; m_devScalar.toHost(&e, 1);
    return e;
}

void SsePostOutputLayer::computeBackwardPass()
{
    check(ctx(), bl_sse_backward(ctx(), size(), curPatterns(), patTypes().data(), _targets().data(), ld(),
                                 _actualOutputs().data(), preceding().ld(), _outputErrors().data(), preceding().ld()));
}

CePostOutputLayer::CePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer)
    : PostOutputLayer(layerChild, precedingLayer, precedingLayer.size()) {}
const std::string &CePostOutputLayer::type() const { static const std::string s("ce"); return s; }

real_t CePostOutputLayer::calculateError()
{
    check(ctx(), bl_ce_error(ctx(), size(), curPatterns(), patTypes().data(), _targets().data(), ld(),
                             _actualOutputs().data(), preceding().ld(), m_devScalar.data()));
    real_t e; m_devScalar.toHost(&e, 1);
    return e;
}

void CePostOutputLayer::computeBackwardPass()
{
    check(ctx(), bl_ce_backward(ctx(), size(), curPatterns(), patTypes().data(), _targets().data(), ld(),
                                _actualOutputs().data(), preceding().ld(), _outputErrors().data(), preceding().ld()));
}

// ------------------------------------------------------------------ weightedsse / wf (layers/WeightedSsePostOutputLayer.cu:101-164,
// layers/SseMaskPostOutputLayer.cu:101-164): size() is twice the output layer's size
WeightedSsePostOutputLayer::WeightedSsePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer)
    : PostOutputLayer(layerChild, precedingLayer, precedingLayer.size() * 2) {}
const std::string &WeightedSsePostOutputLayer::type() const { static const std::string s("weightedsse"); return s; }

real_t WeightedSsePostOutputLayer::calculateError()
{
    check(ctx(), bl_weightedsse_error(ctx(), size() / 2, curPatterns(), patTypes().data(), _targets().data(), ld(),
                                      _actualOutputs().data(), preceding().ld(), m_devScalar.data()));
    real_t e; m_devScalar.toHost(&e, 1);
    return e;
}

void WeightedSsePostOutputLayer::computeBackwardPass()
{
    check(ctx(), bl_weightedsse_backward(ctx(), size() / 2, curPatterns(), patTypes().data(), _targets().data(), ld(),
                                         _actualOutputs().data(), preceding().ld(), _outputErrors().data(), preceding().ld()));
}

SseMaskPostOutputLayer::SseMaskPostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer)
    : PostOutputLayer(layerChild, precedingLayer, precedingLayer.size() * 2) {}
const std::string &SseMaskPostOutputLayer::type() const { static const std::string s("wf"); return s; }

real_t SseMaskPostOutputLayer::calculateError()
{
    check(ctx(), bl_ssemask_error(ctx(), size() / 2, curPatterns(), patTypes().data(), _targets().data(), ld(),
                                  _actualOutputs().data(), preceding().ld(), m_devScalar.data()));
    real_t e; m_devScalar.toHost(&e, 1);
    return e;
}

void SseMaskPostOutputLayer::computeBackwardPass()
{
    check(ctx(), bl_ssemask_backward(ctx(), size() / 2, curPatterns(), patTypes().data(), _targets().data(), ld(),
                                     _actualOutputs().data(), preceding().ld(), _outputErrors().data(), preceding().ld()));
}

// ------------------------------------------------------------------ rmse (layers/RmsePostOutputLayer.cu:103-170)
RmsePostOutputLayer::RmsePostOutputLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer)
    : PostOutputLayer(layerChild, precedingLayer, precedingLayer.size())
    , m_rmses(precedingLayer.ctx(), (size_t)precedingLayer.parallelSequences() * precedingLayer.maxSeqLength(), true) {}
const std::string &RmsePostOutputLayer::type() const { static const std::string s("rmse"); return s; }

void RmsePostOutputLayer::computeForwardPass()
{
    check(ctx(), bl_rmse_forward(ctx(), size(), curPatterns(), patTypes().data(), _targets().data(), ld(),
                                 _actualOutputs().data(), preceding().ld(), m_rmses.data()));
}

real_t RmsePostOutputLayer::calculateError()
{
    check(ctx(), bl_rmse_error(ctx(), curPatterns(), m_rmses.data(), m_devScalar.data()));
    real_t e; m_devScalar.toHost(&e, 1);
    return e;
}

void RmsePostOutputLayer::computeBackwardPass()
{
    check(ctx(), bl_rmse_backward(ctx(), size(), curPatterns(), m_rmses.data(), _targets().data(), ld(),
                                  _actualOutputs().data(), preceding().ld(), _outputErrors().data(), preceding().ld()));
}

// ------------------------------------------------------------------ BinaryClassificationLayer (BinaryClassificationLayer.cu:119-203)
BinaryClassificationLayer::BinaryClassificationLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer)
    : PostOutputLayer(layerChild, precedingLayer, precedingLayer.size())
    , m_devCorrect(precedingLayer.ctx(), 1)
    , m_evaluated(false), m_error(0), m_correct(0)
{
    if (this->size() != 1)
        throw std::runtime_error("The binary classification post output layer cannot be used for an output layer size != 1");
}

const std::string &BinaryClassificationLayer::type() const { static const std::string s("binary_classification"); return s; }

void BinaryClassificationLayer::loadSequences(const data_sets::DataSetFraction &fraction)
{
    PostOutputLayer::loadSequences(fraction);
    if (fraction.targetClasses().size() < (size_t)curPatterns())
        throw std::runtime_error("The data fraction carries no target classes");
    // the target classes are the real target values 0/1 (-1 for padded patterns), :155-162
    m_hostTargets.resize((size_t)curPatterns());
    for (size_t n = 0; n < m_hostTargets.size(); ++n) m_hostTargets[n] = (real_t)fraction.targetClasses()[n];
    _targets().fromHost(m_hostTargets.data(), m_hostTargets.size());           // ld() == size() == 1
    m_evaluated = false;
}

void BinaryClassificationLayer::evaluate()
{
    check(ctx(), bl_binary_error(ctx(), curPatterns(), patTypes().data(), _targets().data(), ld(), _actualOutputs().data(),
                                 preceding().ld(), m_devScalar.data(), m_devCorrect.data()));
    check(ctx(), bl_memcpy_d2h(ctx(), &m_error, m_devScalar.data(), sizeof(real_t)));
    check(ctx(), bl_memcpy_d2h(ctx(), &m_correct, m_devCorrect.data(), sizeof(int)));
    check(ctx(), bl_sync(ctx()));
    m_evaluated = true;
}

real_t BinaryClassificationLayer::calculateError()
{
    evaluate();
    return m_error;
}

int BinaryClassificationLayer::countCorrectClassifications()
{
    if (!m_evaluated) evaluate();
    m_evaluated = false;
    return m_correct;
}

void BinaryClassificationLayer::computeBackwardPass()
{
    check(ctx(), bl_binary_backward(ctx(), curPatterns(), patTypes().data(), _targets().data(), ld(), _actualOutputs().data(),
                                    preceding().ld(), _outputErrors().data(), preceding().ld()));
}

// ------------------------------------------------------------------ MulticlassClassificationLayer (MulticlassClassificationLayer.cu:141-240)
MulticlassClassificationLayer::MulticlassClassificationLayer(const helpers::JsonValue &layerChild, Layer &precedingLayer)
    : PostOutputLayer(layerChild, precedingLayer, precedingLayer.size(), false)
    , m_patTargetClasses(precedingLayer.ctx(), (size_t)precedingLayer.parallelSequences() * precedingLayer.maxSeqLength())
    , m_devCorrect(precedingLayer.ctx(), 1)
    , m_evaluated(false), m_error(0), m_correct(0)
{
    if (this->size() == 1)
        throw std::runtime_error("The multiclass classification post output layer cannot be used for an output layer size of 1");
}

const std::string &MulticlassClassificationLayer::type() const { static const std::string s("multiclass_classification"); return s; }

void MulticlassClassificationLayer::loadSequences(const data_sets::DataSetFraction &fraction)
{
    PostOutputLayer::loadSequences(fraction);
    if (fraction.targetClasses().size() < (size_t)curPatterns())
        throw std::runtime_error("The data fraction carries no target classes");
    m_patTargetClasses.fromHost(fraction.targetClasses().data(), (size_t)curPatterns());
    m_evaluated = false;
}

// error and correct-classification count come out of one fused pass; both scalars cross to the host together
void MulticlassClassificationLayer::evaluate()
{
    check(ctx(), bl_multiclass_error(ctx(), size(), curPatterns(), m_patTargetClasses.data(), _actualOutputs().data(),
                                     preceding().ld(), m_devScalar.data(), m_devCorrect.data()));
    check(ctx(), bl_memcpy_d2h(ctx(), &m_error, m_devScalar.data(), sizeof(real_t)));
    check(ctx(), bl_memcpy_d2h(ctx(), &m_correct, m_devCorrect.data(), sizeof(int)));
    check(ctx(), bl_sync(ctx()));
    m_evaluated = true;
}

real_t MulticlassClassificationLayer::calculateError()
{
    evaluate();
    return m_error;
}

int MulticlassClassificationLayer::countCorrectClassifications()
{
    if (!m_evaluated) evaluate();
    m_evaluated = false;
    return m_correct;
}

void MulticlassClassificationLayer::computeBackwardPass()
{
    check(ctx(), bl_multiclass_backward(ctx(), size(), curPatterns(), m_patTargetClasses.data(), _actualOutputs().data(),
                                        preceding().ld(), _outputErrors().data(), preceding().ld()));
}

} // namespace layers
