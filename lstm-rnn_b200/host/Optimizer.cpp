#include "Optimizer.hpp"
#include <cmath>
#include <limits>
#include <stdexcept>

using device::check;

namespace optimizers {

SteepestDescentOptimizer::SteepestDescentOptimizer(NeuralNetwork &nn, real_t learningRate, real_t momentum, bool hybridOnlineBatch)
    : m_nn(nn), m_learningRate(learningRate), m_momentum(momentum), m_hybridOnlineBatch(hybridOnlineBatch)
    , m_lowestValidationError(std::numeric_limits<real_t>::max())
    , m_curTrainingError(std::numeric_limits<real_t>::max()), m_curValidationError(std::numeric_limits<real_t>::max())
    , m_curTestError(std::numeric_limits<real_t>::max())
{
    m_stats.allocate(nn.ctx(), 4, true);
    for (const auto &layer : nn.layers()) {
        layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer.get());
        const size_t n = tl ? tl->weights().size() : 0;
        m_weightDeltas.emplace_back(new device::real_vector(nn.ctx(), n, true));                       // zeros, SteepestDescentOptimizer.cu:105-108
        m_curWeightUpdates.emplace_back(new device::real_vector(nn.ctx(), hybridOnlineBatch ? 0 : n, true));
    }
    storeWeights();          // best weights = the initial weights until a validation pass says otherwise (Optimizer.cu:106-127)
}

void SteepestDescentOptimizer::updateWeights()
{
    bl_ctx *ctx = m_nn.ctx();
    for (size_t i = 1; i + 1 < m_nn.layers().size(); ++i) {
        layers::TrainableLayer *layer = dynamic_cast<layers::TrainableLayer *>(m_nn.layers()[i].get());
        if (!layer) continue;
        real_t lr = m_learningRate;
        if (layer->learningRate() >= 0.0) lr = layer->learningRate();                                  // SteepestDescentOptimizer.cu:78-80
        // in hybrid online/batch mode the reference first copies weightUpdates() into m_curWeightUpdates (Optimizer.cu:82);
        // the update reads the layer's buffer directly instead -- same values, one HBM pass less
        const real_t *grad = m_hybridOnlineBatch ? layer->weightUpdates().data() : m_curWeightUpdates[i]->data();
        check(ctx, bl_sgd_update(ctx, layer->weights().size(), lr, m_momentum, layer->weights().data(), grad, m_weightDeltas[i]->data()));
    }
}

StepResult SteepestDescentOptimizer::evalFraction(const data_sets::DataSetFraction &frac)
{
    StepResult r{0, 0, frac.validFrames()};
    if (frac.numSequences() == 0) return r;
    m_nn.loadSequences(frac);
    m_nn.computeForwardPass();
    r.error = m_nn.calculateError();
    layers::MulticlassClassificationLayer *mc = dynamic_cast<layers::MulticlassClassificationLayer *>(&m_nn.postOutputLayer());
    if (mc) r.correct = mc->countCorrectClassifications();
    layers::BinaryClassificationLayer *bc = dynamic_cast<layers::BinaryClassificationLayer *>(&m_nn.postOutputLayer());
    if (bc) r.correct = bc->countCorrectClassifications();                                       // Optimizer.cu:52-55
    return r;
}

StepResult SteepestDescentOptimizer::trainFraction(const data_sets::DataSetFraction &frac, bool firstFraction)
{
    bl_ctx *ctx = m_nn.ctx();
    StepResult r{0, 0, frac.validFrames()};
    if (frac.numSequences() == 0) {
        m_nn.contributeZeroGradients();           // empty shard of a data-parallel fraction
        if (m_weightNoiseSigma > 0)                 // keep the noise counter in step with the ranks that do compute
            for (const auto &layer : m_nn.layers()) {
                layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer.get());
                if (tl) m_weightNoiseCounter += tl->weights().size();
            }
    } else {
        r = evalFraction(frac);
        const bool noise = m_weightNoiseSigma > 0;
        if (noise) {                                                              // Optimizer.cu:58-69
            if (m_origWeights.empty())
                for (const auto &layer : m_nn.layers()) {
                    layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer.get());
                    m_origWeights.emplace_back(new device::real_vector(ctx, tl ? tl->weights().size() : 0, false));
                }
            for (size_t i = 1; i + 1 < m_nn.layers().size(); ++i) {
                layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(m_nn.layers()[i].get());
                if (!tl) continue;
                const size_t n = tl->weights().size();
                check(ctx, bl_memcpy_d2d(ctx, m_origWeights[i]->data(), tl->weights().data(), n * sizeof(real_t)));
                check(ctx, bl_add_gaussian_noise(ctx, n, m_weightNoiseSigma, m_weightNoiseSeed, m_weightNoiseCounter, tl->weights().data()));
                m_weightNoiseCounter += n;
            }
        }
        m_nn.computeBackwardPass();
        if (noise)                                                                // restore before the update (Optimizer.cu:84-86)
            for (size_t i = 1; i + 1 < m_nn.layers().size(); ++i) {
                layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(m_nn.layers()[i].get());
                if (tl) check(ctx, bl_memcpy_d2d(ctx, tl->weights().data(), m_origWeights[i]->data(), tl->weights().size() * sizeof(real_t)));
            }
    }
    m_nn.joinGradients();
    if (!m_hybridOnlineBatch) {
        for (size_t i = 1; i + 1 < m_nn.layers().size(); ++i) {
            layers::TrainableLayer *layer = dynamic_cast<layers::TrainableLayer *>(m_nn.layers()[i].get());
            if (!layer) continue;
            const size_t n = layer->weightUpdates().size();
            if (firstFraction) check(ctx, bl_memcpy_d2d(ctx, m_curWeightUpdates[i]->data(), layer->weightUpdates().data(), n * sizeof(real_t)));
            else check(ctx, bl_vector_add(ctx, n, layer->weightUpdates().data(), m_curWeightUpdates[i]->data()));
        }
    } else {
        updateWeights();
    }
    return r;
}

real_t SteepestDescentOptimizer::processDataSet(data_sets::DataSet &ds, bool calcWeightUpdates, real_t *classError)
{
    real_t error = 0;
    *classError = (real_t)ds.totalTimesteps();
    std::shared_ptr<data_sets::DataSetFraction> frac;
    bool firstFraction = true;
    while ((frac = ds.getNextFraction())) {
        const StepResult r = calcWeightUpdates ? trainFraction(*frac, firstFraction) : evalFraction(*frac);
        error += r.error;
        *classError -= (real_t)r.correct;
        firstFraction = false;
    }
    if (calcWeightUpdates && !m_hybridOnlineBatch) updateWeights();
    check(m_nn.ctx(), bl_sync(m_nn.ctx()));
    if (m_nn.communicator()) {
        // every rank saw only its columns of each fraction: sum the statistics (the count is split so that fp32 stays exact)
        const long missed = std::lround((double)*classError);
        real_t st[4] = {error, (real_t)(missed / 4096), (real_t)(missed % 4096), 0};
        long base = 0;
        if (ds.totalTimesteps() > 0) base = (long)ds.totalTimesteps();
        m_stats.fromHost(st, 4);
        check(m_nn.ctx(), bl_allreduce_sum_f32(m_nn.communicator(), m_stats.data(), 4));
        check(m_nn.ctx(), bl_comm_join(m_nn.communicator()));
        m_stats.toHost(st, 4);
        int world = 1;
        bl_comm_info(m_nn.communicator(), nullptr, &world);
        error = st[0];
        // each rank started from totalTimesteps and subtracted its own correct count
        *classError = (real_t)((double)st[1] * 4096.0 + (double)st[2] - (double)(world - 1) * (double)base);
    }
    error /= ds.totalSequences();
    *classError /= (real_t)ds.totalTimesteps();
    return error;
}

void SteepestDescentOptimizer::setDataSets(data_sets::DataSet *trainingSet, data_sets::DataSet *validationSet, data_sets::DataSet *testSet,
                                           int maxEpochs, int maxEpochsNoBest, int validateEvery, int testEvery)
{
    m_trainingSet = trainingSet; m_validationSet = validationSet; m_testSet = testSet;
    m_maxEpochs = maxEpochs; m_maxEpochsNoBest = maxEpochsNoBest; m_validateEvery = validateEvery; m_testEvery = testEvery;
}

void SteepestDescentOptimizer::storeWeights()
{
    bl_ctx *ctx = m_nn.ctx();
    if (m_bestWeights.empty())
        for (const auto &layer : m_nn.layers()) {
            layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer.get());
            m_bestWeights.emplace_back(new device::real_vector(ctx, tl ? tl->weights().size() : 0, false));
        }
    for (size_t i = 0; i < m_nn.layers().size(); ++i) {
        layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(m_nn.layers()[i].get());
        if (tl) check(ctx, bl_memcpy_d2d(ctx, m_bestWeights[i]->data(), tl->weights().data(), tl->weights().size() * sizeof(real_t)));
    }
}

void SteepestDescentOptimizer::restoreWeights()
{
    bl_ctx *ctx = m_nn.ctx();
    for (size_t i = 0; i < m_nn.layers().size(); ++i) {
        layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(m_nn.layers()[i].get());
        if (tl) check(ctx, bl_memcpy_d2d(ctx, tl->weights().data(), m_bestWeights[i]->data(), tl->weights().size() * sizeof(real_t)));
    }
    check(ctx, bl_sync(ctx));
}

bool SteepestDescentOptimizer::train()
{
    if (!m_trainingSet) throw std::runtime_error("Optimizer::train: no training set");
    const bool haveVal = m_validationSet && !m_validationSet->empty();
    const bool haveTest = m_testSet && !m_testSet->empty();
    if (!m_finished) {
        ++m_curEpoch;
        m_curTrainingError = processDataSet(*m_trainingSet, true, &m_curTrainingClassError);
        if (haveVal && m_curEpoch % m_validateEvery == 0) {
            m_curValidationError = processDataSet(*m_validationSet, false, &m_curValidationClassError);
            if (m_curValidationError < m_lowestValidationError) {
                m_lowestValidationError = m_curValidationError;
                m_epochsSinceLowestError = 0;
                storeWeights();
            } else {
                m_epochsSinceLowestError += m_validateEvery;
            }
        } else if (!haveVal) {
            m_epochsSinceLowestError = 0;
            storeWeights();
        }
        if (haveTest && m_curEpoch % m_testEvery == 0)
            m_curTestError = processDataSet(*m_testSet, false, &m_curTestClassError);
        if (m_epochsSinceLowestError >= m_maxEpochsNoBest || (m_maxEpochs >= 0 && m_curEpoch >= m_maxEpochs)) {
            restoreWeights();
            m_finished = true;
        }
    }
    return m_finished;
}

namespace {
void exportVectors(helpers::JsonDocument &doc, const char *name, const std::vector<std::unique_ptr<device::real_vector>> &vs)
{
    helpers::JsonValue arr = helpers::JsonValue::makeArray();                     // Optimizer.cu:107-122: one array per layer
    for (const auto &v : vs) {
        helpers::JsonValue a = helpers::JsonValue::makeArray();
        const std::vector<real_t> h = v->toHost();
        a.Reserve(h.size());
        for (real_t x : h) a.PushBack(helpers::JsonValue::makeNumber(x));
        arr.PushBack(a);
    }
    doc.member(name) = arr;
}

void importVectors(const helpers::JsonDocument &doc, const char *name, std::vector<std::unique_ptr<device::real_vector>> &vs)
{
    if (!doc.HasMember(name) || !doc[name].IsArray())                             // Optimizer.cu:124-150, same messages
        throw std::runtime_error(std::string("Array '") + name + "' is missing or has the wrong type");
    if (doc[name].Size() != vs.size()) throw std::runtime_error(std::string("Array '") + name + "' has a wrong size");
    for (size_t i = 0; i < vs.size(); ++i) {
        const helpers::JsonValue &a = doc[name].at(i);
        if (!a.IsArray()) throw std::runtime_error(std::string("Object in '") + name + "' is not an array");
        if (a.Size() != vs[i]->size()) throw std::runtime_error(std::string("Subarray in '") + name + "' has a wrong size");
        std::vector<real_t> h(a.Size());
        for (size_t j = 0; j < h.size(); ++j) h[j] = (real_t)a.at(j).GetDouble();
        if (!h.empty()) vs[i]->fromHost(h.data(), h.size());
    }
}

template <typename T> T checkedGet(const helpers::JsonDocument &doc, const char *name);
template <> bool checkedGet<bool>(const helpers::JsonDocument &doc, const char *name)
{
    if (!doc.HasMember(name) || !doc[name].IsBool()) throw std::runtime_error(std::string("Missing or invalid value '") + name + "'");
    return doc[name].GetBool();
}
template <> int checkedGet<int>(const helpers::JsonDocument &doc, const char *name)
{
    if (!doc.HasMember(name) || !doc[name].IsNumber()) throw std::runtime_error(std::string("Missing or invalid value '") + name + "'");
    return doc[name].GetInt();
}
template <> real_t checkedGet<real_t>(const helpers::JsonDocument &doc, const char *name)
{
    if (!doc.HasMember(name) || !doc[name].IsNumber()) throw std::runtime_error(std::string("Missing or invalid value '") + name + "'");
    return (real_t)doc[name].GetDouble();
}
} // namespace

void SteepestDescentOptimizer::exportState(helpers::JsonDocument &doc)
{
    using helpers::JsonValue;
    doc.member("optimizer_finished") = JsonValue::makeBool(m_finished);
    doc.member("optimizer_cur_epoch") = JsonValue::makeNumber(m_curEpoch, true);
    doc.member("optimizer_epochs_since_lowest_error") = JsonValue::makeNumber(m_epochsSinceLowestError, true);
    doc.member("optimizer_lowest_validation_error") = JsonValue::makeNumber(m_lowestValidationError);
    doc.member("optimizer_cur_training_error") = JsonValue::makeNumber(m_curTrainingError);
    doc.member("optimizer_cur_validation_error") = JsonValue::makeNumber(m_curValidationError);
    doc.member("optimizer_cur_test_error") = JsonValue::makeNumber(m_curTestError);
    doc.member("optimizer_cur_training_class_error") = JsonValue::makeNumber(m_curTrainingClassError);
    doc.member("optimizer_cur_validation_class_error") = JsonValue::makeNumber(m_curValidationClassError);
    doc.member("optimizer_cur_test_class_error") = JsonValue::makeNumber(m_curTestClassError);
    exportVectors(doc, "optimizer_best_weights", m_bestWeights);
    exportVectors(doc, "steepest_descent_optimizer_weight_deltas", m_weightDeltas);
}

void SteepestDescentOptimizer::importState(const helpers::JsonDocument &doc)
{
    m_finished = checkedGet<bool>(doc, "optimizer_finished");
    m_curEpoch = checkedGet<int>(doc, "optimizer_cur_epoch");
    m_epochsSinceLowestError = checkedGet<int>(doc, "optimizer_epochs_since_lowest_error");
    m_lowestValidationError = checkedGet<real_t>(doc, "optimizer_lowest_validation_error");
    m_curTrainingError = checkedGet<real_t>(doc, "optimizer_cur_training_error");
    m_curValidationError = checkedGet<real_t>(doc, "optimizer_cur_validation_error");
    m_curTestError = checkedGet<real_t>(doc, "optimizer_cur_test_error");
    m_curTrainingClassError = checkedGet<real_t>(doc, "optimizer_cur_training_class_error");
    m_curValidationClassError = checkedGet<real_t>(doc, "optimizer_cur_validation_class_error");
    m_curTestClassError = checkedGet<real_t>(doc, "optimizer_cur_test_class_error");
    importVectors(doc, "optimizer_best_weights", m_bestWeights);
    importVectors(doc, "steepest_descent_optimizer_weight_deltas", m_weightDeltas);
    check(m_nn.ctx(), bl_sync(m_nn.ctx()));
}

std::vector<std::vector<real_t>> SteepestDescentOptimizer::weightDeltasToHost() const
{
    std::vector<std::vector<real_t>> out;
    for (const auto &v : m_weightDeltas) out.push_back(v->toHost());
    return out;
}

} // namespace optimizers
