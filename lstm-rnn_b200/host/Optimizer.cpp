#include "Optimizer.hpp"

using device::check;

namespace optimizers {

SteepestDescentOptimizer::SteepestDescentOptimizer(NeuralNetwork &nn, real_t learningRate, real_t momentum, bool hybridOnlineBatch)
    : m_nn(nn), m_learningRate(learningRate), m_momentum(momentum), m_hybridOnlineBatch(hybridOnlineBatch)
{
    for (const auto &layer : nn.layers()) {
        layers::TrainableLayer *tl = dynamic_cast<layers::TrainableLayer *>(layer.get());
        const size_t n = tl ? tl->weights().size() : 0;
        m_weightDeltas.emplace_back(new device::real_vector(nn.ctx(), n, true));                       // zeros, SteepestDescentOptimizer.cu:105-108
        m_curWeightUpdates.emplace_back(new device::real_vector(nn.ctx(), hybridOnlineBatch ? 0 : n, true));
    }
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
    return r;
}

StepResult SteepestDescentOptimizer::trainFraction(const data_sets::DataSetFraction &frac, bool firstFraction)
{
    bl_ctx *ctx = m_nn.ctx();
    StepResult r{0, 0, frac.validFrames()};
    if (frac.numSequences() == 0) {
        m_nn.contributeZeroGradients();           // empty shard of a data-parallel fraction
    } else {
        r = evalFraction(frac);
        m_nn.computeBackwardPass();
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
    error /= ds.totalSequences();
    *classError /= (real_t)ds.totalTimesteps();
    return error;
}

std::vector<std::vector<real_t>> SteepestDescentOptimizer::weightDeltasToHost() const
{
    std::vector<std::vector<real_t>> out;
    for (const auto &v : m_weightDeltas) out.push_back(v->toHost());
    return out;
}

} // namespace optimizers
