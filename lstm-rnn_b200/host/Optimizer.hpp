// Training step and epoch loop (optimizers/Optimizer.cu:37-104, 283-324; SteepestDescentOptimizer.cu:39-94).
#pragma once
#include <memory>
#include <vector>
#include "DataSet.hpp"
#include "NeuralNetwork.hpp"

namespace optimizers {

struct StepResult {
    real_t error;          // objective summed over the fraction (this rank's shard, or all ranks after reduceStats)
    int    correct;        // correctly classified timesteps (multiclass only)
    long   frames;         // valid timesteps processed
};

class SteepestDescentOptimizer {
public:
    // hybridOnlineBatch == the reference's --hybrid_online_batch / --stochastic: update after every fraction
    SteepestDescentOptimizer(NeuralNetwork &neuralNetwork, real_t learningRate, real_t momentum, bool hybridOnlineBatch = true);

    // one fraction: loadSequences -> forward -> calculateError (+ countCorrect) -> backward (+ overlapped gradient
    // all-reduce) -> weight update.  The body of the while loop of Optimizer::_processDataSet (Optimizer.cu:46-97).
    StepResult trainFraction(const data_sets::DataSetFraction &frac, bool firstFraction = true);
    // forward + error only (validation / test sets)
    StepResult evalFraction(const data_sets::DataSetFraction &frac);
    // one pass over a data set; returns error / totalSequences and fills *classError (Optimizer.cu:99-101)
    real_t processDataSet(data_sets::DataSet &ds, bool calcWeightUpdates, real_t *classError);
    // applies the accumulated updates (batch mode) -- _updateWeights(), SteepestDescentOptimizer.cu:67-94
    void updateWeights();

    real_t learningRate() const { return m_learningRate; }
    void setLearningRate(real_t lr) { m_learningRate = lr; }
    std::vector<std::vector<real_t>> weightDeltasToHost() const;

private:
    NeuralNetwork &m_nn;
    real_t m_learningRate, m_momentum;
    bool m_hybridOnlineBatch;
    std::vector<std::unique_ptr<device::real_vector>> m_curWeightUpdates;   // batch mode accumulators
    std::vector<std::unique_ptr<device::real_vector>> m_weightDeltas;       // momentum state
};

} // namespace optimizers
