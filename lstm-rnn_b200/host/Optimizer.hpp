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

    // one fraction: loadSequences -> forward -> calculateError (+ countCorrect) -> backward (+ gradient
    // all-reduce) -> weight update.  The body of the while loop of Optimizer::_processDataSet (Optimizer.cu:46-97).
    StepResult trainFraction(const data_sets::DataSetFraction &frac, bool firstFraction = true);
    // forward + error only (validation / test sets)
    StepResult evalFraction(const data_sets::DataSetFraction &frac);
    // one pass over a data set; returns error / totalSequences and fills *classError (Optimizer.cu:99-101)
    real_t processDataSet(data_sets::DataSet &ds, bool calcWeightUpdates, real_t *classError);
    // applies the accumulated updates (batch mode) -- _updateWeights(), SteepestDescentOptimizer.cu:67-94
    void updateWeights();

    // ---- epoch loop with early stopping (Optimizer.cu:283-324) ----
    // data sets may be null / empty; maxEpochs < 0 = unlimited
    void setDataSets(data_sets::DataSet *trainingSet, data_sets::DataSet *validationSet, data_sets::DataSet *testSet,
                     int maxEpochs, int maxEpochsNoBest, int validateEvery, int testEvery);
    // trains one epoch, evaluates validation / test sets when due, keeps the best weights; true once training is finished
    // (the best weights are restored then)
    bool train();
    bool finished() const { return m_finished; }
    int currentEpoch() const { return m_curEpoch; }
    real_t lowestValidationError() const { return m_lowestValidationError; }
    int epochsSinceLowestValidationError() const { return m_epochsSinceLowestError; }
    real_t curTrainingError() const { return m_curTrainingError; }
    real_t curValidationError() const { return m_curValidationError; }
    real_t curTestError() const { return m_curTestError; }
    real_t curTrainingClassError() const { return m_curTrainingClassError; }
    real_t curValidationClassError() const { return m_curValidationClassError; }
    real_t curTestClassError() const { return m_curTestClassError; }

    // autosave state in the reference's JSON fields (Optimizer.cu:326-360, SteepestDescentOptimizer.cu:118-131):
    // epoch counters, current / lowest errors, "optimizer_best_weights" and "steepest_descent_optimizer_weight_deltas"
    void exportState(helpers::JsonDocument &jsonDoc);
    void importState(const helpers::JsonDocument &jsonDoc);

    // --weight_noise_sigma (Optimizer.cu:58-69, 84-86): Gaussian noise on every trainable layer's weights between the forward and
    // the backward pass of a training fraction; the clean weights are restored before the update
    void setWeightNoise(real_t sigma, unsigned seed) { m_weightNoiseSigma = sigma; m_weightNoiseSeed = seed; }

    real_t learningRate() const { return m_learningRate; }
    void setLearningRate(real_t lr) { m_learningRate = lr; }
    std::vector<std::vector<real_t>> weightDeltasToHost() const;

private:
    NeuralNetwork &m_nn;
    real_t m_learningRate, m_momentum;
    bool m_hybridOnlineBatch;
    std::vector<std::unique_ptr<device::real_vector>> m_curWeightUpdates;   // batch mode accumulators
    std::vector<std::unique_ptr<device::real_vector>> m_weightDeltas;       // momentum state
    std::vector<std::unique_ptr<device::real_vector>> m_bestWeights;        // Optimizer.cu:106-127
    std::vector<std::unique_ptr<device::real_vector>> m_origWeights;        // clean weights while the noise is injected
    real_t m_weightNoiseSigma = 0; unsigned m_weightNoiseSeed = 0; unsigned long long m_weightNoiseCounter = 0;
    device::real_vector m_stats;                                            // {error, correct/4096, correct%4096} summed over ranks

    data_sets::DataSet *m_trainingSet = nullptr, *m_validationSet = nullptr, *m_testSet = nullptr;
    int m_maxEpochs = -1, m_maxEpochsNoBest = 20, m_validateEvery = 1, m_testEvery = 1;
    bool m_finished = false;
    int m_curEpoch = 0, m_epochsSinceLowestError = 0;
    real_t m_lowestValidationError;
    real_t m_curTrainingError, m_curValidationError, m_curTestError;
    real_t m_curTrainingClassError = 0, m_curValidationClassError = 0, m_curTestClassError = 0;

    void storeWeights();
    void restoreWeights();
};

} // namespace optimizers
