// In-memory data set + fraction builder with the reference's semantics (data_sets/DataSet.cpp:300-414 packing,
// :527-542 --truncate_seq chunking, :603-605 sort by length in training mode, :632-668 fraction iteration) and the
// data-parallel extension: rank r of W takes columns [r*S, (r+1)*S) of each global fraction of W*S sequences.
// The NetCDF front end (DataSet.cpp:44-144, 443-606) lives in NetCdf.hpp; it only fills the arrays handed in here.
#pragma once
#include <future>
#include <memory>
#include <random>
#include <string>
#include <vector>
#include "DataSetFraction.hpp"

namespace data_sets {

class DataSet {
public:
    struct sequence_t {
        int         originalSeqIdx;   // chunk index within the source sequence (DataSet.cpp:529)
        int         sourceSeq;        // index of the source sequence in the file(s)
        int         length;
        std::string seqTag;
        size_t      inputsBegin;      // offsets in patterns into the frame arrays
        size_t      targetsBegin;
    };

    // frames: inputs [totalFrames][P]; exactly one of targetClasses [totalFrames] / targets [totalFrames][O].
    // parSeq = parallel sequences PER RANK; truncSeqLength = --truncate_seq; trainingMode sorts by length.
    DataSet(bl_ctx *ctx, int numSeqs, const int *seqLengths, int P, int O, const real_t *inputs, const int *targetClasses,
            const real_t *targets, int parSeq, int truncSeqLength = 0, bool trainingMode = true, int rank = 0, int world = 1);

    ~DataSet();

    // per source sequence tags (the NetCDF seqTags variable); chunks of a truncated sequence share their source's tag
    void setSequenceTags(const std::vector<std::string> &tags);
    // --shuffle_sequences / --shuffle_fractions (DataSet.cpp:225-246, 416-426), applied before the first fraction of every
    // epoch.  Every rank must pass the same seed: the shuffle is of the GLOBAL sequence list, ranks then take their columns.
    void setShuffling(bool fractionShuffling, bool sequenceShuffling, unsigned seed);
    // --input_noise_sigma: Gaussian noise added to the inputs of every fraction (DataSet.cpp:250-266)
    void setInputNoise(real_t sigma, unsigned seed);
    // --input_left_context / --input_right_context: every input pattern becomes the concatenation of the frames
    // t-left .. t+right of its sequence, the first / last frame repeated at the edges (DataSet.cpp:302-304, 348-363);
    // --output_time_lag: the target of frame t is the target of frame t-lag, class 0 / value 1 for t < lag (:370-393)
    void setContext(int left, int right, int outputTimeLag);
    int fractionInputPatternSize() const { return m_inputPatternSize * (m_contextLeft + m_contextRight + 1); }
    // per-output statistics from the file (outputMeans / outputStdevs), used by --revert_std in forward-pass mode
    void setOutputStatistics(const std::vector<real_t> &means, const std::vector<real_t> &stdevs) { m_outputMeans = means; m_outputStdevs = stdevs; }
    const std::vector<real_t> &outputMeans() const { return m_outputMeans; }
    const std::vector<real_t> &outputStdevs() const { return m_outputStdevs; }

    bool isClassificationData() const { return m_isClassificationData; }
    bool empty() const { return m_totalTimesteps == 0; }
    int totalSequences() const { return m_totalSequences; }
    int totalTimesteps() const { return m_totalTimesteps; }
    int minSeqLength() const { return m_minSeqLength; }
    int maxSeqLength() const { return m_maxSeqLength; }
    int inputPatternSize() const { return m_inputPatternSize; }
    int outputPatternSize() const { return m_outputPatternSize; }
    int numFractions() const;
    const std::vector<sequence_t> &sequences() const { return m_sequences; }

    // next fraction of the epoch, or null once at the end of each epoch (then the iteration restarts).  Like the
    // reference (DataSet.cpp:632-668) the following fraction is packed by a worker thread while this one is computed.
    // With world > 1 a rank whose shard is empty gets a fraction with numSequences() == 0 and maxSeqLength() == 0.
    std::shared_ptr<DataSetFraction> getNextFraction();
    // fraction starting at global sequence index firstSeqIdx (this rank's shard of it)
    std::shared_ptr<DataSetFraction> makeFraction(int firstSeqIdx);

private:
    bl_ctx *m_ctx;
    bool m_isClassificationData;
    int m_parallelSequences, m_rank, m_world;
    int m_totalSequences, m_totalTimesteps, m_minSeqLength, m_maxSeqLength, m_inputPatternSize, m_outputPatternSize;
    std::vector<real_t> m_inputs, m_targets;
    std::vector<int> m_targetClasses;
    std::vector<sequence_t> m_sequences;
    int m_curFirstSeqIdx;
    bool m_fractionShuffling, m_sequenceShuffling;
    real_t m_noiseDeviation;
    int m_contextLeft = 0, m_contextRight = 0, m_outputLag = 0;
    std::mt19937 m_shuffleGen, m_noiseGen;
    std::vector<real_t> m_outputMeans, m_outputStdevs;
    std::future<std::shared_ptr<DataSetFraction>> m_pending;

    std::shared_ptr<DataSetFraction> makeFirstFraction();
    void shuffleSequences();
    void shuffleFractions();
};

} // namespace data_sets
