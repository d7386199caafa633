#include "DataSet.hpp"
#include <algorithm>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>

namespace data_sets {

DataSetFraction *DataSetFraction::fromPacked(bl_ctx *ctx, int S, int T, int Tmin, int numSeqs, const int *seqLengths, int P, int O,
                                             const real_t *inputs, const char *patTypes, const int *targetClasses, const real_t *targets)
{
    DataSetFraction *f = new DataSetFraction;
    f->m_inputPatternSize = P; f->m_outputPatternSize = O; f->m_maxSeqLength = T; f->m_minSeqLength = Tmin; f->m_parallelSequences = S;
    for (int i = 0; i < numSeqs; ++i) {
        seq_info_t si; si.originalSeqIdx = i; si.length = seqLengths ? seqLengths[i] : T; si.seqTag = "seq";
        f->m_seqInfo.push_back(si);
    }
    const size_t n = (size_t)T * S;
    f->m_inputs.resize(ctx, n * P, 0);
    std::memcpy(f->m_inputs.data(), inputs, n * P * sizeof(real_t));
    f->m_patTypes.resize(ctx, n, PATTYPE_NONE);
    std::memcpy(f->m_patTypes.data(), patTypes, n);
    if (targetClasses) { f->m_targetClasses.resize(ctx, n, -1); std::memcpy(f->m_targetClasses.data(), targetClasses, n * sizeof(int)); }
    if (targets) { f->m_outputs.resize(ctx, n * O, 0); std::memcpy(f->m_outputs.data(), targets, n * O * sizeof(real_t)); }
    return f;
}

namespace {
std::mutex g_poolMutex;
std::multimap<size_t, void *> &pool() { static std::multimap<size_t, void *> *p = new std::multimap<size_t, void *>; return *p; }   // never destroyed: outlives the CUDA context
}

void *DataSetFraction::PinnedPool::get(bl_ctx *ctx, size_t bytes, size_t *capacity)
{
    if (bytes == 0) bytes = 1;
    {
        std::lock_guard<std::mutex> lock(g_poolMutex);
        auto it = pool().lower_bound(bytes);
        if (it != pool().end() && it->first <= 2 * bytes + 4096) { void *p = it->second; *capacity = it->first; pool().erase(it); return p; }
    }
    void *p = nullptr;
    device::check(ctx, bl_malloc_host(ctx, &p, bytes));
    *capacity = bytes;
    return p;
}

void DataSetFraction::PinnedPool::put(void *ptr, size_t capacity)
{
    std::lock_guard<std::mutex> lock(g_poolMutex);
    pool().emplace(capacity, ptr);
}

static bool comp_seqs(const DataSet::sequence_t &a, const DataSet::sequence_t &b) { return a.length < b.length; }   // DataSet.cpp:165-168

DataSet::DataSet(bl_ctx *ctx, int numSeqs, const int *seqLengths, int P, int O, const real_t *inputs, const int *targetClasses,
                 const real_t *targets, int parSeq, int truncSeqLength, bool trainingMode, int rank, int world)
    : m_ctx(ctx), m_isClassificationData(targetClasses != nullptr), m_parallelSequences(parSeq), m_rank(rank), m_world(world)
    , m_totalSequences(0), m_totalTimesteps(0)
    , m_minSeqLength(std::numeric_limits<int>::max()), m_maxSeqLength(std::numeric_limits<int>::min())
    , m_inputPatternSize(P), m_outputPatternSize(O), m_curFirstSeqIdx(-1)
    , m_fractionShuffling(false), m_sequenceShuffling(false), m_noiseDeviation(0)
{
    if ((targetClasses == nullptr) == (targets == nullptr))
        throw std::runtime_error("DataSet needs exactly one of target classes / target patterns");
    if (parSeq < 1 || world < 1 || rank < 0 || rank >= world)
        throw std::runtime_error("DataSet: bad parallel_sequences / rank / world");
    size_t begin = 0;
    for (int i = 0; i < numSeqs; ++i) {
        int seqLength = seqLengths[i];
        m_totalTimesteps += seqLength;                                            // DataSet.cpp:523 (original frames)
        int k = 0;
        while (seqLength > 0) {                                                   // DataSet.cpp:527-542
            sequence_t seq;
            seq.originalSeqIdx = k;
            seq.sourceSeq = i;
            if (truncSeqLength > 0 && seqLength > 1.5 * truncSeqLength)
                seq.length = std::min(truncSeqLength, seqLength);
            else
                seq.length = seqLength;
            seq.seqTag = "seq" + std::to_string(i);
            seq.inputsBegin = begin; seq.targetsBegin = begin;
            m_sequences.push_back(seq);
            begin += (size_t)seq.length;
            seqLength -= seq.length;
            ++k;
        }
    }
    for (const sequence_t &s : m_sequences) {
        m_minSeqLength = std::min(m_minSeqLength, s.length);
        m_maxSeqLength = std::max(m_maxSeqLength, s.length);
    }
    m_inputs.assign(inputs, inputs + begin * P);
    if (targetClasses) m_targetClasses.assign(targetClasses, targetClasses + begin);
    else m_targets.assign(targets, targets + begin * O);
    m_totalSequences = (int)m_sequences.size();                                   // DataSet.cpp:602 (chunks)
    if (trainingMode)
        std::sort(m_sequences.begin(), m_sequences.end(), comp_seqs);             // DataSet.cpp:603-605
}

DataSet::~DataSet()
{
    if (m_pending.valid()) m_pending.wait();
}

void DataSet::setSequenceTags(const std::vector<std::string> &tags)
{
    for (sequence_t &s : m_sequences)
        if (s.sourceSeq < (int)tags.size()) s.seqTag = tags[s.sourceSeq];
}

void DataSet::setShuffling(bool fractionShuffling, bool sequenceShuffling, unsigned seed)
{
    m_fractionShuffling = fractionShuffling; m_sequenceShuffling = sequenceShuffling;
    m_shuffleGen.seed(seed);
}

void DataSet::setInputNoise(real_t sigma, unsigned seed)
{
    m_noiseDeviation = sigma;
    m_noiseGen.seed(seed + 7919u * (unsigned)m_rank);                             // independent noise per rank
}

void DataSet::setContext(int left, int right, int outputTimeLag)
{
    if (left < 0 || right < 0 || outputTimeLag < 0) throw std::runtime_error("DataSet: negative context / time lag");
    if (m_pending.valid()) m_pending.wait();
    m_contextLeft = left; m_contextRight = right; m_outputLag = outputTimeLag;
}

void DataSet::shuffleSequences()
{
    std::shuffle(m_sequences.begin(), m_sequences.end(), m_shuffleGen);           // DataSet.cpp:225-229
}

void DataSet::shuffleFractions()
{
    // blocks of one GLOBAL fraction (S per rank x world) keep their members, the blocks are permuted (DataSet.cpp:231-246)
    const size_t gs = (size_t)m_parallelSequences * m_world;
    std::vector<std::vector<sequence_t>> fractions;
    for (size_t i = 0; i < m_sequences.size(); ++i) {
        if (i % gs == 0) fractions.resize(fractions.size() + 1);
        fractions.back().push_back(m_sequences[i]);
    }
    std::shuffle(fractions.begin(), fractions.end(), m_shuffleGen);
    m_sequences.clear();
    for (const auto &f : fractions) m_sequences.insert(m_sequences.end(), f.begin(), f.end());
}

std::shared_ptr<DataSetFraction> DataSet::makeFirstFraction()
{
    if (m_sequenceShuffling) shuffleSequences();                                  // DataSet.cpp:416-426
    if (m_fractionShuffling) shuffleFractions();
    return makeFraction(0);
}

int DataSet::numFractions() const
{
    const int gs = m_parallelSequences * m_world;
    return ((int)m_sequences.size() + gs - 1) / gs;
}

std::shared_ptr<DataSetFraction> DataSet::makeFraction(int firstSeqIdx)
{
    const int S = m_parallelSequences, P = m_inputPatternSize, O = m_outputPatternSize;
    const int ctx = m_contextLeft + m_contextRight + 1, PF = P * ctx;
    std::shared_ptr<DataSetFraction> frac(new DataSetFraction);
    frac->m_inputPatternSize = PF;
    frac->m_outputPatternSize = O;
    frac->m_parallelSequences = S;
    frac->m_maxSeqLength = std::numeric_limits<int>::min();
    frac->m_minSeqLength = std::numeric_limits<int>::max();
    const int first = firstSeqIdx + m_rank * S;
    for (int seqIdx = first; seqIdx < first + S; ++seqIdx) {
        if (seqIdx < (int)m_sequences.size()) {
            frac->m_maxSeqLength = std::max(frac->m_maxSeqLength, m_sequences[seqIdx].length);
            frac->m_minSeqLength = std::min(frac->m_minSeqLength, m_sequences[seqIdx].length);
            DataSetFraction::seq_info_t si;
            si.originalSeqIdx = m_sequences[seqIdx].originalSeqIdx;
            si.length = m_sequences[seqIdx].length;
            si.seqTag = m_sequences[seqIdx].seqTag;
            frac->m_seqInfo.push_back(si);
        }
    }
    if (frac->m_seqInfo.empty()) { frac->m_maxSeqLength = 0; frac->m_minSeqLength = 0; return frac; }   // empty shard

    const size_t slots = (size_t)frac->m_maxSeqLength * S, maxSlots = (size_t)m_maxSeqLength * S;
    frac->m_inputs.resize(m_ctx, slots * PF, 0, maxSlots * PF);
    frac->m_patTypes.resize(m_ctx, slots, PATTYPE_NONE, maxSlots);
    if (m_isClassificationData) frac->m_targetClasses.resize(m_ctx, slots, -1, maxSlots);
    else frac->m_outputs.resize(m_ctx, slots * O, 0, maxSlots * O);

    std::vector<real_t> noisy;
    for (int i = 0; i < S; ++i) {
        if (first + i >= (int)m_sequences.size()) continue;
        const sequence_t &seq = m_sequences[first + i];
        const real_t *src = m_inputs.data() + seq.inputsBegin * P;
        if (m_noiseDeviation > 0) {                                               // noise on the sequence, before the splicing (:346-347, 250-266)
            std::normal_distribution<real_t> dist((real_t)0, m_noiseDeviation);
            noisy.assign(src, src + (size_t)seq.length * P);
            for (real_t &x : noisy) x += dist(m_noiseGen);
            src = noisy.data();
        }
        for (int t = 0; t < seq.length; ++t) {
            const size_t slot = (size_t)t * S + i;                                // DataSet.cpp:358
            real_t *dst = frac->m_inputs.data() + slot * PF;
            for (int off = -m_contextLeft; off <= m_contextRight; ++off, dst += P) {   // :348-363 (edges repeat the first / last frame)
                const int ts = std::min(std::max(t + off, 0), seq.length - 1);
                std::memcpy(dst, src + (size_t)ts * P, sizeof(real_t) * P);
            }
            if (m_isClassificationData)                                           // :370-377
                frac->m_targetClasses[slot] = (t >= m_outputLag) ? m_targetClasses[seq.targetsBegin + t - m_outputLag] : 0;
            else if (t >= m_outputLag)                                            // :380-393
                std::memcpy(frac->m_outputs.data() + slot * O, m_targets.data() + (seq.targetsBegin + t - m_outputLag) * O, sizeof(real_t) * O);
            else
                for (int o = 0; o < O; ++o) frac->m_outputs[slot * O + o] = 1.0f;
            frac->m_patTypes[slot] = (t == 0) ? PATTYPE_FIRST : (t == seq.length - 1) ? PATTYPE_LAST : PATTYPE_NORMAL;   // :397-406
        }
    }
    return frac;
}

std::shared_ptr<DataSetFraction> DataSet::getNextFraction()
{
    const int gs = m_parallelSequences * m_world;
    if (m_curFirstSeqIdx == -1) {
        m_pending = std::async(std::launch::async, &DataSet::makeFirstFraction, this);
        m_curFirstSeqIdx = 0;
    }
    std::shared_ptr<DataSetFraction> frac;
    if (m_curFirstSeqIdx < (int)m_sequences.size()) {
        frac = m_pending.get();
        m_curFirstSeqIdx += gs;
        if (m_curFirstSeqIdx < (int)m_sequences.size())
            m_pending = std::async(std::launch::async, &DataSet::makeFraction, this, m_curFirstSeqIdx);
        else
            m_pending = std::async(std::launch::async, &DataSet::makeFirstFraction, this);    // next epoch's first fraction
    } else {
        m_curFirstSeqIdx = 0;                                                     // DataSet.cpp:662-664
    }
    return frac;
}

} // namespace data_sets
