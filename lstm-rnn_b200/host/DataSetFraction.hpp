// A fraction (minibatch) of parallel sequences in the reference's packing: slot n = t*S + s, features fastest
// (data_sets/DataSetFraction.hpp:38-143; filled by data_sets::DataSet, DataSet.cpp:300-414).
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "Types.hpp"

namespace data_sets {

class DataSet;

class DataSetFraction {
    friend class DataSet;

public:
    struct seq_info_t {
        int         originalSeqIdx;
        int         length;
        std::string seqTag;
    };

    // host buffers; pinned when the fraction was built with a device context.  Pinned blocks come from a process-wide
    // pool (cudaMallocHost / cudaFreeHost cost milliseconds and cudaFreeHost synchronises the device): a block goes back
    // to the pool when its fraction dies.  NeuralNetwork::loadSequences copies out of these buffers asynchronously and leaves an
    // upload ticket with the fraction (noteUpload); the destructor waits for that point of the stream before the blocks are
    // returned, so a forward-only caller that never synchronises cannot have a queued copy read a recycled block.
    struct PinnedPool {
        static void *get(bl_ctx *ctx, size_t bytes, size_t *capacity);
        static void put(void *ptr, size_t capacity);
    };
    template <typename T>
    struct HostBuffer {
        T *ptr = nullptr; size_t n = 0, cap = 0; bl_ctx *ctx = nullptr; std::vector<T> pageable;
        HostBuffer() {}
        HostBuffer(const HostBuffer &) = delete;
        HostBuffer &operator=(const HostBuffer &) = delete;
        ~HostBuffer() { if (ctx && ptr) PinnedPool::put(ptr, cap); }
        // reserve >= count: blocks of one data set all have the capacity of its largest fraction, so they are interchangeable
        void resize(bl_ctx *c, size_t count, T fill, size_t reserve = 0)
        {
            n = count; ctx = c;
            if (c) { ptr = (T *)PinnedPool::get(c, std::max(count, reserve) * sizeof(T), &cap); std::fill(ptr, ptr + count, fill); }
            else { pageable.assign(count, fill); ptr = pageable.data(); }
        }
        const T *data() const { return ptr; }
        T *data() { return ptr; }
        size_t size() const { return n; }
        bool empty() const { return n == 0; }
        const T &operator[](size_t i) const { return ptr[i]; }
        T &operator[](size_t i) { return ptr[i]; }
    };

    DataSetFraction() : m_inputPatternSize(0), m_outputPatternSize(0), m_maxSeqLength(0), m_minSeqLength(0), m_parallelSequences(0) {}
    ~DataSetFraction() { if (m_uploadCtx) bl_upload_wait(m_uploadCtx, m_uploadTicket); }       // before the buffers go back to the pool
    // called by NeuralNetwork::loadSequences after the last asynchronous H2D copy out of this fraction's buffers
    void noteUpload(bl_ctx *ctx, unsigned long long ticket) const { m_uploadCtx = ctx; m_uploadTicket = ticket; }

    int inputPatternSize() const { return m_inputPatternSize; }
    int outputPatternSize() const { return m_outputPatternSize; }
    int maxSeqLength() const { return m_maxSeqLength; }
    int minSeqLength() const { return m_minSeqLength; }
    int numSequences() const { return (int)m_seqInfo.size(); }
    int parallelSequences() const { return m_parallelSequences; }
    const seq_info_t &seqInfo(int seqIdx) const { return m_seqInfo[seqIdx]; }
    const HostBuffer<char>   &patTypes() const { return m_patTypes; }
    const HostBuffer<real_t> &inputs() const { return m_inputs; }
    const HostBuffer<real_t> &outputs() const { return m_outputs; }
    const HostBuffer<int>    &targetClasses() const { return m_targetClasses; }

    // valid (non-padded) timesteps in this fraction
    long validFrames() const { long n = 0; for (const auto &s : m_seqInfo) n += s.length; return n; }

    // builds a fraction from already-packed arrays (used by the C API for externally packed data)
    static DataSetFraction *fromPacked(bl_ctx *ctx, int S, int T, int Tmin, int numSeqs, const int *seqLengths, int P, int O,
                                       const real_t *inputs, const char *patTypes, const int *targetClasses, const real_t *targets);

private:
    int m_inputPatternSize;
    int m_outputPatternSize;
    int m_maxSeqLength;
    int m_minSeqLength;
    int m_parallelSequences;
    std::vector<seq_info_t> m_seqInfo;
    HostBuffer<real_t> m_inputs;
    HostBuffer<real_t> m_outputs;
    HostBuffer<char>   m_patTypes;
    HostBuffer<int>    m_targetClasses;
    mutable bl_ctx *m_uploadCtx = nullptr;
    mutable unsigned long long m_uploadTicket = 0;
};

} // namespace data_sets
