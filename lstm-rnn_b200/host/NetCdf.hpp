// Minimal NetCDF-3 "classic" reader (CDF-1 / CDF-2 on-disk format, big-endian) for the reference's data files
// (schema: data_sets/DataSet.cpp:486-583 -- dims numSeqs, numTimesteps, inputPattSize, numLabels | targetPattSize,
// maxSeqTagLength; vars seqTags, seqLengths, inputs, targetClasses | targetPatterns).  libnetcdf is not needed.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "DataSet.hpp"

namespace data_sets {

class NetCdfFile {
public:
    explicit NetCdfFile(const std::string &path);
    bool hasDimension(const std::string &name) const { return m_dims.count(name) != 0; }
    int dimension(const std::string &name) const;
    bool hasVariable(const std::string &name) const { return m_vars.count(name) != 0; }
    std::vector<int>   readInts(const std::string &name) const;
    std::vector<float> readFloats(const std::string &name) const;
    std::vector<char>  readChars(const std::string &name) const;

private:
    struct Var { int type; std::vector<int> dimids; uint64_t vsize, begin; bool record; size_t count; };
    template <typename T> std::vector<T> readAs(const std::string &name) const;
    std::string m_path;
    std::vector<std::pair<std::string, uint64_t>> m_dimList;
    std::map<std::string, uint64_t> m_dims;
    std::map<std::string, Var> m_vars;
    uint64_t m_numrecs, m_recsize;
};

// Builds a DataSet from one or more .nc files exactly like DataSet::DataSet (DataSet.cpp:443-606): `fraction` keeps the
// first max(1, int(numSeqs*fraction)) sequences of each file.
std::unique_ptr<DataSet> loadNetCdfDataSet(bl_ctx *ctx, const std::vector<std::string> &ncfiles, int parSeq, real_t fraction = 1,
                                           int truncSeqLength = 0, bool trainingMode = true, int rank = 0, int world = 1);

} // namespace data_sets
