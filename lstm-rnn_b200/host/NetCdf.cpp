#include "NetCdf.hpp"
#include <algorithm>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace data_sets {

namespace {
struct Reader {
    std::ifstream f; bool cdf2 = false;
    uint32_t u32() { unsigned char b[4]; f.read((char *)b, 4); if (!f) throw std::runtime_error("truncated NetCDF header"); return (uint32_t)b[0] << 24 | b[1] << 16 | b[2] << 8 | b[3]; }
    uint64_t u64() { uint64_t hi = u32(); return hi << 32 | u32(); }
    std::string name()
    {
        uint32_t n = u32();
        if (n > 4096) throw std::runtime_error("bad NetCDF header: name of " + std::to_string(n) + " bytes");
        std::string s(n, '\0'); f.read(&s[0], n); f.seekg((4 - n % 4) % 4, std::ios::cur);
        if (!f) throw std::runtime_error("truncated NetCDF header");
        return s;
    }
    void skipAttrs()
    {
        uint32_t tag = u32(), n = u32();
        if (tag == 0 && n == 0) return;
        if (tag != 0x0C) throw std::runtime_error("bad NetCDF attribute list");
        static const int tsz[] = {0, 1, 1, 2, 4, 4, 8};
        for (uint32_t i = 0; i < n; ++i) {
            name();
            uint32_t type = u32(), cnt = u32();
            if (type < 1 || type > 6) throw std::runtime_error("bad NetCDF attribute type");
            uint64_t bytes = (uint64_t)cnt * tsz[type];
            f.seekg((bytes + 3) / 4 * 4, std::ios::cur);
        }
    }
};
const int kTypeSize[] = {0, 1, 1, 2, 4, 4, 8};
}

NetCdfFile::NetCdfFile(const std::string &path) : m_path(path), m_numrecs(0), m_recsize(0)
{
    Reader r;
    r.f.open(path, std::ios::binary);
    if (!r.f) throw std::runtime_error("Could not open '" + path + "'");
    char magic[4]; r.f.read(magic, 4);
    if (std::memcmp(magic, "CDF", 3) != 0 || (magic[3] != 1 && magic[3] != 2))
        throw std::runtime_error("'" + path + "' is not a NetCDF-3 classic file (netCDF-4/HDF5 files must be converted with nccopy -k classic)");
    r.cdf2 = (magic[3] == 2);
    m_numrecs = r.u32();
    uint32_t tag = r.u32(), n = r.u32();
    if (!(tag == 0 && n == 0)) {
        if (tag != 0x0A) throw std::runtime_error("bad NetCDF dimension list");
        for (uint32_t i = 0; i < n; ++i) { std::string nm = r.name(); uint64_t len = r.u32(); m_dimList.emplace_back(nm, len); }
    }
    r.skipAttrs();
    tag = r.u32(); n = r.u32();
    if (!(tag == 0 && n == 0)) {
        if (tag != 0x0B) throw std::runtime_error("bad NetCDF variable list");
        for (uint32_t i = 0; i < n; ++i) {
            std::string nm = r.name();
            Var v; uint32_t nd = r.u32();
            if (nd > 1024) throw std::runtime_error("bad NetCDF header: variable '" + nm + "' has " + std::to_string(nd) + " dimensions");
            for (uint32_t d = 0; d < nd; ++d) {
                const uint32_t id = r.u32();
                if (id >= m_dimList.size()) throw std::runtime_error("bad NetCDF header: variable '" + nm + "' uses an undefined dimension");
                v.dimids.push_back((int)id);
            }
            r.skipAttrs();
            v.type = (int)r.u32(); v.vsize = r.u32(); v.begin = r.cdf2 ? r.u64() : r.u32();
            if (v.type < 1 || v.type > 6) throw std::runtime_error("bad NetCDF variable type");
            v.record = !v.dimids.empty() && m_dimList[v.dimids[0]].second == 0;
            v.count = 1;
            for (size_t d = v.record ? 1 : 0; d < v.dimids.size(); ++d) v.count *= m_dimList[v.dimids[d]].second;
            if (v.record) m_recsize += v.vsize;
            m_vars[nm] = v;
        }
    }
    for (auto &d : m_dimList) m_dims[d.first] = d.second ? d.second : m_numrecs;
}

int NetCdfFile::dimension(const std::string &name) const
{
    auto it = m_dims.find(name);
    if (it == m_dims.end()) throw std::runtime_error("Cannot get dimension '" + name + "'");
    return (int)it->second;
}

template <typename T>
std::vector<T> NetCdfFile::readAs(const std::string &name) const
{
    auto it = m_vars.find(name);
    if (it == m_vars.end()) throw std::runtime_error("Cannot read variable '" + name + "'");
    const Var &v = it->second;
    const int ts = kTypeSize[v.type];
    const size_t nrec = v.record ? (size_t)m_numrecs : 1;
    std::vector<T> out; out.reserve(nrec * v.count);
    std::ifstream f(m_path, std::ios::binary);
    std::vector<unsigned char> buf(v.count * ts);
    for (size_t rec = 0; rec < nrec; ++rec) {
        f.seekg((std::streamoff)(v.begin + (v.record ? rec * m_recsize : 0)));
        f.read((char *)buf.data(), (std::streamsize)buf.size());
        if (!f) throw std::runtime_error("Cannot read array '" + name + "': truncated file");
        for (size_t i = 0; i < v.count; ++i) {
            const unsigned char *p = buf.data() + i * ts;
            switch (v.type) {
                case 1: case 2: out.push_back((T)(signed char)p[0]); break;
                case 3: out.push_back((T)(int16_t)(p[0] << 8 | p[1])); break;
                case 4: out.push_back((T)(int32_t)((uint32_t)p[0] << 24 | p[1] << 16 | p[2] << 8 | p[3])); break;
                case 5: { uint32_t u = (uint32_t)p[0] << 24 | p[1] << 16 | p[2] << 8 | p[3]; float x; std::memcpy(&x, &u, 4); out.push_back((T)x); break; }
                case 6: { uint64_t u = 0; for (int b = 0; b < 8; ++b) u = u << 8 | p[b]; double x; std::memcpy(&x, &u, 8); out.push_back((T)x); break; }
            }
        }
    }
    return out;
}

std::vector<int> NetCdfFile::readInts(const std::string &name) const { return readAs<int>(name); }
std::vector<float> NetCdfFile::readFloats(const std::string &name) const { return readAs<float>(name); }
std::vector<char> NetCdfFile::readChars(const std::string &name) const { return readAs<char>(name); }

std::unique_ptr<DataSet> loadNetCdfDataSet(bl_ctx *ctx, const std::vector<std::string> &ncfiles, int parSeq, real_t fraction,
                                           int truncSeqLength, bool trainingMode, int rank, int world)
{
    if (fraction <= 0 || fraction > 1) throw std::runtime_error("Invalid fraction");
    std::vector<int> seqLengths, classes;
    std::vector<float> inputs, targets, means, stdevs;
    std::vector<std::string> tags;
    bool first = true, isClassification = false;
    int P = 0, O = 0;
    for (const std::string &path : ncfiles) {
        NetCdfFile nc(path);
        const bool cls = nc.hasDimension("numLabels");
        const int p = nc.dimension("inputPattSize");
        int o;
        if (cls) { const int numLabels = nc.dimension("numLabels"); o = (numLabels == 2 ? 1 : numLabels); }      // DataSet.cpp:490-493
        else o = nc.dimension("targetPattSize");
        if (first) { isClassification = cls; P = p; O = o; }
        else {
            if (cls != isClassification) throw std::runtime_error("Cannot combine classification with regression NC");
            if (o != O) throw std::runtime_error(cls ? "Number of classes mismatch in NC files" : "Number of targets mismatch in NC files");
            if (p != P) throw std::runtime_error("Number of inputs mismatch in NC files");
        }
        int nSeq = nc.dimension("numSeqs");
        nSeq = std::max((int)((real_t)nSeq * fraction), 1);                                                       // DataSet.cpp:514-516
        const std::vector<int> lens = nc.readInts("seqLengths");
        // the header is not trusted: every variable must hold what the dimensions promise (a truncated or inconsistent file is an error)
        auto need = [&path](const char *what, size_t have, size_t want) {
            if (have < want)
                throw std::runtime_error("Inconsistent NC file '" + path + "': '" + what + "' holds " + std::to_string(have) +
                                         " values, " + std::to_string(want) + " needed");
        };
        if (p <= 0 || o <= 0) throw std::runtime_error("Inconsistent NC file '" + path + "': pattern sizes must be positive");
        need("seqLengths", lens.size(), (size_t)nSeq);
        size_t frames = 0;
        for (int i = 0; i < nSeq; ++i) {
            if (lens[i] <= 0) throw std::runtime_error("Inconsistent NC file '" + path + "': sequence " + std::to_string(i) + " has length " + std::to_string(lens[i]));
            seqLengths.push_back(lens[i]); frames += (size_t)lens[i];
        }
        const int tagLen = nc.dimension("maxSeqTagLength");
        const std::vector<char> tagChars = nc.readChars("seqTags");
        if (tagLen <= 0) throw std::runtime_error("Inconsistent NC file '" + path + "': maxSeqTagLength must be positive");
        need("seqTags", tagChars.size(), (size_t)nSeq * (size_t)tagLen);
        for (int i = 0; i < nSeq; ++i) {                                                                          // DataSet.cpp:525
            const char *t = tagChars.data() + (size_t)i * tagLen;
            size_t n = 0; while (n < (size_t)tagLen && t[n]) ++n;
            tags.emplace_back(t, n);
        }
        if (first && !cls) {                                                                                      // DataSet.cpp:572-582
            if (nc.hasVariable("outputMeans") && nc.hasVariable("outputStdevs")) { means = nc.readFloats("outputMeans"); stdevs = nc.readFloats("outputStdevs"); }
        }
        const std::vector<float> in = nc.readFloats("inputs");
        need("inputs", in.size(), frames * (size_t)P);
        inputs.insert(inputs.end(), in.begin(), in.begin() + frames * P);
        if (cls) {
            const std::vector<int> tc = nc.readInts("targetClasses");
            need("targetClasses", tc.size(), frames);
            const int numLabels = nc.dimension("numLabels");
            for (size_t i = 0; i < frames; ++i)
                if (tc[i] < 0 || tc[i] >= numLabels)
                    throw std::runtime_error("Inconsistent NC file '" + path + "': target class " + std::to_string(tc[i]) + " outside [0, " + std::to_string(numLabels) + ")");
            classes.insert(classes.end(), tc.begin(), tc.begin() + frames);
        } else {
            const std::vector<float> tp = nc.readFloats("targetPatterns");
            need("targetPatterns", tp.size(), frames * (size_t)O);
            targets.insert(targets.end(), tp.begin(), tp.begin() + frames * O);
        }
        first = false;
    }
    std::unique_ptr<DataSet> ds(new DataSet(ctx, (int)seqLengths.size(), seqLengths.data(), P, O, inputs.data(),
                                            isClassification ? classes.data() : nullptr, isClassification ? nullptr : targets.data(),
                                            parSeq, truncSeqLength, trainingMode, rank, world));
    ds->setSequenceTags(tags);
    if ((int)means.size() != O || (int)stdevs.size() != O) { means.assign(O, 0.0f); stdevs.assign(O, 1.0f); }
    ds->setOutputStatistics(means, stdevs);
    return ds;
}

} // namespace data_sets
