// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI harness around the UNMODIFIED reference CPU path (NeuralNetwork<Cpu> and the
// layers::*<Cpu> classes of /root/reference/currennt_lib/src), compiled by
// oracle/build_ref.sh into oracle/_ref/libcurrennt_ref.so.  It lets the tests and the
// cpu_baseline leg of bench.py drive the reference's own
//   loadSequences -> computeForwardPass -> calculateError -> computeBackwardPass
// (NeuralNetwork.cpp:161-190) on in-memory fractions and read back every layer tensor.
//
// Two reference translation units cannot be built here (Boost.program_options /
// libnetcdf are absent): Configuration.cpp and data_sets/DataSet.cpp.  This file
// supplies the handful of their symbols the hot-path objects reference:
//   * Configuration: ctor + the accessors used by TrainableLayer.cu:103-125 (random init)
//   * data_sets::DataSet: only as the `friend` of DataSetFraction (DataSetFraction.hpp:40)
//     so that a fraction can be filled from caller-provided arrays.
// Nothing here restates arithmetic; all numerics come from the reference objects.

#include "NeuralNetwork.hpp"
#include "Configuration.hpp"
#include "data_sets/DataSet.hpp"
#include "data_sets/DataSetFraction.hpp"
#include "layers/Layer.hpp"
#include "layers/TrainableLayer.hpp"
#include "layers/LstmLayer.hpp"
#include "layers/PostOutputLayer.hpp"
#include "layers/MulticlassClassificationLayer.hpp"
#include "layers/BinaryClassificationLayer.hpp"
#include "helpers/JsonClasses.hpp"
#include "rapidjson/document.h"

#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>

// ---------------------------------------------------------------- Configuration stub
Configuration *Configuration::ms_instance = NULL;

Configuration::Configuration(int, const char *[])
{
    m_trainingMode        = true;
    m_hybridOnlineBatch   = true;
    m_useCuda             = false;
    m_randomSeed          = 0;
    m_weightsDistribution = DISTRIBUTION_UNIFORM;
    m_weightsUniformMin   = -0.1f;
    m_weightsUniformMax   = +0.1f;
    m_weightsNormalSigma  = 0.1f;
    m_weightsNormalMean   = 0.0f;
    m_inputNoiseSigma     = 0;
    m_weightNoiseSigma    = 0;
    m_inputLeftContext    = 0;
    m_inputRightContext   = 0;
    m_outputTimeLag       = 0;
    ms_instance = this;
}
Configuration::~Configuration() {}
const Configuration& Configuration::instance()
{
    if (!ms_instance) {
        static const char *argv[] = { "cref" };
        new Configuration(1, argv);
    }
    return *ms_instance;
}
unsigned Configuration::randomSeed() const { return m_randomSeed; }
Configuration::distribution_type_t Configuration::weightsDistributionType() const { return m_weightsDistribution; }
real_t Configuration::weightsDistributionUniformMin()  const { return m_weightsUniformMin; }
real_t Configuration::weightsDistributionUniformMax()  const { return m_weightsUniformMax; }
real_t Configuration::weightsDistributionNormalSigma() const { return m_weightsNormalSigma; }
real_t Configuration::weightsDistributionNormalMean()  const { return m_weightsNormalMean; }

// ---------------------------------------------------------------- DataSet stub (friend of DataSetFraction)
namespace data_sets {
    struct thread_data_t {};

    struct fraction_source_t {
        int P, O, S, T, Tmin;
        const int   *seqLengths; int numSeqs;
        const float *inputs; const char *patTypes; const int *targetClasses; const float *targets;
    };
    static fraction_source_t g_src;

    DataSet::DataSet() : m_curFirstSeqIdx(-1) {}
    DataSet::~DataSet() {}

    // Fills a DataSetFraction verbatim from the caller's already-packed arrays.
    boost::shared_ptr<DataSetFraction> DataSet::getNextFraction()
    {
        boost::shared_ptr<DataSetFraction> frac(new DataSetFraction);
        const fraction_source_t &s = g_src;
        frac->m_inputPatternSize  = s.P;
        frac->m_outputPatternSize = s.O;
        frac->m_maxSeqLength      = s.T;
        frac->m_minSeqLength      = s.Tmin;
        for (int i = 0; i < s.numSeqs; ++i) {
            DataSetFraction::seq_info_t si;
            si.originalSeqIdx = i;
            si.length         = s.seqLengths ? s.seqLengths[i] : s.T;
            si.seqTag         = "seq";
            frac->m_seqInfo.push_back(si);
        }
        size_t n = (size_t)s.T * s.S;
        frac->m_inputs.assign(s.inputs, s.inputs + n * s.P);
        frac->m_patTypes.assign(s.patTypes, s.patTypes + n);
        if (s.targetClasses) frac->m_targetClasses.assign(s.targetClasses, s.targetClasses + n);
        if (s.targets)       frac->m_outputs.assign(s.targets, s.targets + n * s.O);
        return frac;
    }
}

// ---------------------------------------------------------------- C ABI
namespace {
    struct RefNet {
        rapidjson::Document doc;
        NeuralNetwork<Cpu> *net;
        data_sets::DataSet  ds;
        boost::shared_ptr<data_sets::DataSetFraction> frac;
        int S, maxT;
        RefNet() : net(NULL) {}
        ~RefNet() { delete net; }
    };
    thread_local std::string g_err;

    layers::TrainableLayer<Cpu>* trainable(RefNet *h, int i) {
        return dynamic_cast<layers::TrainableLayer<Cpu>*>(h->net->layers()[i].get());
    }
}

#define CREF_TRY   try {
#define CREF_CATCH(ret) } catch (const std::exception &e) { g_err = e.what(); return ret; }

extern "C" {

const char* cref_last_error() { return g_err.c_str(); }

void* cref_net_create(const char *json, int parallelSequences, int maxSeqLength)
{
    CREF_TRY
    Configuration::instance();
    RefNet *h = new RefNet;
    h->S = parallelSequences; h->maxT = maxSeqLength;
    if (h->doc.Parse<0>(json).HasParseError()) { std::string e = h->doc.GetParseError(); delete h; throw std::runtime_error("JSON parse error: " + e); }
    try { h->net = new NeuralNetwork<Cpu>(h->doc, parallelSequences, maxSeqLength); }
    catch (...) { delete h; throw; }
    return h;
    CREF_CATCH(NULL)
}

void cref_net_destroy(void *p) { delete (RefNet*)p; }

int cref_net_num_layers(void *p) { return (int)((RefNet*)p)->net->layers().size(); }

int cref_layer_size(void *p, int i) { return ((RefNet*)p)->net->layers()[i]->size(); }

const char* cref_layer_type(void *p, int i) { return ((RefNet*)p)->net->layers()[i]->type().c_str(); }

long cref_layer_num_weights(void *p, int i)
{
    layers::TrainableLayer<Cpu> *l = trainable((RefNet*)p, i);
    return l ? (long)l->weights().size() : 0;
}

int cref_layer_set_weights(void *p, int i, const float *w, long n)
{
    layers::TrainableLayer<Cpu> *l = trainable((RefNet*)p, i);
    if (!l || (long)l->weights().size() != n) { g_err = "bad layer / weight count"; return 1; }
    std::copy(w, w + n, l->weights().begin());
    return 0;
}

int cref_layer_get_weights(void *p, int i, float *w, long n)
{
    layers::TrainableLayer<Cpu> *l = trainable((RefNet*)p, i);
    if (!l || (long)l->weights().size() != n) { g_err = "bad layer / weight count"; return 1; }
    std::copy(l->weights().begin(), l->weights().end(), w);
    return 0;
}

int cref_layer_get_weight_updates(void *p, int i, float *w, long n)
{
    layers::TrainableLayer<Cpu> *l = trainable((RefNet*)p, i);
    if (!l || (long)l->weightUpdates().size() != n) { g_err = "bad layer / weight count"; return 1; }
    std::copy(l->weightUpdates().begin(), l->weightUpdates().end(), w);
    return 0;
}

// inputs [T][S][P], patTypes [T][S], targetClasses [T][S] (or NULL), targets [T][S][O] (or NULL)
int cref_net_load_fraction(void *p, int T, int Tmin, int numSeqs, const int *seqLengths,
                           int P, int O, const float *inputs, const char *patTypes,
                           const int *targetClasses, const float *targets)
{
    CREF_TRY
    RefNet *h = (RefNet*)p;
    data_sets::fraction_source_t &s = data_sets::g_src;
    s.P = P; s.O = O; s.S = h->S; s.T = T; s.Tmin = Tmin; s.numSeqs = numSeqs; s.seqLengths = seqLengths;
    s.inputs = inputs; s.patTypes = patTypes; s.targetClasses = targetClasses; s.targets = targets;
    h->frac = h->ds.getNextFraction();
    h->net->loadSequences(*h->frac);
    return 0;
    CREF_CATCH(1)
}

int cref_net_forward(void *p)  { CREF_TRY ((RefNet*)p)->net->computeForwardPass();  return 0; CREF_CATCH(1) }
int cref_net_backward(void *p) { CREF_TRY ((RefNet*)p)->net->computeBackwardPass(); return 0; CREF_CATCH(1) }

int cref_net_calculate_error(void *p, float *err)
{
    CREF_TRY *err = ((RefNet*)p)->net->calculateError(); return 0; CREF_CATCH(1)
}

int cref_net_count_correct(void *p, int *n)
{
    CREF_TRY
    layers::MulticlassClassificationLayer<Cpu> *l =
        dynamic_cast<layers::MulticlassClassificationLayer<Cpu>*>(&((RefNet*)p)->net->postOutputLayer());
    if (l) { *n = l->countCorrectClassifications(); return 0; }
    layers::BinaryClassificationLayer<Cpu> *b =
        dynamic_cast<layers::BinaryClassificationLayer<Cpu>*>(&((RefNet*)p)->net->postOutputLayer());
    if (!b) throw std::runtime_error("post output layer is not a classification layer");
    *n = b->countCorrectClassifications();
    return 0;
    CREF_CATCH(1)
}

// copies the first n values of a layer's outputs() / outputErrors()
int cref_layer_get_outputs(void *p, int i, float *dst, long n)
{
    CREF_TRY
    Cpu::real_vector &v = ((RefNet*)p)->net->layers()[i]->outputs();
    if ((long)v.size() < n) throw std::runtime_error("outputs() smaller than requested");
    std::copy(v.begin(), v.begin() + n, dst);
    return 0;
    CREF_CATCH(1)
}

int cref_layer_get_output_errors(void *p, int i, float *dst, long n)
{
    CREF_TRY
    Cpu::real_vector &v = ((RefNet*)p)->net->layers()[i]->outputErrors();
    if ((long)v.size() < n) throw std::runtime_error("outputErrors() smaller than requested");
    std::copy(v.begin(), v.begin() + n, dst);
    return 0;
    CREF_CATCH(1)
}

int cref_layer_set_output_errors(void *p, int i, const float *src, long n)
{
    CREF_TRY
    Cpu::real_vector &v = ((RefNet*)p)->net->layers()[i]->outputErrors();
    if ((long)v.size() < n) throw std::runtime_error("outputErrors() smaller than requested");
    std::copy(src, src + n, v.begin());
    return 0;
    CREF_CATCH(1)
}

// internals of a unidirectional lstm layer (LstmLayer.cu:646-734); which: 0 cellStates, 1 cellStateErrors,
// 2 niActs, 3 igActs, 4 fgActs, 5 ogActs, 6 niDeltas, 7 igDeltas, 8 fgDeltas, 9 ogDeltas
int cref_lstm_get_internal(void *p, int i, int which, float *dst, long n)
{
    CREF_TRY
    layers::LstmLayer<Cpu> *l = dynamic_cast<layers::LstmLayer<Cpu>*>(((RefNet*)p)->net->layers()[i].get());
    if (!l) throw std::runtime_error("not an lstm layer");
    const Cpu::real_vector *v = NULL;
    switch (which) {
        case 0: v = &l->cellStates(); break;      case 1: v = &l->cellStateErrors(); break;
        case 2: v = &l->netInputActs(); break;    case 3: v = &l->inputGateActs(); break;
        case 4: v = &l->forgetGateActs(); break;  case 5: v = &l->outputGateActs(); break;
        case 6: v = &l->netInputDeltas(); break;  case 7: v = &l->inputGateDeltas(); break;
        case 8: v = &l->forgetGateDeltas(); break; case 9: v = &l->outputGateDeltas(); break;
        default: throw std::runtime_error("bad selector");
    }
    if ((long)v->size() < n) throw std::runtime_error("vector smaller than requested");
    std::copy(v->begin(), v->begin() + n, dst);
    return 0;
    CREF_CATCH(1)
}

} // extern "C"
