// C ABI over the host layer (include/currennt_b200.h).
#include "currennt_b200.h"
#include <cstring>
#include <memory>
#include <string>
#include "DataSet.hpp"
#include "NetCdf.hpp"
#include "NeuralNetwork.hpp"
#include "Optimizer.hpp"

struct cn_net {
    helpers::JsonDocument doc;
    std::unique_ptr<NeuralNetwork> net;
};
struct cn_opt { std::unique_ptr<optimizers::SteepestDescentOptimizer> opt; cn_net *net; };
struct cn_dataset { std::unique_ptr<data_sets::DataSet> ds; };
struct cn_fraction { std::shared_ptr<data_sets::DataSetFraction> frac; };

static thread_local std::string g_err;

#define CN_TRY try {
#define CN_CATCH(ret) } catch (const std::exception &e) { g_err = e.what(); return ret; } catch (...) { g_err = "unknown error"; return ret; }

static layers::TrainableLayer *trainable(const cn_net *n, int i)
{
    if (i < 0 || i >= (int)n->net->layers().size()) throw std::runtime_error("layer index out of range");
    return dynamic_cast<layers::TrainableLayer *>(n->net->layers()[i].get());
}
static layers::Layer *layerAt(const cn_net *n, int i)
{
    if (i < 0 || i >= (int)n->net->layers().size()) throw std::runtime_error("layer index out of range");
    return n->net->layers()[i].get();
}

extern "C" {

const char *cn_last_error(void) { return g_err.c_str(); }

cn_net *cn_net_create(bl_ctx *ctx, const char *json, int S, int maxT)
{
    CN_TRY
    if (!ctx) throw std::runtime_error("cn_net_create: ctx is NULL");
    std::unique_ptr<cn_net> h(new cn_net);
    h->doc = helpers::parseJson(json);
    h->net.reset(new NeuralNetwork(ctx, h->doc, S, maxT));
    return h.release();
    CN_CATCH(nullptr)
}
void cn_net_destroy(cn_net *net) { delete net; }
int cn_net_num_layers(const cn_net *net) { return (int)net->net->layers().size(); }
int cn_layer_size(const cn_net *net, int i) { CN_TRY return layerAt(net, i)->size(); CN_CATCH(-1) }
const char *cn_layer_type(const cn_net *net, int i) { CN_TRY return layerAt(net, i)->type().c_str(); CN_CATCH(nullptr) }
const char *cn_layer_name(const cn_net *net, int i) { CN_TRY return layerAt(net, i)->name().c_str(); CN_CATCH(nullptr) }
long cn_layer_num_weights(const cn_net *net, int i) { CN_TRY layers::TrainableLayer *l = trainable(net, i); return l ? (long)l->weights().size() : 0; CN_CATCH(-1) }

int cn_layer_set_weights(cn_net *net, int i, const float *w, long n)
{
    CN_TRY
    layers::TrainableLayer *l = trainable(net, i);
    if (!l) throw std::runtime_error("layer is not trainable");
    l->setWeights(w, (size_t)n);
    return 0;
    CN_CATCH(1)
}
int cn_layer_get_weights(cn_net *net, int i, float *w, long n)
{
    CN_TRY
    layers::TrainableLayer *l = trainable(net, i);
    if (!l || (long)l->weights().size() != n) throw std::runtime_error("bad layer / weight count");
    if (n) l->weights().toHost(w, (size_t)n);
    return 0;
    CN_CATCH(1)
}
int cn_layer_get_weight_updates(cn_net *net, int i, float *w, long n)
{
    CN_TRY
    layers::TrainableLayer *l = trainable(net, i);
    if (!l || (long)l->weightUpdates().size() != n) throw std::runtime_error("bad layer / weight count");
    if (n) l->weightUpdates().toHost(w, (size_t)n);
    return 0;
    CN_CATCH(1)
}
int cn_layer_get_outputs(cn_net *net, int i, float *dst, long n)
{
    CN_TRY
    std::vector<real_t> v = layerAt(net, i)->outputsToHost();
    if ((long)v.size() != n) throw std::runtime_error("outputs: expected " + std::to_string(v.size()) + " values");
    std::memcpy(dst, v.data(), v.size() * sizeof(real_t));
    return 0;
    CN_CATCH(1)
}
int cn_layer_get_output_errors(cn_net *net, int i, float *dst, long n)
{
    CN_TRY
    std::vector<real_t> v = layerAt(net, i)->outputErrorsToHost();
    if ((long)v.size() != n) throw std::runtime_error("outputErrors: expected " + std::to_string(v.size()) + " values");
    std::memcpy(dst, v.data(), v.size() * sizeof(real_t));
    return 0;
    CN_CATCH(1)
}
int cn_lstm_get_internal(cn_net *net, int i, int dir, int which, float *dst, long n)
{
    CN_TRY
    layers::LstmLayer *l = dynamic_cast<layers::LstmLayer *>(layerAt(net, i));
    if (!l) throw std::runtime_error("not an lstm layer");
    std::vector<real_t> v = l->internalOfDirection(dir, which);
    if ((long)v.size() != n) throw std::runtime_error("internal: expected " + std::to_string(v.size()) + " values");
    std::memcpy(dst, v.data(), v.size() * sizeof(real_t));
    return 0;
    CN_CATCH(1)
}
int cn_lstm_plan_info(cn_net *net, int i, int *out8)
{
    CN_TRY
    layers::LstmLayer *l = dynamic_cast<layers::LstmLayer *>(layerAt(net, i));
    if (!l) throw std::runtime_error("not an lstm layer");
    l->planInfo(out8);
    return 0;
    CN_CATCH(1)
}
int cn_lstm_debug_trace(cn_net *net, int i, int T, long long *dst, int *rows)
{
    CN_TRY
    layers::LstmLayer *l = dynamic_cast<layers::LstmLayer *>(layerAt(net, i));
    if (!l) throw std::runtime_error("not an lstm layer");
    device::check(l->ctx(), bl_lstm_debug_trace(l->plan(), T, dst, rows));
    return 0;
    CN_CATCH(1)
}
int cn_lstm_debug_trace2(cn_net *net, int i, int backward, int T, long long *dst, int *rows)
{
    CN_TRY
    layers::LstmLayer *l = dynamic_cast<layers::LstmLayer *>(layerAt(net, i));
    if (!l) throw std::runtime_error("not an lstm layer");
    device::check(l->ctx(), bl_lstm_debug_trace2(l->plan(), backward, T, dst, rows));
    return 0;
    CN_CATCH(1)
}
long cn_net_export_json(cn_net *net, char *buf, long cap)
{
    CN_TRY
    helpers::JsonDocument doc = helpers::JsonValue::makeObject();
    net->net->exportLayers(doc);
    net->net->exportWeights(doc);
    const std::string s = doc.serialize(true);
    if (buf && cap > 0) { const size_t n = std::min((size_t)cap - 1, s.size()); std::memcpy(buf, s.data(), n); buf[n] = 0; }
    return (long)s.size() + 1;
    CN_CATCH(-1)
}

int cn_net_load_fraction(cn_net *net, const cn_fraction *f) { CN_TRY net->net->loadSequences(*f->frac); return 0; CN_CATCH(1) }
int cn_net_forward(cn_net *net)  { CN_TRY net->net->computeForwardPass();  return 0; CN_CATCH(1) }
int cn_net_backward(cn_net *net) { CN_TRY net->net->computeBackwardPass(); net->net->joinGradients(); return 0; CN_CATCH(1) }
int cn_net_calculate_error(cn_net *net, float *e) { CN_TRY *e = net->net->calculateError(); return 0; CN_CATCH(1) }
int cn_net_count_correct(cn_net *net, int *c)
{
    CN_TRY
    layers::MulticlassClassificationLayer *l = dynamic_cast<layers::MulticlassClassificationLayer *>(&net->net->postOutputLayer());
    if (l) { *c = l->countCorrectClassifications(); return 0; }
    layers::BinaryClassificationLayer *b = dynamic_cast<layers::BinaryClassificationLayer *>(&net->net->postOutputLayer());
    if (!b) throw std::runtime_error("post output layer is not a classification layer");
    *c = b->countCorrectClassifications();
    return 0;
    CN_CATCH(1)
}
int cn_net_set_comm(cn_net *net, bl_comm *comm) { net->net->setCommunicator(comm); return 0; }

cn_fraction *cn_fraction_create(bl_ctx *ctx, int S, int T, int Tmin, int numSeqs, const int *seqLengths, int P, int O,
                                const float *inputs, const char *pat, const int *tc, const float *targets)
{
    CN_TRY
    std::unique_ptr<cn_fraction> f(new cn_fraction);
    f->frac.reset(data_sets::DataSetFraction::fromPacked(ctx, S, T, Tmin, numSeqs, seqLengths, P, O, inputs, pat, tc, targets));
    return f.release();
    CN_CATCH(nullptr)
}
void cn_fraction_destroy(cn_fraction *f) { delete f; }
int cn_fraction_info(const cn_fraction *f, long *o)
{
    const data_sets::DataSetFraction &d = *f->frac;
    o[0] = d.maxSeqLength(); o[1] = d.minSeqLength(); o[2] = d.numSequences(); o[3] = d.parallelSequences();
    o[4] = d.inputPatternSize(); o[5] = d.outputPatternSize(); o[6] = d.validFrames();
    return 0;
}
int cn_fraction_get(const cn_fraction *f, float *inputs, char *pat, int *tc, float *targets, int *seqLengths)
{
    const data_sets::DataSetFraction &d = *f->frac;
    if (inputs && d.inputs().size()) std::memcpy(inputs, d.inputs().data(), d.inputs().size() * sizeof(float));
    if (pat && d.patTypes().size()) std::memcpy(pat, d.patTypes().data(), d.patTypes().size());
    if (tc && d.targetClasses().size()) std::memcpy(tc, d.targetClasses().data(), d.targetClasses().size() * sizeof(int));
    if (targets && d.outputs().size()) std::memcpy(targets, d.outputs().data(), d.outputs().size() * sizeof(float));
    if (seqLengths) for (int i = 0; i < d.numSequences(); ++i) seqLengths[i] = d.seqInfo(i).length;
    return 0;
}

cn_dataset *cn_dataset_create(bl_ctx *ctx, int numSeqs, const int *seqLengths, int P, int O, const float *inputs,
                              const int *tc, const float *targets, int parSeq, int trunc, int training, int rank, int world)
{
    CN_TRY
    std::unique_ptr<cn_dataset> d(new cn_dataset);
    d->ds.reset(new data_sets::DataSet(ctx, numSeqs, seqLengths, P, O, inputs, tc, targets, parSeq, trunc, training != 0, rank, world));
    return d.release();
    CN_CATCH(nullptr)
}
cn_dataset *cn_dataset_load_netcdf(bl_ctx *ctx, const char *path, int parSeq, float fraction, int trunc, int training, int rank, int world)
{
    CN_TRY
    std::unique_ptr<cn_dataset> d(new cn_dataset);
    std::vector<std::string> files;
    std::string all(path);                                   // comma separated list, like the reference's --train_file
    size_t pos = 0;
    while (pos <= all.size()) { size_t c = all.find(',', pos); if (c == std::string::npos) c = all.size(); if (c > pos) files.push_back(all.substr(pos, c - pos)); pos = c + 1; }
    d->ds = data_sets::loadNetCdfDataSet(ctx, files, parSeq, fraction, trunc, training != 0, rank, world);
    return d.release();
    CN_CATCH(nullptr)
}
int cn_dataset_set_context(cn_dataset *ds, int left, int right, int lag) { CN_TRY ds->ds->setContext(left, right, lag); return 0; CN_CATCH(1) }
void cn_dataset_destroy(cn_dataset *ds) { delete ds; }
int cn_dataset_info(const cn_dataset *ds, long *o)
{
    const data_sets::DataSet &d = *ds->ds;
    o[0] = d.totalSequences(); o[1] = d.totalTimesteps(); o[2] = d.minSeqLength(); o[3] = d.maxSeqLength();
    o[4] = d.numFractions(); o[5] = d.isClassificationData();
    return 0;
}
int cn_dataset_sequence_lengths(const cn_dataset *ds, int *out, int cap)
{
    const auto &s = ds->ds->sequences();
    for (int i = 0; i < (int)s.size() && i < cap; ++i) out[i] = s[i].length;
    return (int)s.size();
}
cn_fraction *cn_dataset_next_fraction(cn_dataset *ds)
{
    CN_TRY
    std::shared_ptr<data_sets::DataSetFraction> f = ds->ds->getNextFraction();
    if (!f) { g_err.clear(); return nullptr; }
    cn_fraction *h = new cn_fraction; h->frac = f;
    return h;
    CN_CATCH(nullptr)
}
cn_fraction *cn_dataset_make_fraction(cn_dataset *ds, int first)
{
    CN_TRY
    cn_fraction *h = new cn_fraction; h->frac = ds->ds->makeFraction(first);
    return h;
    CN_CATCH(nullptr)
}

cn_opt *cn_opt_create(cn_net *net, float lr, float momentum, int hybrid)
{
    CN_TRY
    std::unique_ptr<cn_opt> o(new cn_opt);
    o->net = net;
    o->opt.reset(new optimizers::SteepestDescentOptimizer(*net->net, lr, momentum, hybrid != 0));
    return o.release();
    CN_CATCH(nullptr)
}
void cn_opt_destroy(cn_opt *opt) { delete opt; }
int cn_opt_train_fraction(cn_opt *opt, const cn_fraction *f, int first, float *error, int *correct, long *frames)
{
    CN_TRY
    const optimizers::StepResult r = opt->opt->trainFraction(*f->frac, first != 0);
    if (error) *error = r.error;
    if (correct) *correct = r.correct;
    if (frames) *frames = r.frames;
    return 0;
    CN_CATCH(1)
}
int cn_opt_eval_fraction(cn_opt *opt, const cn_fraction *f, float *error, int *correct, long *frames)
{
    CN_TRY
    const optimizers::StepResult r = opt->opt->evalFraction(*f->frac);
    if (error) *error = r.error;
    if (correct) *correct = r.correct;
    if (frames) *frames = r.frames;
    return 0;
    CN_CATCH(1)
}
int cn_opt_update_weights(cn_opt *opt) { CN_TRY opt->opt->updateWeights(); return 0; CN_CATCH(1) }
int cn_opt_process_dataset(cn_opt *opt, cn_dataset *ds, int calc, float *error, float *classError)
{
    CN_TRY
    real_t ce = 0;
    const real_t e = opt->opt->processDataSet(*ds->ds, calc != 0, &ce);
    if (error) *error = e;
    if (classError) *classError = ce;
    return 0;
    CN_CATCH(1)
}
int cn_opt_get_weight_deltas(cn_opt *opt, int layer, float *dst, long n)
{
    CN_TRY
    const auto all = opt->opt->weightDeltasToHost();
    if (layer < 0 || layer >= (int)all.size() || (long)all[layer].size() != n) throw std::runtime_error("bad layer / count");
    if (n) std::memcpy(dst, all[layer].data(), (size_t)n * sizeof(float));
    return 0;
    CN_CATCH(1)
}

} // extern "C"
