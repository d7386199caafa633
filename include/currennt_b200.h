/*
 * currennt_b200.h -- C ABI over the C++ host layer (lstm-rnn_b200/host), i.e. the reference-facing call path
 * NeuralNetwork::{loadSequences, computeForwardPass, calculateError, computeBackwardPass} + the optimizer step
 * (reference: NeuralNetwork.cpp:161-190, optimizers/Optimizer.cu:46-97) with HOST buffers in and out.
 * This is what bench.py's `e2e` number and the `-m gpu` parity tests drive (through ctypes); a C++ caller uses
 * the classes in lstm-rnn_b200/host directly.  All functions return 0 on success (or a handle), never throw;
 * cn_last_error() returns the message of the last failure on the calling thread.
 */
#ifndef CURRENNT_B200_H
#define CURRENNT_B200_H

#include "blstm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cn_net      cn_net;
typedef struct cn_opt      cn_opt;
typedef struct cn_dataset  cn_dataset;
typedef struct cn_fraction cn_fraction;

const char *cn_last_error(void);

/* network (NeuralNetwork.cpp:37-130): JSON text in the reference's schema; weights random unless the JSON carries them */
cn_net *cn_net_create(bl_ctx *ctx, const char *network_json, int parallel_sequences, int max_seq_length);
void    cn_net_destroy(cn_net *net);
int     cn_net_num_layers(const cn_net *net);
int     cn_layer_size(const cn_net *net, int layer);
const char *cn_layer_type(const cn_net *net, int layer);
const char *cn_layer_name(const cn_net *net, int layer);
long    cn_layer_num_weights(const cn_net *net, int layer);
int     cn_layer_set_weights(cn_net *net, int layer, const float *host_w, long n);
int     cn_layer_get_weights(cn_net *net, int layer, float *host_w, long n);
int     cn_layer_get_weight_updates(cn_net *net, int layer, float *host_w, long n);
/* packed host copies [T*S][size] of the layer's outputs() / outputErrors() */
int     cn_layer_get_outputs(cn_net *net, int layer, float *host_dst, long n);
int     cn_layer_get_output_errors(cn_net *net, int layer, float *host_dst, long n);
/* [T*S][H] host copy of an LSTM-internal tensor; `which` as bl_lstm_get_internal */
int     cn_lstm_get_internal(cn_net *net, int layer, int dir, int which, float *host_dst, long n);
int     cn_lstm_plan_info(cn_net *net, int layer, int *out8);
int     cn_lstm_debug_trace(cn_net *net, int layer, int T, long long *host_dst, int *rows);
int     cn_lstm_debug_trace2(cn_net *net, int layer, int backward, int T, long long *host_dst, int *rows);   /* tm2 kernels: 8 stamps per step */
/* serialises {"layers":..., "weights":...} (NeuralNetwork.cpp:192-235); returns the length needed (incl. NUL) */
long    cn_net_export_json(cn_net *net, char *buf, long cap);

int cn_net_load_fraction(cn_net *net, const cn_fraction *frac);
int cn_net_forward(cn_net *net);
int cn_net_backward(cn_net *net);
int cn_net_calculate_error(cn_net *net, float *error);
int cn_net_count_correct(cn_net *net, int *correct);
int cn_net_set_comm(cn_net *net, bl_comm *comm);

/* fractions (data_sets/DataSetFraction.hpp): from already packed host arrays, or produced by a data set */
cn_fraction *cn_fraction_create(bl_ctx *ctx, int S, int T, int Tmin, int num_seqs, const int *seq_lengths, int P, int O,
                                const float *inputs, const char *pat_types, const int *target_classes, const float *targets);
void cn_fraction_destroy(cn_fraction *frac);
/* out7 = {T, Tmin, numSeqs, S, P, O, validFrames} */
int  cn_fraction_info(const cn_fraction *frac, long *out7);
/* copies the packed arrays out (any pointer may be NULL) */
int  cn_fraction_get(const cn_fraction *frac, float *inputs, char *pat_types, int *target_classes, float *targets, int *seq_lengths);

/* in-memory data set with the reference's truncation / sort / packing (data_sets/DataSet.cpp:300-414, 527-542, 603-605) */
cn_dataset *cn_dataset_create(bl_ctx *ctx, int num_seqs, const int *seq_lengths, int P, int O, const float *inputs,
                              const int *target_classes, const float *targets, int parallel_sequences, int truncate_seq,
                              int training_mode, int rank, int world);
/* NetCDF-3 classic data files in the reference's schema (data_sets/DataSet.cpp:443-606); `path` may be a comma separated list */
cn_dataset *cn_dataset_load_netcdf(bl_ctx *ctx, const char *path, int parallel_sequences, float fraction, int truncate_seq,
                                   int training_mode, int rank, int world);
/* --input_left_context / --input_right_context / --output_time_lag (DataSet.cpp:302-305, 348-393); fractions then carry
 * input patterns of (left + right + 1) x inputPattSize values */
int  cn_dataset_set_context(cn_dataset *ds, int left, int right, int output_time_lag);
void cn_dataset_destroy(cn_dataset *ds);
/* out6 = {totalSequences, totalTimesteps, minSeqLength, maxSeqLength, numFractions, isClassification} */
int  cn_dataset_info(const cn_dataset *ds, long *out6);
/* sequence lengths after truncation+sort, capacity cap; returns the count */
int  cn_dataset_sequence_lengths(const cn_dataset *ds, int *out, int cap);
cn_fraction *cn_dataset_next_fraction(cn_dataset *ds);            /* NULL at the end of each epoch */
cn_fraction *cn_dataset_make_fraction(cn_dataset *ds, int first_seq_idx);

/* optimizer (SteepestDescentOptimizer.cu:39-94 + the step body of Optimizer.cu:46-97) */
cn_opt *cn_opt_create(cn_net *net, float learning_rate, float momentum, int hybrid_online_batch);
void    cn_opt_destroy(cn_opt *opt);
int     cn_opt_train_fraction(cn_opt *opt, const cn_fraction *frac, int first_fraction, float *error, int *correct, long *frames);
int     cn_opt_eval_fraction(cn_opt *opt, const cn_fraction *frac, float *error, int *correct, long *frames);
int     cn_opt_update_weights(cn_opt *opt);
int     cn_opt_process_dataset(cn_opt *opt, cn_dataset *ds, int calc_weight_updates, float *error, float *class_error);
int     cn_opt_get_weight_deltas(cn_opt *opt, int layer, float *host_dst, long n);

#ifdef __cplusplus
}
#endif
#endif /* CURRENNT_B200_H */
