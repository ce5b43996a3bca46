/*
 * ctcasr.h — C-ABI of libctcasr.so, the B200 (sm_100a) CTC training core.
 *
 * The reference (mdangschat/ctc-asr) has no native layer: its hot path is two Python functions
 * that call TensorFlow library ops.  Each entry point below replaces one of those call sites
 * (cited as asr/<file>:<line> of the reference) and is what a ctypes/cffi binding on the
 * reference side would bind (see INTEGRATION.md for that stub).
 *
 * Conventions
 *   - plain C, no torch types; all tensor pointers are DEVICE pointers owned by the caller,
 *     fp32 row-major unless noted; activations are time-major ([T, B, C], row = t*B + b).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates
 *     nothing, and keeps no global mutable state apart from the last-error string and lazily
 *     created tensor-map / attribute caches.
 *   - return value: 0 = CTCASR_OK, negative = error (ctcasr_last_error() has the text).
 *     Device-detected per-utterance conditions are reported through `status` arrays.
 *   - `ws` is caller-provided scratch of at least *_workspace_bytes(...) bytes, 256-B aligned.
 */
#ifndef CTCASR_H
#define CTCASR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCASR_ABI_VERSION 1

enum {
    CTCASR_OK = 0,
    CTCASR_ERR_INVALID = -1,      /* bad argument (shape, alignment, null pointer) */
    CTCASR_ERR_UNSUPPORTED = -2,  /* legal but not implemented for this shape/cell */
    CTCASR_ERR_WORKSPACE = -3,    /* ws_bytes too small */
    CTCASR_ERR_CUDA = -4          /* a CUDA runtime/driver call failed */
};

/* rnn cell menu = asr/params.py:48-50 / asr/model.py:194-199 */
enum { CTCASR_CELL_RNN_TANH = 0, CTCASR_CELL_RNN_RELU = 1, CTCASR_CELL_LSTM = 2, CTCASR_CELL_GRU = 3 };

/* arithmetic of the GEMM-shaped work */
enum {
    CTCASR_COMPUTE_FP32 = 0,      /* SIMT FFMA, fp32 everywhere (parity mode, any shape) */
    CTCASR_COMPUTE_TF32 = 1,      /* tcgen05 kind::tf32, fp32 storage + fp32 accumulate in TMEM */
    CTCASR_COMPUTE_BF16X3 = 2,    /* tcgen05 kind::f16 on bf16-split operands (a = a1 + a2 [+ a3]),
                                     3 or 6 products accumulated in fp32 TMEM (chains of <= 8192, gemm_tc.cu): 1e-5 .. 3e-5 of fp64 */
    CTCASR_COMPUTE_BF16 = 3       /* GEMMs and recurrence: operands rounded to bf16, one tcgen05 kind::f16 product, fp32
                                     accumulate (fp32 master weights, state and CTC; the 29-class logits layer stays bf16x3) */
};

/* per-utterance CTC status words (TF raises InvalidArgumentError for 1..3) */
enum { CTCASR_CTC_OK = 0, CTCASR_CTC_INFEASIBLE = 1, CTCASR_CTC_BAD_LABEL = 2, CTCASR_CTC_BAD_LENGTH = 3 };

int ctcasr_abi_version(void);
const char *ctcasr_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t ctcasr_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * CTC loss forward-backward.  Replaces tf.nn.ctc_loss at asr/model.py:259-264
 * (ctc_merge_repeated=True, preprocess_collapse_repeated=False, time_major=True) plus the
 * 1/B of tf.reduce_mean (asr/model.py:267) through `grad_scale`.
 *   logits        [T,B,V]   labels [B,label_stride] int32 (0-padded like asr/model.py:71 strips)
 *   label_len[B], seq_len[B] int32;   blank = V-1 in the reference (asr/labels.py:6)
 *   loss[B]       -log p(labels|logits) per utterance (+inf when status != 0)
 *   grad          [T,B,V] or NULL:  grad_scale * d loss_b / d logits, zero for t >= seq_len[b]
 *   status[B]     CTCASR_CTC_*
 * -------------------------------------------------------------------------------------------- */
size_t ctcasr_ctc_workspace_bytes(int T, int B, int V, int max_label_len);
int ctcasr_ctc_loss(const float *logits, int T, int B, int V, int blank,
                    const int32_t *labels, int label_stride, const int32_t *label_len,
                    const int32_t *seq_len, float *loss, float *grad, float grad_scale,
                    int32_t *status, int max_label_len, void *ws, size_t ws_bytes, void *stream);

/* Same call with HOST buffers (pageable or pinned): copies in, runs, copies out, synchronises.
 * This is the drop-in for the reference's CPU-only CTCLoss kernel, which receives the logits by
 * D2H copy every step. */
int ctcasr_ctc_loss_host(const float *logits, int T, int B, int V, int blank,
                         const int32_t *labels, int label_stride, const int32_t *label_len,
                         const int32_t *seq_len, float *loss, float *grad, float grad_scale,
                         int32_t *status);

/* ---------------------------------------------------------------------------------------------
 * Greedy CTC decode (argmax, merge repeats, drop blank) — the parity-checked stand-in for
 * decode_fn's tf.nn.ctc_beam_search_decoder call at asr/model.py:292-296.
 *   out_ids [B,T] int32 (-1 padded), out_len[B]
 * -------------------------------------------------------------------------------------------- */
int ctcasr_greedy_decode(const float *logits, int T, int B, int V, int blank,
                         const int32_t *seq_len, int32_t *out_ids, int32_t *out_len, void *stream);

/* Levenshtein distance between label sequences — tf.edit_distance(decoded, labels) at
 * asr/model.py:338 (normalize: divided by the truth length; empty truth -> 0 or +inf like TF).
 *   hyp [B,hyp_stride], truth [B,truth_stride] int32 (only the first *_len[b] entries are read) */
int ctcasr_edit_distance(const int32_t *hyp, int hyp_stride, const int32_t *hyp_len,
                         const int32_t *truth, int truth_stride, const int32_t *truth_len,
                         int B, int max_hyp_len, int normalize, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Dense layer.  Replaces tf.layers.dense + tf.minimum + tf.layers.dropout at
 * asr/util/tf_contrib.py:52-58 and asr/model.py:220-226,232.
 *   y[M,N] = dropout( act( x[M,K] w[K,N] + bias[N] ) ),  act: 0 linear, 1 min(relu(.), cutoff)
 *   dropout keep-mask = counter hash of (seed, m*N+n); rate 0 disables it.
 * Backward: dy[M,N] -> dx[M,K] (nullable), dw[K,N], db[N] (overwritten).  `y` is the forward
 * output (the activation/dropout mask is recovered from it).  dy is clobbered (it may become dz).
 * -------------------------------------------------------------------------------------------- */
int ctcasr_dense_fwd(const float *x, const float *w, const float *bias, float *y,
                     int M, int K, int N, int act, float cutoff, float drop_rate, uint32_t seed,
                     int compute, void *stream);
int ctcasr_dense_bwd(const float *x, const float *w, const float *y, float *dy,
                     float *dx, float *dw, float *db, int M, int K, int N,
                     int act, float cutoff, float drop_rate, uint32_t seed,
                     int compute, void *stream);

/* ---------------------------------------------------------------------------------------------
 * One bidirectional recurrent layer.  Replaces one layer of
 * tfc.rnn.stack_bidirectional_dynamic_rnn (asr/model.py:176-183, cells from
 * asr/util/tf_contrib.py:183-189) / tfc.cudnn_rnn.Cudnn* (asr/model.py:194-215).
 *   x [T,B,in]  ->  y [T,B,2H]  (fw | bw)
 *   wx [in, 2*G*H] (fw gate columns | bw gate columns), wh [2][H, G*H], bias [2*G*H]
 *   LSTM gate order i, j, f, o with forget_bias (TF LSTMCell); G = 4 (LSTM), 3 (GRU), 1 (tanh/relu)
 *   GRU: cuDNN formulation (what CudnnGRU computes, asr/model.py:197), gate order r, z, n;
 *        bias / dbias carry 2*H extra entries after the 2*G*H input-side ones: b_rn [2][H], the
 *        recurrent bias of the candidate gate  n = tanh(x Wn + bn + r * (h Rn + b_rn))
 *   use_len != 0: dynamic_rnn(sequence_length) semantics (zero output + frozen state past
 *                 seq_len[b]; backward cell starts at seq_len[b]-1).  0: every frame (cuDNN path).
 *   reserve: >= ctcasr_birnn_reserve_bytes(); written by fwd, consumed AND clobbered by bwd.
 * Backward: dy [T,B,2H] -> dx [T,B,in] (nullable), dwx, dwh, dbias (overwritten).
 * -------------------------------------------------------------------------------------------- */
size_t ctcasr_birnn_reserve_bytes(int T, int B, int in, int H, int cell);
size_t ctcasr_birnn_workspace_bytes(int T, int B, int in, int H, int cell);
/* bytes one recurrence launch of this shape pulls through TMA from L2 / HBM (weight tiles not resident on chip +
 * the state tiles of every step, all CTAs): the numerator of bench.py's L2-streaming bound.  0 for the stepwise path. */
double ctcasr_birnn_stream_bytes(int T, int B, int H, int cell, int compute, int backward);
int ctcasr_birnn_fwd(const float *x, const int32_t *seq_len, const float *wx, const float *wh,
                     const float *bias, float *y, void *reserve,
                     int T, int B, int in, int H, int cell, int use_len, float forget_bias,
                     int compute, void *ws, size_t ws_bytes, void *stream);
int ctcasr_birnn_bwd(const float *x, const int32_t *seq_len, const float *wx, const float *wh,
                     const float *y, void *reserve, const float *dy,
                     float *dx, float *dwx, float *dwh, float *dbias,
                     int T, int B, int in, int H, int cell, int use_len,
                     int compute, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * CTC prefix beam search.  Replaces tf.nn.ctc_beam_search_decoder(inputs=logits, sequence_length,
 * beam_width=FLAGS.beam_width, top_paths=1, merge_repeated=False) at asr/model.py:292-296
 * (beam_width default 1024, asr/params.py:85).  No language-model scorer, like the reference.
 *   logits [T,B,V] time-major; blank must be V-1 (TF's convention, asr/labels.py:6); V <= 32;
 *   1 <= beam_width <= 1024; frames t >= seq_len[b] are not read.
 *   out_ids [B,T] (row b: the best path's labels, then -1), out_len [B],
 *   out_logp [B] (nullable): log-score of the best path (frame scores relative to the frame maximum,
 *   as TF r1.12's Step() computes them).
 * Exactly equal scores at the beam boundary are ordered by candidate index (TF: heap order).
 * -------------------------------------------------------------------------------------------- */
size_t ctcasr_beam_search_workspace_bytes(int T, int B, int V, int beam_width);
int ctcasr_beam_search(const float *logits, int T, int B, int V, int blank, const int32_t *seq_len,
                       int beam_width, int merge_repeated, int32_t *out_ids, int32_t *out_len,
                       float *out_logp, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * 2-D convolution layer of the 'ds2' front-end.  Replaces one iteration of the loop at
 * asr/util/tf_contrib.py:123-134: tf.layers.conv2d(padding='SAME', activation=relu) followed by
 * tf.minimum(., relu_cutoff); the image height is TIME and its width the feature axis
 * (asr/model.py:157-158), and tf.layers.dropout(rate=conv_dropout_rate) (asr/util/tf_contrib.py:135):
 *   y = dropout( act( conv(x, w) + bias ) ), keep-mask = counter hash of (seed, position*N + n) as in
 *   ctcasr_dense_fwd; drop_rate 0 (the reference's default, asr/params.py:89) disables it.
 *   x  [T, B, F, x_pitch]   time-major, the C real channels first in every x_pitch-float pixel
 *   w  [Kp, N]              TF's HWIO kernel [kt, kf, C, filters] flattened to rows
 *                           (it*kf + jf)*C + c and zero-padded to Kp = roundup(kt*kf*C, 8) rows and
 *                           N >= filters columns (N = the pitch of y; pad entries must be zero)
 *   y  [To, B, Fo, N]       To = ceil(T/st), Fo = ceil(F/sf) (ctcasr_conv2d_out_dims)
 *   act: 0 linear, 1 min(relu, cutoff).
 * Backward: dy [To,B,Fo,N] (clobbered: may become dz) -> dx [T,B,F,x_pitch] (nullable; pad channels
 * are written as zeros), dw [Kp,N], db [N] (overwritten).
 * ws: >= ctcasr_conv2d_workspace_bytes() (the patch matrix; rebuilt in the backward pass).
 * -------------------------------------------------------------------------------------------- */
int ctcasr_conv2d_out_dims(int T, int F, int kt, int kf, int st, int sf, int *To, int *Fo);
size_t ctcasr_conv2d_workspace_bytes(int T, int B, int F, int C, int kt, int kf, int st, int sf);
int ctcasr_conv2d_fwd(const float *x, int x_pitch, const float *w, const float *bias, float *y,
                      int T, int B, int F, int C, int kt, int kf, int st, int sf, int N,
                      int act, float cutoff, float drop_rate, uint32_t seed, int compute,
                      void *ws, size_t ws_bytes, void *stream);
int ctcasr_conv2d_bwd(const float *x, int x_pitch, const float *w, const float *y, float *dy,
                      float *dx, float *dw, float *db,
                      int T, int B, int F, int C, int kt, int kf, int st, int sf, int N,
                      int act, float cutoff, float drop_rate, uint32_t seed, int compute,
                      void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Element-wise dropout of a contiguous buffer: y[i] = keep(seed, i) ? x[i] / (1 - rate) : 0, the keep-mask
 * being the counter hash of ctcasr_dense_fwd over the flat index i.  y may be x (in place).  Applied to a
 * gradient with the same (rate, seed) it is the layer's backward pass.  Replaces the dropout on the
 * non-recurrent connections of the RNN stack: tf.nn.rnn_cell.DropoutWrapper(input_keep_prob, output_keep_prob)
 * (asr/util/tf_contrib.py:190-194) and the `dropout` argument of the cuDNN RNNs, applied between layers
 * (asr/model.py:201-206); rate = rnn_dropout_rate (asr/params.py:91).
 * -------------------------------------------------------------------------------------------- */
int ctcasr_dropout(const float *x, float *y, size_t n, float drop_rate, uint32_t seed, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Feature extraction.  Replaces the arithmetic of load_sample (asr/input_functions.py:156-262) from
 * the decoded samples on: python_speech_features' mfcc + delta (:264-294) or logfbank (:297-318) with
 * the reference's parameters (25 ms / 10 ms frames, pre-emphasis 0.97, nfft 1024, `num_features` mel
 * filters from 64 Hz to Nyquist; MFCC: num_features/2 cepstra, lifter 22, c0 := log energy, + deltas
 * over +-2 frames) and __feature_normalization (:321-349).  Reading the WAV file stays on the host.
 *   audio [B, max_samples] int16 (DEVICE), utterance b in its first num_samples_host[b] samples;
 *   num_samples_host [B]: HOST array (the caller knows its file lengths; each must be >= 401, the
 *   reference's own lower bound, :213-214);
 *   feature_type 0 'mel' | 1 'mfcc';  normalization 0 'none' | 1 'local' | 2 'local_scalar';
 *   features [B, Tmax, num_features] float32 (DEVICE), zero past each utterance's frames — the
 *   `sequences` layout of asr/model.py:129;  num_frames [B] (DEVICE).
 * -------------------------------------------------------------------------------------------- */
int ctcasr_feature_frames(int num_samples, int sampling_rate);
int ctcasr_feature_filterbank_bins(int sampling_rate, int num_filters, int32_t *bins_host);
size_t ctcasr_featurize_workspace_bytes(int B, int max_samples, int sampling_rate, int num_features);
int ctcasr_featurize(const int16_t *audio, int B, int max_samples, const int32_t *num_samples_host,
                     int feature_type, int normalization, int drop_every_second_frame,
                     int sampling_rate, int num_features,
                     float *features, int Tmax, int32_t *num_frames, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Plumbing around the path
 * -------------------------------------------------------------------------------------------- */
/* Optional timing of the hot kernels with CUDA events recorded on the launching stream.
 * classes: 0 LSTM recurrence fwd, 1 LSTM recurrence bwd, 2 tcgen05 GEMM, 3 CTC, 4 operand split.
 * collect() synchronises the device and returns summed milliseconds and launch counts per class. */
int ctcasr_profile_enable(int on);
int ctcasr_profile_collect(double *ms, int *count, int ntags);
/* Scratch arena for the bf16-split operand copies of CTCASR_COMPUTE_BF16X3 (device memory owned by the
 * caller, 1024-B aligned).  A call that needs more returns CTCASR_ERR_WORKSPACE and
 * ctcasr_scratch_needed() tells how much. */
int ctcasr_set_scratch(void *ptr, size_t bytes);
size_t ctcasr_scratch_needed(void);
size_t ctcasr_scratch_bytes(void);
/* Contexts (no global mutable state except an opaque handle).  The scratch arena and the cache of
 * split operands that lives in it belong to a context.  Every host thread works on the context it last bound with
 * ctcasr_use() — the process-default context until then, or after ctcasr_use(NULL) — so two host threads that drive two
 * CUDA streams (TensorFlow's inter-op pool running two towers) give each its own handle and its own arena and never
 * share split operands.  Entry points that take a workspace argument use nothing else.  The reference has no
 * counterpart: TF owns this state in its per-stream scratch allocators. */
typedef struct ctcasr_context *ctcasr_handle_t;
int ctcasr_create(ctcasr_handle_t *out);
int ctcasr_use(ctcasr_handle_t handle);
int ctcasr_destroy(ctcasr_handle_t handle);
/* [A,B,C] -> [B,A,C]: batch-major sequences (asr/model.py:129) <-> the time-major layout used
 * internally and by the logits (asr/model.py:233) */
int ctcasr_transpose01(const float *in, float *out, int A, int B, int C, void *stream);
/* Adam, TF1 formulation (tf.train.AdamOptimizer, asr/model.py:79-83) on a flat buffer;
 * grad_scale multiplies g first (e.g. 1/world_size after a sum all-reduce). */
int ctcasr_adam(float *p, float *m, float *v, const float *g, size_t n, int step,
                float lr, float beta1, float beta2, float eps, float grad_scale, void *stream);
/* C[M,N] = op(A) op(B): generic GEMM used by the layers above, exposed for tests.
 *   ta == 0: A is [M,K] row-major, ta != 0: A is [K,M];  tb == 0: B is [K,N], tb != 0: B is [N,K] */
int ctcasr_gemm(const float *A, const float *B, float *C, int M, int N, int K,
                int ta, int tb, int lda, int ldb, int ldc, int accumulate, int compute, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCASR_H */
