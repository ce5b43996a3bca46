/*
 * oracle_impl.h — TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * Scalar CPU restatement of the hot path of mdangschat/ctc-asr, included twice by
 * oracle.c (REAL = float, REAL = double).  "parity unpinned": the reference ships no
 * tests or golden vectors for this path and its arithmetic lives in
 * the un-vendored dependency tensorflow>=1.12 (requirements.txt:1-3), which cannot be
 * installed offline.  The restatement therefore follows TensorFlow r1.12's published
 * algorithm and is pinned on external known-answer vectors (TF ctc_loss_op_test
 * `testBasic`, tests/golden/ctc_tf_known_answer.json), brute-force path enumeration and
 * torch-CPU cross-checks (tests/test_oracle_*.py).
 *
 * Reference call sites each function stands in for are cited per function as
 * asr/<file>:<line> (relative to the reference checkout).
 *
 * Layout conventions (all row-major, time-major activations like the reference's logits,
 * asr/model.py:233-235):
 *   logits / grad : [T, B, V]
 *   rnn in/out    : [T, B, *]
 *   weights       : [in, out]  (tf.layers.dense / rnn_cell kernel orientation)
 */

#ifndef REAL
#error "include from oracle.c"
#endif

#define NEG_INF (-(REAL)INFINITY)

/* Operand rounding (oracle_set_operand_rounding(1)): every matrix product below rounds BOTH operands to bfloat16
 * (round-to-nearest-even) before multiplying and accumulates in REAL — the arithmetic of the product path's
 * compute = 'bf16' mode (BASELINE cfg3: bf16 tensor-core operands, fp32 accumulation, fp32 everything else), so
 * that mode can be compared at the same 1e-3 bar as the exact modes.  Off (0) by default: exact products. */
extern int oracle_operand_rounding;
static inline REAL SUF(rb)(REAL v)
{
    if (!oracle_operand_rounding) return v;
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return v;            /* inf / nan */
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    memcpy(&f, &u, 4);
    return (REAL)f;
}
#define RB(v) SUF(rb)(v)

static inline REAL SUF(lse2)(REAL a, REAL b)
{
    /* ctc_loss_util.h LogSumExp: log-zero is -inf */
    if (a == NEG_INF) return b;
    if (b == NEG_INF) return a;
    return a > b ? a + (REAL)log1p(exp((double)(b - a))) : b + (REAL)log1p(exp((double)(a - b)));
}

/* ------------------------------------------------------------------------------------------
 * CTC loss, forward-backward.  Stands in for tf.nn.ctc_loss as called at asr/model.py:259-264
 * (preprocess_collapse_repeated=False, ctc_merge_repeated=True, time_major=True,
 * ignore_longer_outputs_than_inputs=False).  blank = V-1 (asr/labels.py:6, asr/params.py:100).
 *
 * labels: [B, lstride] int32, label_len[B], seq_len[B].
 * loss[B]; grad[T,B,V] = d loss_b / d logits (NOT yet multiplied by the upstream 1/B of
 * tf.reduce_mean, asr/model.py:267); status[B]: 0 ok, 1 infeasible alignment, 2 bad label,
 * 3 seq_len > T.  Infeasible/bad rows get loss=+inf, grad=0 (the TF op raises instead).
 * ------------------------------------------------------------------------------------------ */
int SUF(oracle_ctc_loss)(const REAL *logits, int T, int B, int V, int blank,
                         const int *labels, int lstride, const int *label_len,
                         const int *seq_len, REAL *loss, REAL *grad, int *status)
{
    if (grad) memset(grad, 0, sizeof(REAL) * (size_t)T * B * V);
    for (int b = 0; b < B; ++b) {
        const int Tb = seq_len[b], L = label_len[b], S = 2 * L + 1;
        const int *lab = labels + (size_t)b * lstride;
        status[b] = 0;
        loss[b] = 0;
        if (Tb > T || Tb < 0) { status[b] = 3; loss[b] = (REAL)INFINITY; continue; }
        int repeats = 0, bad = 0;
        for (int i = 0; i < L; ++i) {
            if (lab[i] < 0 || lab[i] >= V || lab[i] == blank) bad = 1;
            if (i > 0 && lab[i] == lab[i - 1]) ++repeats;
        }
        if (bad) { status[b] = 2; loss[b] = (REAL)INFINITY; continue; }
        if (Tb == 0) { if (L > 0) { status[b] = 1; loss[b] = (REAL)INFINITY; } continue; }
        if (Tb < L + repeats) { status[b] = 1; loss[b] = (REAL)INFINITY; continue; }

        int *lp = (int *)malloc(sizeof(int) * S);
        for (int s = 0; s < S; ++s) lp[s] = (s & 1) ? lab[s >> 1] : blank;
        REAL *y = (REAL *)malloc(sizeof(REAL) * (size_t)Tb * V);    /* softmax, prob space */
        REAL *ly = (REAL *)malloc(sizeof(REAL) * (size_t)Tb * V);   /* log y */
        REAL *al = (REAL *)malloc(sizeof(REAL) * (size_t)Tb * S);
        REAL *be = (REAL *)malloc(sizeof(REAL) * (size_t)Tb * S);
        for (int t = 0; t < Tb; ++t) {
            const REAL *x = logits + ((size_t)t * B + b) * V;
            REAL m = x[0];
            for (int k = 1; k < V; ++k) m = x[k] > m ? x[k] : m;
            REAL sum = 0;
            for (int k = 0; k < V; ++k) { y[t * V + k] = (REAL)exp((double)(x[k] - m)); sum += y[t * V + k]; }
            for (int k = 0; k < V; ++k) { y[t * V + k] /= sum; ly[t * V + k] = (REAL)log((double)y[t * V + k]); }
        }
        for (size_t i = 0; i < (size_t)Tb * S; ++i) { al[i] = NEG_INF; be[i] = NEG_INF; }
        /* forward variables */
        al[0 * S + 0] = ly[0 * V + blank];
        if (S > 1) al[0 * S + 1] = ly[0 * V + lp[1]];
        for (int t = 1; t < Tb; ++t) {
            int lo = S - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > S) hi = S;
            for (int s = lo; s < hi; ++s) {
                REAL sum = al[(t - 1) * S + s];
                if (s > 0) sum = SUF(lse2)(sum, al[(t - 1) * S + s - 1]);
                if (s > 1 && lp[s] != blank && lp[s] != lp[s - 2]) sum = SUF(lse2)(sum, al[(t - 1) * S + s - 2]);
                if (sum != NEG_INF) al[t * S + s] = ly[t * V + lp[s]] + sum;
            }
        }
        REAL logp = al[(Tb - 1) * S + S - 1];
        if (S > 1) logp = SUF(lse2)(logp, al[(Tb - 1) * S + S - 2]);
        /* backward variables (TF convention: beta excludes y at its own t) */
        be[(Tb - 1) * S + S - 1] = 0;
        if (S > 1) be[(Tb - 1) * S + S - 2] = 0;
        for (int t = Tb - 2; t >= 0; --t) {
            int lo = S - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > S) hi = S;
            for (int s = lo; s < hi; ++s) {
                REAL sum = NEG_INF;
                if (be[(t + 1) * S + s] != NEG_INF) sum = be[(t + 1) * S + s] + ly[(t + 1) * V + lp[s]];
                if (s < S - 1 && be[(t + 1) * S + s + 1] != NEG_INF)
                    sum = SUF(lse2)(sum, be[(t + 1) * S + s + 1] + ly[(t + 1) * V + lp[s + 1]]);
                if (s < S - 2 && lp[s + 2] != blank && lp[s + 2] != lp[s] && be[(t + 1) * S + s + 2] != NEG_INF)
                    sum = SUF(lse2)(sum, be[(t + 1) * S + s + 2] + ly[(t + 1) * V + lp[s + 2]]);
                be[t * S + s] = sum;
            }
        }
        loss[b] = -logp;
        if (grad) {
            REAL *ps = (REAL *)malloc(sizeof(REAL) * V);
            for (int t = 0; t < Tb; ++t) {
                for (int k = 0; k < V; ++k) ps[k] = NEG_INF;
                for (int s = 0; s < S; ++s) {
                    REAL a = al[t * S + s], c = be[t * S + s];
                    if (a != NEG_INF && c != NEG_INF) ps[lp[s]] = SUF(lse2)(ps[lp[s]], a + c);
                }
                REAL *g = grad + ((size_t)t * B + b) * V;
                for (int k = 0; k < V; ++k) {
                    REAL post = (ps[k] == NEG_INF || logp == NEG_INF) ? (REAL)0 : (REAL)exp((double)(ps[k] - logp));
                    g[k] = y[t * V + k] - post;
                }
            }
            free(ps);
        }
        free(lp); free(y); free(ly); free(al); free(be);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Greedy decode.  Parity-checked decoder named by BASELINE.json's north_star; restates
 * tf.nn.ctc_greedy_decoder(merge_repeated=True) — argmax (first max wins), merge repeats,
 * drop blank.  The reference's decode_fn (asr/model.py:271-309) calls beam search width 1024;
 * see oracle_beam_search below for that row.  out_ids [B, T] (padded with -1), out_len [B].
 * ------------------------------------------------------------------------------------------ */
int SUF(oracle_greedy_decode)(const REAL *logits, int T, int B, int V, int blank,
                              const int *seq_len, int *out_ids, int *out_len)
{
    for (int b = 0; b < B; ++b) {
        int n = 0, prev = -1;
        for (int t = 0; t < T; ++t) out_ids[(size_t)b * T + t] = -1;
        int Tb = seq_len[b] < T ? seq_len[b] : T;
        for (int t = 0; t < Tb; ++t) {
            const REAL *x = logits + ((size_t)t * B + b) * V;
            int am = 0;
            for (int k = 1; k < V; ++k) if (x[k] > x[am]) am = k;
            if (am != blank && am != prev) out_ids[(size_t)b * T + n++] = am;
            prev = am;
        }
        out_len[b] = n;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Dense layer.  tf.layers.dense + tf.minimum(relu, cutoff) + tf.layers.dropout as composed at
 * asr/util/tf_contrib.py:52-58 and asr/model.py:220-226,232.
 *   act: 0 = linear, 1 = min(relu(z), cutoff)
 *   dropout: keep mask from the counter hash shared with the CUDA path (drop_keep below);
 *            rate==0 -> identity (parity runs).  Inverted scaling 1/(1-rate).
 * ------------------------------------------------------------------------------------------ */
static inline uint32_t oracle_hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
static inline int oracle_drop_keep(uint32_t seed, uint64_t idx, float rate)
{
    uint32_t h = oracle_hash32((uint32_t)idx ^ oracle_hash32(seed ^ (uint32_t)(idx >> 32) * 0x9e3779b9U));
    /* 24-bit uniform in [0,1) */
    return ((float)(h >> 8) * (1.0f / 16777216.0f)) >= rate;
}

int SUF(oracle_dense_fwd)(const REAL *x, const REAL *w, const REAL *bias, REAL *y,
                          int M, int K, int N, int act, REAL cutoff, float drop_rate, uint32_t seed)
{
    const REAL inv_keep = drop_rate > 0 ? (REAL)(1.0 / (1.0 - (double)drop_rate)) : (REAL)1;
#pragma omp parallel for
    for (int m = 0; m < M; ++m) {
        for (int n = 0; n < N; ++n) {
            REAL acc = bias ? bias[n] : 0;
            for (int k = 0; k < K; ++k) acc += RB(x[(size_t)m * K + k]) * RB(w[(size_t)k * N + n]);
            if (act == 1) { acc = acc > 0 ? acc : 0; acc = acc < cutoff ? acc : cutoff; }
            if (drop_rate > 0) acc = oracle_drop_keep(seed, (uint64_t)m * N + n, drop_rate) ? acc * inv_keep : 0;
            y[(size_t)m * N + n] = acc;
        }
    }
    return 0;
}

/* Backward of the layer above given its OUTPUT y (mask is recoverable from y: 0<y/inv_keep<cutoff)
 * dy [M,N] -> dx [M,K] (nullable), dw [K,N], db [N].  dw/db are overwritten. */
int SUF(oracle_dense_bwd)(const REAL *x, const REAL *w, const REAL *y, const REAL *dy,
                          REAL *dx, REAL *dw, REAL *db,
                          int M, int K, int N, int act, REAL cutoff, float drop_rate, uint32_t seed)
{
    const REAL inv_keep = drop_rate > 0 ? (REAL)(1.0 / (1.0 - (double)drop_rate)) : (REAL)1;
    REAL *dz = (REAL *)malloc(sizeof(REAL) * (size_t)M * N);
#pragma omp parallel for
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            /* forward: a = min(relu(z), c); y = keep ? a / (1-rate) : 0 */
            REAL g = dy[(size_t)m * N + n];
            const int keep = drop_rate > 0 ? oracle_drop_keep(seed, (uint64_t)m * N + n, drop_rate) : 1;
            g = keep ? g * inv_keep : 0;
            if (act == 1 && keep) {
                /* tf.minimum(relu(z), c): gradient passes where 0 < z < c (ties are measure-zero) */
                /* compare in the scaled domain: a clipped unit stored exactly c * inv_keep */
                const REAL v = y[(size_t)m * N + n];
                if (!(v > 0 && v < cutoff * inv_keep)) g = 0;
            }
            dz[(size_t)m * N + n] = g;
        }
    if (db) for (int n = 0; n < N; ++n) { REAL s = 0; for (int m = 0; m < M; ++m) s += dz[(size_t)m * N + n]; db[n] = s; }
    if (dw) {
#pragma omp parallel for
        for (int k = 0; k < K; ++k)
            for (int n = 0; n < N; ++n) {
                REAL s = 0;
                for (int m = 0; m < M; ++m) s += RB(x[(size_t)m * K + k]) * RB(dz[(size_t)m * N + n]);
                dw[(size_t)k * N + n] = s;
            }
    }
    if (dx) {
#pragma omp parallel for
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) {
                REAL s = 0;
                for (int n = 0; n < N; ++n) s += RB(dz[(size_t)m * N + n]) * RB(w[(size_t)k * N + n]);
                dx[(size_t)m * K + k] = s;
            }
    }
    free(dz);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * One bidirectional recurrent layer.  Restates tfc.rnn.stack_bidirectional_dynamic_rnn's
 * per-layer step (asr/model.py:176-183) with the cells of asr/util/tf_contrib.py:183-189 and
 * the cuDNN cell menu of asr/model.py:194-199 (rnn_relu / rnn_tanh / lstm; GRU = cell 3,
 * cuDNN formulation).  Semantics:
 *   cell 0: h' = tanh(x Wx + h Wh + b)         cell 1: h' = relu(...)
 *   cell 2: TF LSTMCell, gate order i, j, f, o; c' = sig(f + forget_bias) c + sig(i) tanh(j);
 *           h' = sig(o) tanh(c')
 *   cell 3: GRU in the cuDNN formulation the reference reaches through CudnnGRU (asr/model.py:197),
 *           gate order r, z, n:  r = sig(x Wr + h Rr + br), z = sig(x Wz + h Rz + bz),
 *           n = tanh(x Wn + bn + r * (h Rn + b_rn)),  h' = (1 - z) n + z h.
 *           bias then has 2*3H input-side entries followed by b_rn [2*H]; `cstate` stores h Rn + b_rn.
 *   use_len != 0: dynamic_rnn(sequence_length): for t >= len_b output row is zero and the state
 *           is carried through; the backward direction starts at t = len_b - 1 with zero state
 *           (reverse_sequence semantics).  use_len == 0: cuDNN-path behaviour, all T frames.
 * Weights (this repo's packing; a sub-block view of TF's fused [in+H, G*H] kernels):
 *   wx [in, 2*G*H]  (fw gate columns | bw gate columns),  wh [2][H, G*H],  bias [2*G*H]
 * x [T,B,in] -> y [T,B,2H] (fw | bw).  reserve: gates [2][T,B,G*H] (post-activation) and
 * cstate [2][T,B,H] (LSTM only) for the backward pass.
 * ------------------------------------------------------------------------------------------ */
static inline REAL SUF(sigm)(REAL v) { return (REAL)(1.0 / (1.0 + exp(-(double)v))); }

static int SUF(ngates)(int cell) { return cell == 2 ? 4 : (cell == 3 ? 3 : 1); }

int SUF(oracle_birnn_fwd)(const REAL *x, const int *seq_len, const REAL *wx, const REAL *wh,
                          const REAL *bias, REAL *y, REAL *gates, REAL *cstate,
                          int T, int B, int in, int H, int cell, int use_len, REAL forget_bias)
{
    const int G = SUF(ngates)(cell), GH = G * H;
    memset(y, 0, sizeof(REAL) * (size_t)T * B * 2 * H);
    memset(gates, 0, sizeof(REAL) * (size_t)2 * T * B * GH);
    if (cstate) memset(cstate, 0, sizeof(REAL) * (size_t)2 * T * B * H);
#pragma omp parallel for collapse(2)
    for (int d = 0; d < 2; ++d) {
        for (int b = 0; b < B; ++b) {
            const int len = use_len ? (seq_len[b] < T ? seq_len[b] : T) : T;
            const REAL *whd = wh + (size_t)d * H * GH;
            REAL *h = (REAL *)calloc(H, sizeof(REAL));
            REAL *c = (REAL *)calloc(H, sizeof(REAL));
            REAL *z = (REAL *)malloc(sizeof(REAL) * GH);
            for (int step = 0; step < len; ++step) {
                const int t = d == 0 ? step : len - 1 - step;
                const REAL *xt = x + ((size_t)t * B + b) * in;
                for (int g = 0; g < GH; ++g) z[g] = bias[d * GH + g];
                for (int k = 0; k < in; ++k) {
                    const REAL xv = RB(xt[k]);
                    const REAL *wr = wx + (size_t)k * 2 * GH + d * GH;
                    for (int g = 0; g < GH; ++g) z[g] += xv * RB(wr[g]);
                }
                REAL *gt = gates + (((size_t)d * T + t) * B + b) * GH;
                REAL *yt = y + ((size_t)t * B + b) * 2 * H + d * H;
                if (cell == 3) {
                    /* recurrent part kept apart: the candidate gate multiplies it by r */
                    REAL *rh = (REAL *)calloc(GH, sizeof(REAL));
                    for (int k = 0; k < H; ++k) {
                        const REAL hv = RB(h[k]);
                        const REAL *wr = whd + (size_t)k * GH;
                        for (int g = 0; g < GH; ++g) rh[g] += hv * RB(wr[g]);
                    }
                    REAL *qt = cstate + (((size_t)d * T + t) * B + b) * H;
                    for (int u = 0; u < H; ++u) {
                        const REAL q = rh[2 * H + u] + bias[2 * GH + d * H + u];
                        const REAL gr = SUF(sigm)(z[u] + rh[u]), gz = SUF(sigm)(z[H + u] + rh[H + u]);
                        const REAL gn = (REAL)tanh((double)(z[2 * H + u] + gr * q));
                        gt[u] = gr; gt[H + u] = gz; gt[2 * H + u] = gn; qt[u] = q;
                        h[u] = (1 - gz) * gn + gz * h[u];
                        yt[u] = h[u];
                    }
                    free(rh);
                    continue;
                }
                for (int k = 0; k < H; ++k) {
                    const REAL hv = RB(h[k]);
                    const REAL *wr = whd + (size_t)k * GH;
                    for (int g = 0; g < GH; ++g) z[g] += hv * RB(wr[g]);
                }
                if (cell == 2) {
                    REAL *ct = cstate + (((size_t)d * T + t) * B + b) * H;
                    for (int u = 0; u < H; ++u) {
                        REAL gi = SUF(sigm)(z[u]), gj = (REAL)tanh((double)z[H + u]);
                        REAL gf = SUF(sigm)(z[2 * H + u] + forget_bias), go = SUF(sigm)(z[3 * H + u]);
                        c[u] = gf * c[u] + gi * gj;
                        h[u] = go * (REAL)tanh((double)c[u]);
                        gt[u] = gi; gt[H + u] = gj; gt[2 * H + u] = gf; gt[3 * H + u] = go;
                        ct[u] = c[u]; yt[u] = h[u];
                    }
                } else {
                    for (int u = 0; u < H; ++u) {
                        h[u] = cell == 0 ? (REAL)tanh((double)z[u]) : (z[u] > 0 ? z[u] : 0);
                        gt[u] = h[u]; yt[u] = h[u];
                    }
                }
            }
            free(h); free(c); free(z);
        }
    }
    return 0;
}

/* Backward (BPTT) of the layer above.  dy [T,B,2H] -> dx [T,B,in] (nullable), dwx, dwh, dbias
 * (overwritten).  dgates scratch is allocated internally. */
int SUF(oracle_birnn_bwd)(const REAL *x, const int *seq_len, const REAL *wx, const REAL *wh,
                          const REAL *y, const REAL *gates, const REAL *cstate, const REAL *dy,
                          REAL *dx, REAL *dwx, REAL *dwh, REAL *dbias,
                          int T, int B, int in, int H, int cell, int use_len)
{
    const int G = SUF(ngates)(cell), GH = G * H;
    REAL *dz = (REAL *)calloc((size_t)2 * T * B * GH, sizeof(REAL));   /* [2][T,B,GH]: gradient wrt the input-side pre-activations */
    REAL *dzr = cell == 3 ? (REAL *)calloc((size_t)2 * T * B * GH, sizeof(REAL)) : dz;  /* GRU: wrt h R (n columns scaled by r) */
#pragma omp parallel for collapse(2)
    for (int d = 0; d < 2; ++d) {
        for (int b = 0; b < B; ++b) {
            const int len = use_len ? (seq_len[b] < T ? seq_len[b] : T) : T;
            const REAL *whd = wh + (size_t)d * H * GH;
            REAL *dh = (REAL *)calloc(H, sizeof(REAL));    /* recurrent grad into h_t */
            REAL *dc = (REAL *)calloc(H, sizeof(REAL));
            for (int step = len - 1; step >= 0; --step) {
                const int t = d == 0 ? step : len - 1 - step;
                const int tp = d == 0 ? t - 1 : t + 1;         /* previous processed frame */
                const int has_prev = step > 0;
                const REAL *gt = gates + (((size_t)d * T + t) * B + b) * GH;
                REAL *dzt = dz + (((size_t)d * T + t) * B + b) * GH;
                const REAL *dyt = dy + ((size_t)t * B + b) * 2 * H + d * H;
                if (cell == 3) {
                    const REAL *qt = cstate + (((size_t)d * T + t) * B + b) * H;
                    REAL *dzrt = dzr + (((size_t)d * T + t) * B + b) * GH;
                    REAL *dhd = dc;     /* direct path z * dh into h_{t-1}, added after the matvec below */
                    for (int u = 0; u < H; ++u) {
                        const REAL gr = gt[u], gz = gt[H + u], gn = gt[2 * H + u];
                        const REAL hp = has_prev ? y[((size_t)tp * B + b) * 2 * H + d * H + u] : 0;
                        const REAL dht = dyt[u] + dh[u];
                        const REAL dn_pre = dht * (1 - gz) * (1 - gn * gn);
                        const REAL dz_pre = dht * (hp - gn) * gz * (1 - gz);
                        const REAL dr_pre = dn_pre * qt[u] * gr * (1 - gr);
                        dzt[u] = dr_pre; dzt[H + u] = dz_pre; dzt[2 * H + u] = dn_pre;
                        dzrt[u] = dr_pre; dzrt[H + u] = dz_pre; dzrt[2 * H + u] = dn_pre * gr;
                        dhd[u] = dht * gz;
                    }
                    for (int k = 0; k < H; ++k) {
                        const REAL *wr = whd + (size_t)k * GH;
                        REAL s = 0;
                        for (int g = 0; g < GH; ++g) s += RB(dzrt[g]) * RB(wr[g]);
                        dh[k] = s + dhd[k];
                    }
                    continue;
                }
                if (cell == 2) {
                    const REAL *ct = cstate + (((size_t)d * T + t) * B + b) * H;
                    const REAL *cp = has_prev ? cstate + (((size_t)d * T + tp) * B + b) * H : NULL;
                    for (int u = 0; u < H; ++u) {
                        REAL gi = gt[u], gj = gt[H + u], gf = gt[2 * H + u], go = gt[3 * H + u];
                        REAL dht = dyt[u] + dh[u];
                        REAL tc = (REAL)tanh((double)ct[u]);
                        REAL dct = dht * go * (1 - tc * tc) + dc[u];
                        REAL cprev = cp ? cp[u] : 0;
                        dzt[u] = dct * gj * gi * (1 - gi);
                        dzt[H + u] = dct * gi * (1 - gj * gj);
                        dzt[2 * H + u] = dct * cprev * gf * (1 - gf);
                        dzt[3 * H + u] = dht * tc * go * (1 - go);
                        dc[u] = dct * gf;
                    }
                } else {
                    for (int u = 0; u < H; ++u) {
                        REAL hv = gt[u], dht = dyt[u] + dh[u];
                        dzt[u] = cell == 0 ? dht * (1 - hv * hv) : (hv > 0 ? dht : 0);
                    }
                }
                for (int k = 0; k < H; ++k) {
                    const REAL *wr = whd + (size_t)k * GH;
                    REAL s = 0;
                    for (int g = 0; g < GH; ++g) s += RB(dzt[g]) * RB(wr[g]);
                    dh[k] = s;
                }
            }
            free(dh); free(dc);
        }
    }
    /* parameter and input gradients from dz */
    memset(dwx, 0, sizeof(REAL) * (size_t)in * 2 * GH);
    memset(dwh, 0, sizeof(REAL) * (size_t)2 * H * GH);
    memset(dbias, 0, sizeof(REAL) * ((size_t)2 * GH + (cell == 3 ? 2 * H : 0)));
    if (dx) memset(dx, 0, sizeof(REAL) * (size_t)T * B * in);
    for (int d = 0; d < 2; ++d) {
        for (int t = 0; t < T; ++t) for (int b = 0; b < B; ++b) {
            const int len = use_len ? (seq_len[b] < T ? seq_len[b] : T) : T;
            if (t >= len) continue;
            const REAL *dzt = dz + (((size_t)d * T + t) * B + b) * GH;
            for (int g = 0; g < GH; ++g) dbias[d * GH + g] += dzt[g];
            if (cell == 3) {
                const REAL *dzrt = dzr + (((size_t)d * T + t) * B + b) * GH;
                for (int u = 0; u < H; ++u) dbias[2 * GH + d * H + u] += dzrt[2 * H + u];
            }
        }
#pragma omp parallel for
        for (int k = 0; k < in; ++k) {
            for (int t = 0; t < T; ++t) for (int b = 0; b < B; ++b) {
                const REAL xv = RB(x[((size_t)t * B + b) * in + k]);
                const REAL *dzt = dz + (((size_t)d * T + t) * B + b) * GH;
                REAL *o = dwx + (size_t)k * 2 * GH + d * GH;
                for (int g = 0; g < GH; ++g) o[g] += xv * RB(dzt[g]);
            }
        }
#pragma omp parallel for
        for (int k = 0; k < H; ++k) {
            for (int b = 0; b < B; ++b) {
                const int len = use_len ? (seq_len[b] < T ? seq_len[b] : T) : T;
                for (int step = 1; step < len; ++step) {
                    const int t = d == 0 ? step : len - 1 - step;
                    const int tp = d == 0 ? t - 1 : t + 1;
                    const REAL hv = RB(y[((size_t)tp * B + b) * 2 * H + d * H + k]);
                    const REAL *dzt = dzr + (((size_t)d * T + t) * B + b) * GH;
                    REAL *o = dwh + ((size_t)d * H + k) * GH;
                    for (int g = 0; g < GH; ++g) o[g] += hv * RB(dzt[g]);
                }
            }
        }
        if (dx) {
#pragma omp parallel for
            for (int tb = 0; tb < T * B; ++tb) {
                const REAL *dzt = dz + ((size_t)d * T * B + tb) * GH;
                for (int k = 0; k < in; ++k) {
                    const REAL *wr = wx + (size_t)k * 2 * GH + d * GH;
                    REAL s = 0;
                    for (int g = 0; g < GH; ++g) s += RB(dzt[g]) * RB(wr[g]);
                    dx[(size_t)tb * in + k] += s;
                }
            }
        }
    }
    if (dzr != dz) free(dzr);
    free(dz);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Adam, TF1 formulation (tf.train.AdamOptimizer as used at asr/model.py:79-83;
 * hyper-parameters asr/params.py:66-82):
 *   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr_t * m / (sqrt(v) + eps)
 * ------------------------------------------------------------------------------------------ */
int SUF(oracle_adam)(REAL *p, REAL *m, REAL *v, const REAL *g, size_t n, int step,
                     REAL lr, REAL b1, REAL b2, REAL eps)
{
    const REAL lr_t = (REAL)((double)lr * sqrt(1.0 - pow((double)b2, step)) / (1.0 - pow((double)b1, step)));
    for (size_t i = 0; i < n; ++i) {
        m[i] = b1 * m[i] + (1 - b1) * g[i];
        v[i] = b2 * v[i] + (1 - b2) * g[i] * g[i];
        p[i] -= lr_t * m[i] / ((REAL)sqrt((double)v[i]) + eps);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * 2-D convolution layer of the 'ds2' front-end.  tf.layers.conv2d(padding='SAME',
 * activation=relu) + tf.minimum(., relu_cutoff) as composed at asr/util/tf_contrib.py:123-134
 * (conv dropout: rate 0.0 by default, asr/params.py:89 — not restated).
 *
 * The reference feeds [batch, time, features, 1] (asr/model.py:157-158): the image height is
 * TIME, its width the feature axis.  Here activations are time-major like everywhere else in
 * this oracle:  x [T, B, F, C]  ->  y [To, B, Fo, N],  kernel w [kt, kf, C, N] (TF's HWIO).
 * TF 'SAME' padding [TF-RECALL, SURVEY.md Appendix A]: out = ceil(in / stride),
 * pad_total = max((out - 1) * stride + k - in, 0), pad_before = pad_total / 2 (the odd unit goes
 * after).
 * ------------------------------------------------------------------------------------------ */
static void SUF(conv_geom)(int in, int k, int s, int *out, int *pad0)
{
    *out = (in + s - 1) / s;
    int total = (*out - 1) * s + k - in;
    if (total < 0) total = 0;
    *pad0 = total / 2;
}

int SUF(oracle_conv2d_fwd)(const REAL *x, const REAL *w, const REAL *bias, REAL *y,
                           int T, int B, int F, int C, int kt, int kf, int st, int sf, int N,
                           int act, REAL cutoff)
{
    int To, Fo, pt, pf;
    SUF(conv_geom)(T, kt, st, &To, &pt);
    SUF(conv_geom)(F, kf, sf, &Fo, &pf);
#pragma omp parallel for collapse(2)
    for (int to = 0; to < To; ++to)
        for (int b = 0; b < B; ++b)
            for (int fo = 0; fo < Fo; ++fo) {
                REAL *yr = y + (((size_t)to * B + b) * Fo + fo) * N;
                for (int n = 0; n < N; ++n) yr[n] = bias ? bias[n] : 0;
                for (int it = 0; it < kt; ++it) {
                    const int t = to * st - pt + it;
                    if (t < 0 || t >= T) continue;
                    for (int jf = 0; jf < kf; ++jf) {
                        const int f = fo * sf - pf + jf;
                        if (f < 0 || f >= F) continue;
                        const REAL *xr = x + (((size_t)t * B + b) * F + f) * C;
                        const REAL *wr = w + ((size_t)it * kf + jf) * C * N;
                        for (int c = 0; c < C; ++c) {
                            const REAL xv = xr[c];
                            for (int n = 0; n < N; ++n) yr[n] += RB(xv) * RB(wr[(size_t)c * N + n]);
                        }
                    }
                }
                if (act == 1)
                    for (int n = 0; n < N; ++n) {
                        REAL v = yr[n];
                        v = v > 0 ? v : 0;
                        yr[n] = v < cutoff ? v : cutoff;
                    }
            }
    return 0;
}

/* Backward given the layer's output y (mask 0 < y < cutoff): dy [To,B,Fo,N] -> dx [T,B,F,C]
 * (nullable), dw [kt,kf,C,N], db [N]; all overwritten. */
int SUF(oracle_conv2d_bwd)(const REAL *x, const REAL *w, const REAL *y, const REAL *dy,
                           REAL *dx, REAL *dw, REAL *db,
                           int T, int B, int F, int C, int kt, int kf, int st, int sf, int N,
                           int act, REAL cutoff)
{
    int To, Fo, pt, pf;
    SUF(conv_geom)(T, kt, st, &To, &pt);
    SUF(conv_geom)(F, kf, sf, &Fo, &pf);
    const size_t rows = (size_t)To * B * Fo;
    REAL *dz = (REAL *)malloc(sizeof(REAL) * rows * N);
    for (size_t i = 0; i < rows * N; ++i) {
        REAL g = dy[i];
        if (act == 1 && !(y[i] > 0 && y[i] < cutoff)) g = 0;
        dz[i] = g;
    }
    for (int n = 0; n < N; ++n) {
        REAL s = 0;
        for (size_t r = 0; r < rows; ++r) s += dz[r * N + n];
        db[n] = s;
    }
    /* dw: one (it, jf) tap per thread -> disjoint outputs, deterministic */
#pragma omp parallel for collapse(2)
    for (int it = 0; it < kt; ++it)
        for (int jf = 0; jf < kf; ++jf) {
            REAL *dwr = dw + ((size_t)it * kf + jf) * C * N;
            for (size_t i = 0; i < (size_t)C * N; ++i) dwr[i] = 0;
            for (int to = 0; to < To; ++to) {
                const int t = to * st - pt + it;
                if (t < 0 || t >= T) continue;
                for (int b = 0; b < B; ++b)
                    for (int fo = 0; fo < Fo; ++fo) {
                        const int f = fo * sf - pf + jf;
                        if (f < 0 || f >= F) continue;
                        const REAL *xr = x + (((size_t)t * B + b) * F + f) * C;
                        const REAL *gz = dz + (((size_t)to * B + b) * Fo + fo) * N;
                        for (int c = 0; c < C; ++c)
                            for (int n = 0; n < N; ++n) dwr[(size_t)c * N + n] += RB(xr[c]) * RB(gz[n]);
                    }
            }
        }
    if (dx) {
        /* gather form: every input element sums the taps that touched it */
#pragma omp parallel for collapse(2)
        for (int t = 0; t < T; ++t)
            for (int b = 0; b < B; ++b)
                for (int f = 0; f < F; ++f) {
                    REAL *dxr = dx + (((size_t)t * B + b) * F + f) * C;
                    for (int c = 0; c < C; ++c) dxr[c] = 0;
                    for (int it = 0; it < kt; ++it) {
                        const int tn = t + pt - it;
                        if (tn < 0 || tn % st) continue;
                        const int to = tn / st;
                        if (to >= To) continue;
                        for (int jf = 0; jf < kf; ++jf) {
                            const int fn = f + pf - jf;
                            if (fn < 0 || fn % sf) continue;
                            const int fo = fn / sf;
                            if (fo >= Fo) continue;
                            const REAL *gz = dz + (((size_t)to * B + b) * Fo + fo) * N;
                            const REAL *wr = w + ((size_t)it * kf + jf) * C * N;
                            for (int c = 0; c < C; ++c) {
                                REAL s = 0;
                                for (int n = 0; n < N; ++n) s += RB(gz[n]) * RB(wr[(size_t)c * N + n]);
                                dxr[c] += s;
                            }
                        }
                    }
                }
    }
    free(dz);
    return 0;
}
