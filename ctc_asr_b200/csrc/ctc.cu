// ctc.cu — CTC loss forward-backward and greedy decode for sm_100a.
//
// Replaces tf.nn.ctc_loss (asr/model.py:259-264; a CPU-only op in TF 1.x that costs a D2H copy of
// the logits and an H2D copy of the gradient every step) and the greedy stand-in
// for decode_fn (asr/model.py:271-309).  Semantics: SURVEY.md Appendix A.7 / oracle_impl.h.
//
// Kernel design (one CTA per utterance, the 2L+1 blank-expanded lattice lives in shared memory):
//   phase 1  alpha sweep over t = 0..T_b-1 in chunks of CH frames.  Only one re-normalised alpha
//            row per chunk (a "checkpoint", S floats) goes to global memory, never the [T,S] table.
//   phase 2  for chunks from the last to the first, the alpha warps re-compute the chunk's alpha
//            rows from its checkpoint into shared memory while the beta warps sweep the
//            previously re-computed chunk backwards, turning alpha rows into posteriors in place;
//            then all warps reduce posteriors per class (deterministic CSR gather, no atomics)
//            and write the gradient rows.
//   HBM traffic is therefore the algorithmic one: logits in (re-read from L2 in phase 2),
//   gradient out, plus T/CH checkpoint rows.
//   alpha/beta are carried as extended-range numbers m * 2^e (m a float in [1, 2), e a 32-bit integer):
//   the recursion  alpha_t(s) = y * (alpha(s) + alpha(s-1) + alpha(s-2))  is then two additions and one
//   multiplication per state with 2^-24 relative rounding each (~1e-5 after 1700 frames) and no transcendental
//   at all; the softmax helper warp turns each frame's log-softmax into (m, e) pairs once per (t, class).
//   A plain fp32 LOG-domain recursion (what the TF kernel does) rounds every step at ulp(|alpha|) ~ 1e-3 and
//   carries ~1e-2 of gradient noise at T=1700 (tests/test_oracle_ctc.py); a linear-domain recursion with one
//   scale per row underflows the states that carry the posterior.  Each step is one shared-memory row
//   read + the (m, e) arithmetic + one named barrier per warp group (alpha and beta on separate barriers).
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

namespace ctcasr {
namespace ctc {

constexpr int kMaxSPT = 4;      // lattice states per thread (generic variant)
constexpr int kMaxGT = 448;     // threads per group (alpha / beta); + 1 helper warp <= 1024
constexpr int kHelper = 32;     // one producer warp: log-softmax of the next chunk's frames

struct Params {
    const float *logits; int T, B, V, blank;
    const int *labels; int lstride; const int *label_len; const int *seq_len;
    float *loss; float *grad; float grad_scale; int *status;
    float2 *ckpt;                   // [B][NCH][RS] alpha checkpoint rows (hi, lo)
    int CH, RS, GT, NCH, Lmax, VP;
    float2 *rows;                   // warp kernel: [B][T][32 * SPL] spilled alpha / beta rows
};

struct SmemLayout {
    size_t lab, csr_start, csr_pos, rowA, rowB, A, LY, red, total;
    __host__ __device__ SmemLayout(int Lmax, int V, int RS, int CH, int VP)
    {
        size_t o = 0;
        lab = o;       o += (size_t)((Lmax + 3) / 4 * 4 + 4) * 4;
        csr_start = o; o += (size_t)((V + 1 + 3) / 4 * 4) * 4;
        csr_pos = o;   o += (size_t)((Lmax + 3) / 4 * 4 + 4) * 4;
        rowA = o;      o += (size_t)2 * RS * 8;
        rowB = o;      o += (size_t)2 * RS * 8;
        A = o;         o += (size_t)2 * CH * RS * 8;
        LY = o;        o += (size_t)((3 * CH * VP + 3) / 4 * 4) * 8;        // (m, e) pairs
        red = o;       o += 64 * 4 + 64;
        total = o;
    }
};

// ---- lattice arithmetic ---------------------------------------------------------------------------
// A lattice value is m * 2^e stored as float2 {m, bits of the int e}: m in [1, 2), or m = 0 (then e = kZE)
// for probability zero.  Sums align the operands to the largest exponent with exact power-of-two factors
// built from integer bits; the product with y re-normalises the mantissa with two integer instructions.
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int kZE = -(1 << 29);
__device__ __forceinline__ float2 xf_make(float m, int e) { return make_float2(m, __int_as_float(e)); }
__device__ __forceinline__ int xf_e(const float2 v) { return __float_as_int(v.y); }
// 2^d for d <= 0 (0 below 2^-126: such a term is below the rounding of its partners by > 100 binades)
__device__ __forceinline__ float pow2_le0(int d) { return __int_as_float(max(d + 127, 0) << 23); }
// 2^e clamped to the normal range (posteriors and y are <= 1 up to rounding)
__device__ __forceinline__ float pow2_clamp(int e) { return __int_as_float(min(max(e + 127, 0), 254) << 23); }

// (a0 + a1 + a2) * y
__device__ __forceinline__ float2 xf_sum3_mul(const float2 a0, const float2 a1, const float2 a2, const float2 y)
{
    const int e0 = xf_e(a0), e1 = xf_e(a1), e2 = xf_e(a2);
    const int em = max(e0, max(e1, e2));
    float s = a0.x * pow2_le0(e0 - em);
    s = fmaf(a1.x, pow2_le0(e1 - em), s);
    s = fmaf(a2.x, pow2_le0(e2 - em), s);
    s *= y.x;                                       // [0, 12)
    if (s == 0.f) return xf_make(0.f, kZE);
    const int bits = __float_as_int(s);
    return xf_make(__int_as_float((bits & 0x007fffff) | 0x3f800000), em + xf_e(y) + (bits >> 23) - 127);
}
// log2 of a lattice value, in double (loss only)
__device__ __forceinline__ double xf_log2(const float2 v) { return v.x == 0.f ? -INFINITY : log2((double)v.x) + (double)xf_e(v); }

__device__ __forceinline__ void group_bar(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// softmax of the frames of chunk c into LY as (m, e) pairs, FOUR LANES PER FRAME (a frame's V logits are a
// contiguous 4V-byte run; the 4 lanes take classes k = sub, sub+4, ... and combine with two
// shuffles), executed by `nthreads` (a multiple of 32) consecutive threads starting at `first`.
// Values are re-read from L1 instead of being held in registers (V up to 128).
// Rows are VP = V|1 entries apart to spread the banks.
__device__ __forceinline__ void compute_logy(const Params &p, int b, int Tb, int c, float2 *LY, int first, int nthreads)
{
    const float kLog2e = 1.4426950408889634f;
    const int lt = (int)threadIdx.x - first, sub = lt & 3, rpp = nthreads >> 2;
    const int lo = c * p.CH, hi = min(lo + p.CH, Tb);
    for (int t0 = lo; t0 < hi; t0 += rpp) {            // uniform trip count: shuffles need the whole warp
        const int t = t0 + (lt >> 2);
        const bool valid = t < hi;
        const float *x = p.logits + ((size_t)(valid ? t : lo) * p.B + b) * p.V;
        float m = -INFINITY;
        for (int k = sub; k < p.V; k += 4) m = fmaxf(m, __ldg(x + k));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        float e = 0.f;
        for (int k = sub; k < p.V; k += 4) e += ex2f((__ldg(x + k) - m) * kLog2e);
        e += __shfl_xor_sync(0xffffffffu, e, 1);
        e += __shfl_xor_sync(0xffffffffu, e, 2);
        const float lg = lg2f(e);                       // log2 sum_k 2^((x_k - max) log2 e), in [0, log2 V]
        if (valid) {
            float2 *row = LY + (size_t)(t - lo) * p.VP;
            for (int k = sub; k < p.V; k += 4) {
                const float l2 = (__ldg(x + k) - m) * kLog2e - lg;      // log2 softmax_k <= 0, small near the maximum
                const float fl = floorf(l2);
                row[k] = l2 > -1e9f ? xf_make(ex2f(l2 - fl), (int)fl) : xf_make(0.f, kZE);
            }
        }
    }
}

template <int SPT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
ctc_loss_kernel(const Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int GT = p.GT, NT = 2 * GT + kHelper;
    const bool is_alpha = tid < GT;
    const bool is_beta = !is_alpha && tid < 2 * GT;
    const bool is_helper = tid >= 2 * GT;
    const int gt = is_alpha ? tid : tid - GT;
    const int RS = p.RS, CH = p.CH, V = p.V, VP = p.VP, blank = p.blank;
    const float2 kNegInf = xf_make(0.f, kZE);                        // probability zero
    const float2 kOne = xf_make(1.f, 0);

    const SmemLayout L(p.Lmax, V, RS, CH, VP);
    int *lab = reinterpret_cast<int *>(smem_raw + L.lab);
    int *csr_start = reinterpret_cast<int *>(smem_raw + L.csr_start);
    int *csr_pos = reinterpret_cast<int *>(smem_raw + L.csr_pos);
    float2 *rowA = reinterpret_cast<float2 *>(smem_raw + L.rowA);   // alpha rows: state s at [s+2]
    float2 *rowB = reinterpret_cast<float2 *>(smem_raw + L.rowB);   // beta rows (incl. y_t): state s at [s]
    float2 *Abuf = reinterpret_cast<float2 *>(smem_raw + L.A);      // [2][CH][RS], state s at [s+2]
    float2 *LYbuf = reinterpret_cast<float2 *>(smem_raw + L.LY);    // [3][CH][VP] ring of (m, e), chunk c at c % 3
    float *red = reinterpret_cast<float *>(smem_raw + L.red);
    int *flags = reinterpret_cast<int *>(red + 48);                 // [0] bad label, [1] repeats

    const int Tb = p.seq_len[b];
    const int Ln = p.label_len[b];
    const int S = 2 * Ln + 1;
    float *grad_b = p.grad ? p.grad + (size_t)b * V : nullptr;      // row t at + t*B*V
    const size_t gstride = (size_t)p.B * V;

    // ---- setup: labels, validation ---------------------------------------------------------
    if (tid < 2) flags[tid] = 0;
    __syncthreads();
    int status = CTCASR_CTC_OK;
    if (Tb > p.T || Tb < 0 || Ln < 0 || Ln > p.Lmax) status = CTCASR_CTC_BAD_LENGTH;
    if (status == CTCASR_CTC_OK) {
        int bad = 0, rep = 0;
        for (int i = tid; i < Ln; i += NT) {
            const int v = p.labels[(size_t)b * p.lstride + i];
            lab[i] = v;
            if (v < 0 || v >= V || v == blank) bad = 1;
            if (i > 0 && v == p.labels[(size_t)b * p.lstride + i - 1]) ++rep;
        }
        if (bad) atomicOr(&flags[0], 1);
        if (rep) atomicAdd(&flags[1], rep);
    }
    __syncthreads();
    if (status == CTCASR_CTC_OK) {
        if (flags[0]) status = CTCASR_CTC_BAD_LABEL;
        else if (Tb < Ln + flags[1]) status = CTCASR_CTC_INFEASIBLE;
    }
    // gradient rows the recursion never touches are zero (t >= T_b, or the whole utterance)
    if (grad_b) {
        const int t0 = (status == CTCASR_CTC_OK) ? Tb : 0;
        for (int i = tid; i < (p.T - t0) * V; i += NT)
            grad_b[(size_t)(t0 + i / V) * gstride + i % V] = 0.f;
    }
    if (status != CTCASR_CTC_OK || Tb == 0) {
        if (tid == 0) {
            p.status[b] = status;
            p.loss[b] = status == CTCASR_CTC_OK ? 0.f : INFINITY;
        }
        return;
    }

    // ---- per-class position lists of the label states (odd s), built in label order ----------
    if (tid == 0) {
        for (int k = 0; k <= V; ++k) csr_start[k] = 0;
        for (int i = 0; i < Ln; ++i) csr_start[lab[i] + 1]++;
        for (int k = 0; k < V; ++k) csr_start[k + 1] += csr_start[k];
    }
    // guards: alpha rows [0],[1]; beta rows [S],[S+1]; every A row [0],[1]
    for (int i = tid; i < 2 * RS; i += NT) { rowA[i] = kNegInf; rowB[i] = kNegInf; }
    for (int i = tid; i < 2 * CH; i += NT) { Abuf[(size_t)i * RS] = kNegInf; Abuf[(size_t)i * RS + 1] = kNegInf; }
    __syncthreads();
    if (tid == 0) {
        // csr_pos filled by a serial stable pass (deterministic summation order later)
        int *cursor = reinterpret_cast<int *>(LYbuf);     // scratch (>= V ints); LY is rewritten before use
        for (int k = 0; k < V; ++k) cursor[k] = csr_start[k];
        for (int i = 0; i < Ln; ++i) csr_pos[cursor[lab[i]]++] = 2 * i + 1;
    }
    __syncthreads();

    // per-thread lattice-state constants
    int st_lp[SPT];
    bool st_skA[SPT], st_skB[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
        const int s = gt + q * GT;
        st_lp[q] = blank; st_skA[q] = false; st_skB[q] = false;
        if (!is_helper && s < S && (s & 1)) {
            const int li = s >> 1;
            st_lp[q] = lab[li];
            st_skA[q] = li > 0 && lab[li] != lab[li - 1];
            st_skB[q] = li + 1 < Ln && lab[li + 1] != lab[li];
        }
    }

    const int NCH = (Tb + CH - 1) / CH;
    float2 *ck_b = p.ckpt + (size_t)b * p.NCH * RS;
    const size_t LYS = (size_t)CH * VP;

    // =========================== phase 1: alpha sweep with checkpoints =========================
    int cur = 0;
    if (is_alpha) {       // virtual row t = -1: {0, -inf, ...} reproduces TF's alpha init
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
            const int s = gt + q * GT;
            if (s < S) rowA[cur * RS + s + 2] = s == 0 ? kOne : kNegInf;
        }
    }
    compute_logy(p, b, Tb, 0, LYbuf, 0, NT);
    __syncthreads();
    for (int c = 0; c < NCH; ++c) {
        const float2 *LY = LYbuf + (size_t)(c % 3) * LYS;
        if (is_helper) {
            if (c + 1 < NCH) compute_logy(p, b, Tb, c + 1, LYbuf + (size_t)((c + 1) % 3) * LYS, 2 * GT, kHelper);
        } else if (is_alpha) {
            // checkpoint: the row before the chunk's first frame
#pragma unroll
            for (int q = 0; q < SPT; ++q) {
                const int s = gt + q * GT;
                if (s < S) ck_b[(size_t)c * RS + s] = rowA[cur * RS + s + 2];
            }
            const int lo = c * CH, hi = min(lo + CH, Tb);
            for (int t = lo; t < hi; ++t) {
                const float2 *prev = rowA + cur * RS;
                float2 *next = rowA + (cur ^ 1) * RS;
                const float2 *ly = LY + (size_t)(t - lo) * VP;
#pragma unroll
                for (int q = 0; q < SPT; ++q) {
                    const int s = gt + q * GT;
                    if (s < S) next[s + 2] = xf_sum3_mul(prev[s + 2], prev[s + 1], st_skA[q] ? prev[s] : kNegInf, ly[st_lp[q]]);
                }
                cur ^= 1;
                group_bar(1, GT);
            }
        }
        __syncthreads();
    }
    // p = alpha[S-1] + alpha[S-2] at t = T_b - 1, shared through smem
    float2 *psh = reinterpret_cast<float2 *>(red + 40);
    if (tid == 0) {
        const float2 a = rowA[cur * RS + (S - 1) + 2];
        const float2 c2 = S > 1 ? rowA[cur * RS + (S - 2) + 2] : kNegInf;
        const float2 pp = xf_sum3_mul(a, c2, kNegInf, kOne);
        psh[0] = pp;
        p.loss[b] = (float)(-xf_log2(pp) * 0.6931471805599453);
        p.status[b] = CTCASR_CTC_OK;
    }
    if (!p.grad) return;
    __syncthreads();
    const float2 lp2 = psh[0];

    // =========================== phase 2: alpha re-compute || beta sweep =======================
    // LY ring on entry: chunks NCH-1, NCH-2 (and NCH-3) are resident from phase 1.
    // beta here INCLUDES y_t (Graves' convention); the posterior subtracts log y_t once.
    int curB = 0;
    if (is_beta) {        // virtual row t = T_b: {.., -inf, 0 at S-1}
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
            const int s = gt + q * GT;
            if (s < S) rowB[curB * RS + s] = s == S - 1 ? kOne : kNegInf;
        }
    }
    for (int r = 0; r <= NCH; ++r) {
        const int ca = NCH - 1 - r;     // chunk the alpha group re-computes
        const int cb = NCH - r;         // chunk the beta group sweeps
        if (is_helper) {
            // next round's alpha chunk; (ca-1) % 3 is the ring slot neither group reads this round
            if (ca - 1 >= 0 && r >= 2) compute_logy(p, b, Tb, ca - 1, LYbuf + (size_t)((ca - 1) % 3) * LYS, 2 * GT, kHelper);
        } else if (is_alpha) {
            if (ca >= 0) {
                const float2 *LY = LYbuf + (size_t)(ca % 3) * LYS;
                float2 *A = Abuf + (size_t)(ca & 1) * CH * RS;
                float2 *r0 = rowA;      // checkpoint row of chunk ca
#pragma unroll
                for (int q = 0; q < SPT; ++q) {
                    const int s = gt + q * GT;
                    if (s < S) r0[s + 2] = ck_b[(size_t)ca * RS + s];
                }
                group_bar(1, GT);
                const int lo = ca * CH, hi = min(lo + CH, Tb);
                for (int t = lo; t < hi; ++t) {
                    const float2 *prev = t == lo ? r0 : A + (size_t)(t - lo - 1) * RS;
                    float2 *next = A + (size_t)(t - lo) * RS;
                    const float2 *ly = LY + (size_t)(t - lo) * VP;
#pragma unroll
                    for (int q = 0; q < SPT; ++q) {
                        const int s = gt + q * GT;
                        if (s < S) next[s + 2] = xf_sum3_mul(prev[s + 2], prev[s + 1], st_skA[q] ? prev[s] : kNegInf, ly[st_lp[q]]);
                    }
                    group_bar(1, GT);
                }
            }
        } else if (cb < NCH) {
            const float2 *LY = LYbuf + (size_t)(cb % 3) * LYS;
            float2 *A = Abuf + (size_t)(cb & 1) * CH * RS;
            const int lo = cb * CH, hi = min(lo + CH, Tb);
            const float inv_pm = 1.f / lp2.x;
            const int pe = xf_e(lp2);
            for (int t = hi - 1; t >= lo; --t) {
                const float2 *prev = rowB + curB * RS;
                float2 *next = rowB + (curB ^ 1) * RS;
                const float2 *ly = LY + (size_t)(t - lo) * VP;
                float2 *arow = A + (size_t)(t - lo) * RS;
#pragma unroll
                for (int q = 0; q < SPT; ++q) {
                    const int s = gt + q * GT;
                    if (s < S) {
                        const float2 lys = ly[st_lp[q]];
                        const float2 bt = xf_sum3_mul(prev[s], prev[s + 1], st_skB[q] ? prev[s + 2] : kNegInf, lys);
                        next[s] = bt;
                        // posterior of state s at t: alpha * beta / (y * p)   (beta includes y_t)
                        const float2 al = arow[s + 2];
                        float post = 0.f;
                        if (al.x != 0.f && bt.x != 0.f)
                            post = __fdividef(al.x * bt.x, lys.x) * inv_pm * pow2_clamp(xf_e(al) + xf_e(bt) - xf_e(lys) - pe);
                        arow[s + 2].x = post;
                    }
                }
                curB ^= 1;
                group_bar(2, GT);
            }
        }
        __syncthreads();
        // ---- gradient rows of chunk cb: y - sum_{s in class k} posterior ---------------------
        if (cb < NCH) {
            const float2 *LY = LYbuf + (size_t)(cb % 3) * LYS;
            const float2 *A = Abuf + (size_t)(cb & 1) * CH * RS;
            const int lo = cb * CH, hi = min(lo + CH, Tb);
            const int warp = tid >> 5, lane = tid & 31, nwarps = NT >> 5;
            for (int t = lo + warp; t < hi; t += nwarps) {
                const float2 *arow = A + (size_t)(t - lo) * RS + 2;
                float pb = 0.f;                                     // blank states: even s
                for (int i = lane; 2 * i < S; i += 32) pb += arow[2 * i].x;
                pb = warp_sum(pb);
                for (int k = lane; k < V; k += 32) {
                    float acc = 0.f;
                    if (k == blank) acc = pb;
                    else for (int q = csr_start[k]; q < csr_start[k + 1]; ++q) acc += arow[csr_pos[q]].x;
                    const float2 yk = LY[(size_t)(t - lo) * VP + k];
                    const float y = yk.x * pow2_clamp(xf_e(yk));
                    grad_b[(size_t)t * gstride + k] = (y - acc) * p.grad_scale;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// greedy decode: one warp per utterance, 32 frames per iteration
// ------------------------------------------------------------------------------------------------
__global__ void greedy_decode_kernel(const float *logits, int T, int B, int V, int blank,
                                     const int *seq_len, int *out_ids, int *out_len)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int Tb = min(seq_len[b], T);
    int n = 0, carry = -1;
    for (int t0 = 0; t0 < Tb; t0 += 32) {
        const int t = t0 + lane;
        int am = -1;
        if (t < Tb) {
            const float *x = logits + ((size_t)t * B + b) * V;
            float best = x[0];
            am = 0;
            for (int k = 1; k < V; ++k) {
                const float v = x[k];
                if (v > best) { best = v; am = k; }     // strict >: first max wins
            }
        }
        int prev = __shfl_up_sync(0xffffffffu, am, 1);
        if (lane == 0) prev = carry;
        const bool emit = t < Tb && am != blank && am != prev;
        const unsigned mask = __ballot_sync(0xffffffffu, emit);
        if (emit) out_ids[(size_t)b * T + n + __popc(mask & ((1u << lane) - 1))] = am;
        n += __popc(mask);
        carry = __shfl_sync(0xffffffffu, am, 31);
    }
    for (int i = n + lane; i < T; i += 32) out_ids[(size_t)b * T + i] = -1;
    if (lane == 0) out_len[b] = n;
}

// ------------------------------------------------------------------------------------------------
// Levenshtein distance between label sequences (tf.edit_distance, asr/model.py:338): one CTA per
// pair, anti-diagonal wavefront over three rolling diagonals in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void edit_distance_kernel(const int *hyp, int hstride, const int *hyp_len, const int *truth, int tstride,
                                     const int *truth_len, int normalize, float *out)
{
    extern __shared__ int ed_smem[];
    const int b = blockIdx.x;
    const int n = hyp_len[b], m = truth_len[b];
    const int *h = hyp + (size_t)b * hstride, *t = truth + (size_t)b * tstride;
    int *d0 = ed_smem, *d1 = d0 + (n + 2), *d2 = d1 + (n + 2);       // diagonals k-2, k-1, k indexed by i
    // D[i][j], i in [0,n], j in [0,m]; cell (i, j) lies on diagonal k = i + j
    for (int k = 0; k <= n + m; ++k) {
        const int ilo = max(0, k - m), ihi = min(n, k);
        for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
            const int j = k - i;
            int v;
            if (i == 0) v = j;
            else if (j == 0) v = i;
            else {
                const int sub = d0[i - 1] + (h[i - 1] != t[j - 1] ? 1 : 0);
                v = min(sub, min(d1[i - 1] + 1, d1[i] + 1));          // (i-1, j) and (i, j-1) are on diagonal k-1
            }
            d2[i] = v;
        }
        __syncthreads();
        int *tmp = d0; d0 = d1; d1 = d2; d2 = tmp;
    }
    if (threadIdx.x == 0) {
        const int dist = d1[n];
        float r = (float)dist;
        if (normalize) r = m > 0 ? (float)dist / (float)m : (n > 0 ? INFINITY : 0.f);
        out[b] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-shuffle CTC (the common case: 2L+1 <= 384 lattice states, V <= 32 classes).
//
// One CTA of two warps per utterance, no block barrier inside the time loop:
//   warp 0 walks t = 0 .. T_b-1 with the alpha recursion, warp 1 walks t = T_b-1 .. 0 with the beta recursion, AT THE
//   SAME TIME.  Lane l holds the SPL consecutive lattice states s = l*SPL .. l*SPL+SPL-1 in registers as
//   extended-range (m, e) numbers; a step needs the two neighbouring states of the next lane (4 shuffles), the
//   frame's softmax (computed by the warp itself: lane k owns class k, max / sum by shuffles, prefetched one frame
//   ahead) and y[l'_s] (one shuffle pair per label state).  Nothing of the recursion touches shared memory.
//   During the FIRST half of its walk a warp spills its rows to HBM ([T][32 lanes][SPL] float2, coalesced 16-B
//   stores); the warps meet in the middle (one 64-thread barrier), p(labels | x) = sum_s alpha_{h-1}(s) beta~_h(s)
//   is formed there, and during the SECOND half each warp reads the other warp's spilled row of the same frame,
//   forms the posteriors alpha beta / (y p) and writes the gradient row.  Two concurrent sweeps of T steps
//   replace the three (alpha, alpha re-computation, beta) of the block-per-utterance kernel above, which stays
//   for longer label sequences and wider alphabets.
//   The recursion runs on UN-normalised y'_k = exp(x_k): a per-frame factor common to all states cancels in the
//   posteriors, and log2 Z_t is taken out of the loss, so no reduction sits on the recursion's dependency chain.
//   Per-class sums of the posteriors and the frame's Z_t are integer (fixed-point) warp reductions and shared-memory
//   atomics: integer addition commutes, so the result is the same bits in any order.
//   Logits and the other warp's rows arrive through a cp.async ring kPF frames ahead.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 xf_add(const float2 a, const float2 b)
{
    const int ea = xf_e(a), eb = xf_e(b), em = max(ea, eb);
    const float s = fmaf(a.x, pow2_le0(ea - em), b.x * pow2_le0(eb - em));
    if (s == 0.f) return xf_make(0.f, kZE);
    const int bits = __float_as_int(s);
    return xf_make(__int_as_float((bits & 0x007fffff) | 0x3f800000), em + (bits >> 23) - 127);
}
__device__ __forceinline__ float2 xf_mul(const float2 a, const float2 b)
{
    const float s = a.x * b.x;
    if (s == 0.f) return xf_make(0.f, kZE);
    const int bits = __float_as_int(s);
    return xf_make(__int_as_float((bits & 0x007fffff) | 0x3f800000), xf_e(a) + xf_e(b) + (bits >> 23) - 127);
}
__device__ __forceinline__ float2 shfl2(const float2 v, int src)
{
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

constexpr int kPF = 8;             // frames the warp kernel prefetches ahead (cp.async ring)

struct WarpSmem {
    size_t lab, hand, post, ring, total;
    __host__ __device__ WarpSmem(int Lmax, int V, int RS)
    {
        (void)V;
        size_t o = 0;
        lab = o;       o += (size_t)((Lmax + 3) / 4 * 4 + 4) * 4;
        o = (o + 15) / 16 * 16;
        hand = o;      o += (size_t)2 * (RS + 4) * 8 + 32;         // alpha_{h-1} and beta_h rows at the hand-over, 2 doubles
        post = o;      o += (size_t)2 * 2 * 32 * 4;                // [warp][2 buffers][32] fixed-point class sums of a frame
        ring = o;      o += (size_t)2 * kPF * (128 + (size_t)RS * 8);   // [warp][kPF slots]{32 logits, the other warp's row}
        total = o + 32;
    }
};

// (destinations as shared-space addresses computed once: a generic pointer costs a window conversion per copy)
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 lds_f32x4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int warp_max_int(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- arithmetic of the warp kernel: "probability zero" is the ordinary number 1.0 x 2^kZE (kZE = -2^29): it drops out
// of every sum through the exponent clamp below and stays ~kZE through thousands of products, so neither the recursion
// nor the posteriors need a zero test or a select on the mantissa.
// x * 2^d for x in [1, 2), d <= 0: one integer multiply-add on the exponent field, 2^-127 (flushed) at the clamp
__device__ __forceinline__ float scale_le0(float x, int d) { return __int_as_float(__float_as_int(x) + (max(d, -127) << 23)); }
// (a0 + a1 + a2) * y with a2 optional; every mantissa is in [1, 2)
template <bool THREE>
__device__ __forceinline__ float2 xf_step(const float2 a0, const float2 a1, const float a2m, const int e2, const float2 y)
{
    const int e0 = xf_e(a0), e1 = xf_e(a1);
    const int em = THREE ? max(e0, max(e1, e2)) : max(e0, e1);
    float s = scale_le0(a0.x, e0 - em) + scale_le0(a1.x, e1 - em);
    if (THREE) s += scale_le0(a2m, e2 - em);
    s *= y.x;                                       // [1, 12)
    const int bits = __float_as_int(s);
    // (the clamp keeps a zero times a zero at kZE: exponents must not walk towards the integer range's end)
    return xf_make(__int_as_float((bits & 0x007fffff) | 0x3f800000), max(em + xf_e(y) + (bits >> 23) - 127, kZE));
}

// one direction of the warp kernel: FWD = alpha (frames ascending), !FWD = beta (frames descending)
template <int SPL, bool FWD>
__device__ __forceinline__ void ctc_sweep(const Params &p, unsigned char *smem_raw, const WarpSmem &L, int b, int lane, int Tb, int Ln)
{
    constexpr int RS = 32 * SPL;
    constexpr size_t SLOT = 128 + (size_t)RS * 8;
    constexpr int warp = FWD ? 0 : 1;
    const int V = p.V, blank = p.blank, S = 2 * Ln + 1;
    const float2 kZero = xf_make(1.f, kZE), kOne = xf_make(1.f, 0);     // zero = 1.0 x 2^kZE, see xf_step
    const int *lab = reinterpret_cast<const int *>(smem_raw + L.lab);
    float2 *hand = reinterpret_cast<float2 *>(smem_raw + L.hand);             // [2][RS + 4]
    double *hand_lz = reinterpret_cast<double *>(smem_raw + L.hand + (size_t)2 * (RS + 4) * 8);
    unsigned *postrow = reinterpret_cast<unsigned *>(smem_raw + L.post) + (size_t)warp * 2 * 32;      // [2 buffers][32] fixed-point class sums
    unsigned char *ring = smem_raw + L.ring + (size_t)warp * kPF * SLOT;
    float *grad_b = p.grad ? p.grad + (size_t)b * V : nullptr;                 // row t at + t*B*V
    const size_t gstride = (size_t)p.B * V;

    // ---- per-lane lattice constants: SPL is even, so state j of a lane is a blank for even j, a label for odd j ----
    int cls[SPL];
    unsigned skip = 0, validm = 0;      // skip: the s-2 (alpha) / s+2 (beta) transition exists
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        cls[j] = blank;
        if (s < S) validm |= 1u << j;
        if (s < S && (j & 1)) {
            const int li = s >> 1;
            cls[j] = lab[li];
            if (FWD ? (li > 0 && lab[li] != lab[li - 1]) : (li + 1 < Ln && lab[li + 1] != lab[li])) skip |= 1u << j;
        }
    }
    unsigned skipB = 0;                 // the beta-side skips are also needed at the hand-over
#pragma unroll
    for (int j = 1; j < SPL; j += 2) {
        const int s = lane * SPL + j, li = s >> 1;
        if (s < S && li + 1 < Ln && lab[li + 1] != lab[li]) skipB |= 1u << j;
    }

    constexpr int kOff = -(1 << 28);                    // added to an exponent: "this term does not exist"
    int voff[SPL], skoff[SPL / 2];
#pragma unroll
    for (int j = 0; j < SPL; ++j) voff[j] = (validm >> j) & 1 ? 0 : kOff;
#pragma unroll
    for (int j = 1; j < SPL; j += 2) skoff[j / 2] = (skip >> j) & 1 ? 0 : kOff;
    const int edge_off = (FWD ? lane == 0 : lane == 31) ? kOff : 0;

    const int h = p.grad ? Tb / 2 : Tb;                 // alpha owns frames [0, h) first, then [h, T_b); beta the reverse
    float2 *rows_b = p.rows + (size_t)b * p.T * RS + (size_t)lane * SPL;        // row t at + t*RS
    const float *logit_b = p.logits + (size_t)b * V + (lane < V ? lane : 0);
    const float kLog2e = 1.4426950408889634f;

    float2 a[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {     // virtual rows: alpha_{-1} = 1 at s = 0; beta_{T_b} = 1 at s = S-1
        const int s = lane * SPL + j;
        a[j] = (FWD ? s == 0 : s == S - 1) ? kOne : kZero;
    }
    double lz = 0.0;                                    // sum over my first-half frames of log2 Z_t
    const int nsteps = (!FWD && !p.grad) ? 0 : Tb;
    const int own_first = FWD ? h : Tb - h;             // steps of my walk before the hand-over
    auto frame_of = [&](int n) { return FWD ? n : Tb - 1 - n; };
    // prefetch of the next step (they are issued strictly in order) into ring slot (step % kPF): the frame's logits (lane
    // k: class k) and, in the second half, the other warp's row of that frame (each lane its own SPL states); one
    // cp.async group per step.  Addresses are running pointers: one add per step instead of a 64-bit multiply each.
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t my_logit_s = ring_s + lane * 4, my_row_s = ring_s + 128 + lane * SPL * 8;
    const ptrdiff_t lstep = FWD ? (ptrdiff_t)gstride : -(ptrdiff_t)gstride, rstep = FWD ? (ptrdiff_t)RS : -(ptrdiff_t)RS;
    int iss = 0;
    uint32_t iss_off = 0;                                                   // (iss % kPF) * SLOT
    const float *iss_logit = logit_b + (size_t)frame_of(0) * gstride;
    const float2 *iss_row = rows_b + (size_t)frame_of(0) * RS;
    auto issue = [&](bool with_row) {
        if (iss < nsteps) {
            if (lane < V) cp_async4(my_logit_s + iss_off, iss_logit);
            if (with_row) {
#pragma unroll
                for (int q = 0; q < SPL / 2; ++q) cp_async16(my_row_s + iss_off + q * 16, reinterpret_cast<const float4 *>(iss_row) + q);
            }
        }
        cp_async_commit();
        ++iss; iss_off = iss_off + SLOT == kPF * SLOT ? 0u : iss_off + (uint32_t)SLOT;
        iss_logit += lstep; iss_row += rstep;
    };
    // one recursion step on frame n of my walk; returns y' of my lane's class, the per-state y' and this frame's Z pieces
    float2 ys[SPL];
    float yrel, zs;
    int emax;
    float4 o[SPL / 2];
    uint32_t cur_off = 0;                                                   // (n % kPF) * SLOT of the step being computed
    auto step = [&](bool second) {
        cp_async_wait<kPF - 1>();
        const float x = lane < V ? lds_f32(my_logit_s + cur_off) : 0.f;
        if (second) {
#pragma unroll
            for (int q = 0; q < SPL / 2; ++q) o[q] = lds_f32x4(my_row_s + cur_off + q * 16);
        }
        cur_off = cur_off + SLOT == kPF * SLOT ? 0u : cur_off + (uint32_t)SLOT;
        issue(second);                  // rows only once the other warp has written them (after the hand-over)
        // y'_k = exp(x_k) as (m, e): no normalisation on the recursion's path (a per-frame factor of all states cancels in
        // the posteriors and is taken out of the loss through log2 Z_t)
        const float l2 = x * kLog2e;
        const float fl = floorf(l2);
        const float2 y = (lane < V && l2 > -1e9f) ? xf_make(ex2f(l2 - fl), (int)fl) : kZero;      // logit -inf: probability zero
        const float2 yblank = shfl2(y, blank);
        // the two states beyond my own range: from the previous (alpha) / next (beta) lane; the edge lane's get the
        // exponent of zero through edge_off
        float2 n1, n2;
        if (FWD) {
            n1 = make_float2(__shfl_up_sync(0xffffffffu, a[SPL - 1].x, 1), __shfl_up_sync(0xffffffffu, a[SPL - 1].y, 1));
            n2 = make_float2(__shfl_up_sync(0xffffffffu, a[SPL - 2].x, 1), __shfl_up_sync(0xffffffffu, a[SPL - 2].y, 1));
        } else {
            n1 = make_float2(__shfl_down_sync(0xffffffffu, a[0].x, 1), __shfl_down_sync(0xffffffffu, a[0].y, 1));
            n2 = make_float2(__shfl_down_sync(0xffffffffu, a[1].x, 1), __shfl_down_sync(0xffffffffu, a[1].y, 1));
        }
        n1.y = __int_as_float(max(xf_e(n1) + edge_off, kZE));
        n2.y = __int_as_float(max(xf_e(n2) + edge_off, kZE));
        float2 nw[SPL];
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const float2 yl = (j & 1) ? shfl2(y, cls[j]) : yblank;
            // states beyond 2L+1 and absent skip transitions: exponent pushed to "zero" by a per-state constant (one add that
            // the clamp in xf_step absorbs) instead of a select on a predicate the loop would have to rebuild every step
            ys[j] = make_float2(yl.x, __int_as_float(max(xf_e(yl) + voff[j], kZE)));
            const float2 s1 = FWD ? (j >= 1 ? a[j - 1] : n1) : (j + 1 < SPL ? a[j + 1] : n1);
            if (j & 1) {        // label state: the skip transition exists when the neighbouring labels differ
                const float2 s2 = FWD ? (j >= 2 ? a[j - 2] : n1) : (j + 2 < SPL ? a[j + 2] : n2);      // j is odd: s-2 of j = 1 is the previous lane's last state
                nw[j] = xf_step<true>(a[j], s1, s2.x, max(xf_e(s2) + skoff[j / 2], kZE), ys[j]);
            } else {            // blank state: never skipped into
                nw[j] = xf_step<false>(a[j], s1, 1.f, kZE, ys[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < SPL; ++j) a[j] = nw[j];
        // Z_t = sum_k y'_k, off the recursion's path: for the loss (first half) and the softmax of the gradient.
        // (integer warp reductions — one redux.sync each — instead of shuffle trees: every dependent shuffle is ~25 cycles of
        // an in-order warp that has no other warp to hide behind; fixed point keeps the sums order-independent)
        emax = __reduce_max_sync(0xffffffffu, lane < V ? xf_e(y) : kZE);
        yrel = lane < V ? scale_le0(y.x, xf_e(y) - emax) : 0.f;     // (2^-127 at the clamp: nothing at 2^-22 resolution)                                           // [0, 2)
        zs = (float)__reduce_add_sync(0xffffffffu, __float2uint_rn(yrel * 4194304.f)) * (1.f / 4194304.f);      // 2^22: 32 x 2 x 2^22 < 2^32
    };

    for (int i = 0; i < kPF; ++i) issue(false);
    // ---- first half: recursion, spill my rows -----------------------------------------------------------
    float2 *spill = rows_b + (size_t)frame_of(0) * RS;
    for (int n = 0; n < own_first; ++n) {
        step(false);
        lz += (double)emax + (double)lg2f(zs);
        float4 *dst = reinterpret_cast<float4 *>(spill);
#pragma unroll
        for (int q = 0; q < SPL / 2; ++q) __stcg(dst + q, make_float4(a[2 * q].x, a[2 * q].y, a[2 * q + 1].x, a[2 * q + 1].y));
        spill += rstep;
    }
    // ---- hand-over: both warps have finished their first half ------------------------------------------------
    if (p.grad || FWD) {
#pragma unroll
        for (int j = 0; j < SPL; ++j) hand[(size_t)warp * (RS + 4) + lane * SPL + j] = a[j];
    }
    if (lane == 0) hand_lz[warp] = lz;
    asm volatile("bar.sync 1, 64;" ::: "memory");
    // p' = sum_s alpha'_{h-1}(s) (beta'_h(s) + beta'_h(s+1) + [skip] beta'_h(s+2)); without a gradient h = T_b and beta_h is the virtual row
    float2 lp2;
    {
        const float2 *ha = hand, *hb = hand + (RS + 4);
        float2 acc = kZero;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const int s = lane * SPL + j;
            float2 b0 = hb[s], b1 = hb[s + 1], b2 = (skipB >> j) & 1 ? hb[s + 2] : kZero;
            if (!p.grad) { b0 = s == S - 1 ? kOne : kZero; b1 = s + 1 == S - 1 ? kOne : kZero; b2 = kZero; }
            acc = xf_add(acc, xf_mul(ha[s], xf_sum3_mul(b0, b1, b2, kOne)));
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) acc = xf_add(acc, shfl2(acc, lane ^ o2));
        lp2 = acc;
    }
    const float inv_pm = 1.f / lp2.x;
    const int pe = xf_e(lp2);
    if (FWD && lane == 0) {     // the rows are un-normalised (y' = exp(x)): p = p' / prod_t Z_t
        p.loss[b] = (float)(-(xf_log2(lp2) - (hand_lz[0] + hand_lz[1])) * 0.6931471805599453);
        p.status[b] = CTCASR_CTC_OK;
    }
    if (!p.grad) return;
    // the other warp's rows of my next kPF frames (their logits are already in flight)
    for (int i = own_first; i < own_first + kPF && i < nsteps; ++i) {
        const float4 *src = reinterpret_cast<const float4 *>(rows_b + (size_t)frame_of(i) * RS);
        const uint32_t dst = my_row_s + (uint32_t)(i % kPF) * (uint32_t)SLOT;
#pragma unroll
        for (int q = 0; q < SPL / 2; ++q) cp_async16(dst + q * 16, src + q);
    }
    cp_async_commit();
    cp_async_wait<0>();
    // ---- second half: posteriors from my row and the other warp's row of the frame, gradient row ----------------
    // Per-class sums of the posteriors in fixed point (2^-24; integer adds commute: any order gives the same bits): every
    // label state adds its value to its class's word with a shared-memory reduction (fire and forget, no dependent loads),
    // the blank states go through one warp reduction.  The gradient row of a frame is written ONE ITERATION LATER, when its
    // reductions have long landed: nothing of the class sums sits on the recursion's path.  (A per-class gather over
    // position lists cost a serial chain of dependent shared-memory loads as long as the most frequent label: 37 % of this
    // loop's stall samples.)
    unsigned *csum = postrow;                               // [2 buffers][32 classes]
    csum[lane] = 0u; csum[32 + lane] = 0u;
    __syncwarp();
    float pyrel = 0.f, pzs = 1.f;
    unsigned ppb = 0u;
    float *gptr = grad_b + (size_t)frame_of(own_first) * gstride + (lane < V ? lane : 0);
    auto finish = [&](int n) {                              // gradient row of step n (called in step order) from its class sums
        unsigned *pc = csum + (n & 1) * 32;
        if (lane < V) {
            const unsigned q = lane == blank ? ppb : pc[lane];
            *gptr = (__fdividef(pyrel, pzs) - (float)q * (1.f / 16777216.f)) * p.grad_scale;
        }
        gptr += lstep;
        pc[lane] = 0u;
    };
    for (int n = own_first; n < nsteps; ++n) {
        step(true);
        unsigned *cs = csum + (n & 1) * 32;
        unsigned pblank = 0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const float2 ot = (j & 1) ? make_float2(o[j / 2].z, o[j / 2].w) : make_float2(o[j / 2].x, o[j / 2].y);
            // alpha * beta / (y * p): both rows include y_t; a zero operand has exponent ~kZE and clamps the factor to 0
            const float post = __fdividef(a[j].x * ot.x, ys[j].x) * inv_pm * pow2_clamp(xf_e(a[j]) + xf_e(ot) - xf_e(ys[j]) - pe);
            const unsigned q = __float2uint_rn(fminf(post, 1.f) * 16777216.f);
            if (j & 1) { if ((validm >> j) & 1) atomicAdd(cs + cls[j], q); }
            else pblank += q;
        }
        pblank = __reduce_add_sync(0xffffffffu, pblank);        // <= 2^24 in total (the posteriors of a frame sum to 1)
        if (n > own_first) finish(n - 1);
        __syncwarp();                                           // my reductions before the next iteration's reads, its zeroing before my next adds
        pyrel = yrel; pzs = zs; ppb = pblank;
    }
    if (nsteps > own_first) { __syncwarp(); finish(nsteps - 1); }
}

template <int SPL>
__global__ void __launch_bounds__(64)
ctc_warp_kernel(const Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, blank = p.blank;
    constexpr int RS = 32 * SPL;
    const WarpSmem L(p.Lmax, V, RS);
    int *lab = reinterpret_cast<int *>(smem_raw + L.lab);
    float2 *hand = reinterpret_cast<float2 *>(smem_raw + L.hand);             // [2][RS + 4]
    int *flags = reinterpret_cast<int *>(smem_raw + L.total - 32);

    const int Tb = p.seq_len[b], Ln = p.label_len[b];
    float *grad_b = p.grad ? p.grad + (size_t)b * V : nullptr;                 // row t at + t*B*V
    const size_t gstride = (size_t)p.B * V;

    // ---- setup: labels, validation (as in the block kernel) ------------------------------------
    if (tid < 2) flags[tid] = 0;
    __syncthreads();
    int status = CTCASR_CTC_OK;
    if (Tb > p.T || Tb < 0 || Ln < 0 || Ln > p.Lmax) status = CTCASR_CTC_BAD_LENGTH;
    if (status == CTCASR_CTC_OK) {
        int bad = 0, rep = 0;
        for (int i = tid; i < Ln; i += 64) {
            const int v = p.labels[(size_t)b * p.lstride + i];
            lab[i] = v;
            if (v < 0 || v >= V || v == blank) bad = 1;
            if (i > 0 && v == p.labels[(size_t)b * p.lstride + i - 1]) ++rep;
        }
        if (bad) atomicOr(&flags[0], 1);
        if (rep) atomicAdd(&flags[1], rep);
    }
    __syncthreads();
    if (status == CTCASR_CTC_OK) {
        if (flags[0]) status = CTCASR_CTC_BAD_LABEL;
        else if (Tb < Ln + flags[1]) status = CTCASR_CTC_INFEASIBLE;
    }
    if (grad_b) {       // gradient rows the recursion never touches are zero (t >= T_b, or the whole utterance)
        const int t0 = (status == CTCASR_CTC_OK) ? Tb : 0;
        for (int i = tid; i < (p.T - t0) * V; i += 64)
            grad_b[(size_t)(t0 + i / V) * gstride + i % V] = 0.f;
    }
    if (status != CTCASR_CTC_OK || Tb == 0) {
        if (tid == 0) { p.status[b] = status; p.loss[b] = status == CTCASR_CTC_OK ? 0.f : INFINITY; }
        return;
    }
    for (int i = tid; i < 2 * (RS + 4); i += 64) hand[i] = xf_make(1.f, kZE);        // zero of the warp kernel's arithmetic
    __syncthreads();
    if (warp == 0) ctc_sweep<SPL, true>(p, smem_raw, L, b, lane, Tb, Ln);
    else ctc_sweep<SPL, false>(p, smem_raw, L, b, lane, Tb, Ln);
}

struct Plan { int CH, RS, GT, NCH, VP, SPT, SPL; size_t smem, ws_total; };

static int make_plan(int T, int B, int V, int Lmax, Plan *pl)
{
    const int S = 2 * Lmax + 1;
    if (V < 1 || V > 128) return fail(CTCASR_ERR_UNSUPPORTED, "ctc: num_classes %d > 128", V);
    // the warp-shuffle kernel: SPL lattice states per lane (0 = not eligible: the block kernel below)
    static const int force_block = getenv("CTCASR_CTC_BLOCK") != nullptr;
    pl->SPL = (V <= 32 && S <= 384 && !force_block) ? (S <= 64 ? 2 : (S <= 192 ? 6 : 12)) : 0;
    // Many utterances (more CTAs than SMs): the kernel is bound by instruction issue, so two lattice states
    // per thread — half the warps, the per-step addressing / barrier / loop overhead shared by two states,
    // and 64 registers per thread at 4 CTAs per SM instead of 32 with spills.  Few utterances: one state per
    // thread keeps the per-step latency lowest.
    pl->SPT = 1;
    int GT = (S + 31) / 32 * 32;
    if (B >= 148 && S > 64 && S <= 256) { pl->SPT = 2; GT = ((S + 1) / 2 + 31) / 32 * 32; }
    if (GT > kMaxGT) GT = kMaxGT;
    if (GT < 64) GT = 64;
    if (S > kMaxSPT * GT) return fail(CTCASR_ERR_UNSUPPORTED, "ctc: label length %d too long", Lmax);
    pl->GT = GT;
    pl->RS = (S + 2 + 31) / 32 * 32;
    pl->VP = V | 1;                 // odd row stride: conflict-free lane-per-frame stores
    // chunk length: lattice rows are (hi, lo) pairs, 2 x CH x RS x 8 B of shared memory; many utterances
    // favour a short chunk (4 CTAs per SM), few a long one (less per-chunk overhead)
    pl->CH = pl->RS <= 448 ? (B >= 296 && pl->RS <= 224 ? 8 : 16) : 8;
    pl->NCH = T > 0 ? (T + pl->CH - 1) / pl->CH : 1;
    pl->smem = SmemLayout(Lmax, V, pl->RS, pl->CH, pl->VP).total;
    pl->ws_total = align_up((size_t)B * pl->NCH * pl->RS * sizeof(float2), 256);
    if (pl->SPL) {
        const size_t rows = align_up((size_t)B * (T > 0 ? T : 1) * 32 * pl->SPL * sizeof(float2), 256);
        if (rows > pl->ws_total) pl->ws_total = rows;
    }
    return CTCASR_OK;
}

}  // namespace ctc
}  // namespace ctcasr

using namespace ctcasr;

extern "C" size_t ctcasr_ctc_workspace_bytes(int T, int B, int V, int max_label_len)
{
    ctc::Plan pl{};
    if (ctc::make_plan(T, B, V, max_label_len, &pl) != CTCASR_OK) return 0;
    return pl.ws_total;
}

extern "C" int ctcasr_ctc_loss(const float *logits, int T, int B, int V, int blank,
                               const int32_t *labels, int label_stride, const int32_t *label_len,
                               const int32_t *seq_len, float *loss, float *grad, float grad_scale,
                               int32_t *status, int max_label_len, void *ws, size_t ws_bytes, void *stream)
{
    CTCASR_REQUIRE(logits && labels && label_len && seq_len && loss && status, "ctc: null pointer");
    CTCASR_REQUIRE(T >= 0 && B >= 1 && blank >= 0 && blank < V && max_label_len >= 0 && label_stride >= 1,
                   "ctc: bad dims T=%d B=%d V=%d blank=%d", T, B, V, blank);
    ctc::Plan pl{};
    int rc = ctc::make_plan(T, B, V, max_label_len, &pl);
    if (rc != CTCASR_OK) return rc;
    if (ws_bytes < pl.ws_total || !ws) return fail(CTCASR_ERR_WORKSPACE, "ctc: workspace %zu < %zu", ws_bytes, pl.ws_total);
    ctc::Params p;
    p.logits = logits; p.T = T; p.B = B; p.V = V; p.blank = blank;
    p.labels = labels; p.lstride = label_stride; p.label_len = label_len; p.seq_len = seq_len;
    p.loss = loss; p.grad = grad; p.grad_scale = grad_scale; p.status = status;
    p.ckpt = reinterpret_cast<float2 *>(ws);
    p.rows = reinterpret_cast<float2 *>(ws);
    p.CH = pl.CH; p.RS = pl.RS; p.GT = pl.GT; p.NCH = pl.NCH; p.Lmax = max_label_len; p.VP = pl.VP;
    // three instantiations: one state per thread at high occupancy (the common case, S <= 224),
    // one state per thread up to S = 448, and the generic strided variant for longer labels
    const int NT = 2 * pl.GT + ctc::kHelper;
    const int S = 2 * max_label_len + 1;
    auto launch = [&](auto kernel, size_t &smem_set) -> int {
        if (pl.smem > smem_set) {
            CTCASR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            smem_set = pl.smem;
        }
        ProfScope prof(PROF_CTC, (cudaStream_t)stream);
        kernel<<<B, NT, pl.smem, (cudaStream_t)stream>>>(p);
        CTCASR_LAUNCH_CHECK();
        return CTCASR_OK;
    };
    static size_t set0 = 0, set1 = 0, set2 = 0, set3 = 0;
    if (pl.SPL) {       // two warps per utterance, alpha and beta walking towards each other
        const size_t wsmem = ctc::WarpSmem(max_label_len, V, 32 * pl.SPL).total;
        static size_t wset[3] = {0, 0, 0};
        auto wlaunch = [&](auto kernel, size_t &smem_set) -> int {
            if (wsmem > 48 * 1024 && wsmem > smem_set) {
                CTCASR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
                smem_set = wsmem;
            }
            ProfScope prof(PROF_CTC, (cudaStream_t)stream);
            kernel<<<B, 64, wsmem, (cudaStream_t)stream>>>(p);
            CTCASR_LAUNCH_CHECK();
            return CTCASR_OK;
        };
        if (pl.SPL == 2) return wlaunch(ctc::ctc_warp_kernel<2>, wset[0]);
        if (pl.SPL == 6) return wlaunch(ctc::ctc_warp_kernel<6>, wset[1]);
        return wlaunch(ctc::ctc_warp_kernel<12>, wset[2]);
    }
    if (pl.SPT == 2 && S <= 2 * pl.GT && NT <= 288) return launch(ctc::ctc_loss_kernel<2, 288, 4>, set3);
    if (S <= pl.GT && NT <= 480) return launch(ctc::ctc_loss_kernel<1, 480, 4>, set0);
    if (S <= pl.GT) return launch(ctc::ctc_loss_kernel<1, 2 * ctc::kMaxGT + ctc::kHelper, 1>, set1);
    return launch(ctc::ctc_loss_kernel<ctc::kMaxSPT, 2 * ctc::kMaxGT + ctc::kHelper, 1>, set2);
}

extern "C" int ctcasr_ctc_loss_host(const float *logits, int T, int B, int V, int blank,
                                    const int32_t *labels, int label_stride, const int32_t *label_len,
                                    const int32_t *seq_len, float *loss, float *grad, float grad_scale,
                                    int32_t *status)
{
    CTCASR_REQUIRE(logits && labels && label_len && seq_len && loss && status, "ctc_host: null pointer");
    int lmax = 0;
    for (int b = 0; b < B; ++b) lmax = label_len[b] > lmax ? label_len[b] : lmax;
    if (lmax > label_stride) lmax = label_stride;
    const size_t nlog = (size_t)T * B * V * sizeof(float), nlab = (size_t)B * label_stride * sizeof(int32_t);
    const size_t wsb = ctcasr_ctc_workspace_bytes(T, B, V, lmax);
    if (wsb == 0) return CTCASR_ERR_UNSUPPORTED;
    char *d = nullptr;
    size_t off_logits = 0, off_grad = align_up(nlog, 256), off_lab = off_grad + align_up(nlog, 256);
    size_t off_ll = off_lab + align_up(nlab, 256), off_sl = off_ll + align_up(B * 4, 256);
    size_t off_loss = off_sl + align_up(B * 4, 256), off_st = off_loss + align_up(B * 4, 256);
    size_t off_ws = off_st + align_up(B * 4, 256), total = off_ws + wsb;
    CTCASR_CUDA_CHECK(cudaMalloc(&d, total));
    cudaStream_t s = 0;
    int rc = CTCASR_OK;
    cudaError_t e = cudaSuccess;
    e = cudaMemcpyAsync(d + off_logits, logits, nlog, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + off_lab, labels, nlab, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + off_ll, label_len, B * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + off_sl, seq_len, B * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess)
        rc = ctcasr_ctc_loss((float *)(d + off_logits), T, B, V, blank, (int32_t *)(d + off_lab), label_stride,
                             (int32_t *)(d + off_ll), (int32_t *)(d + off_sl), (float *)(d + off_loss),
                             grad ? (float *)(d + off_grad) : nullptr, grad_scale, (int32_t *)(d + off_st), lmax,
                             d + off_ws, wsb, s);
    if (e == cudaSuccess && rc == CTCASR_OK) e = cudaMemcpyAsync(loss, d + off_loss, B * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && rc == CTCASR_OK) e = cudaMemcpyAsync(status, d + off_st, B * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && rc == CTCASR_OK && grad) e = cudaMemcpyAsync(grad, d + off_grad, nlog, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (e != cudaSuccess) return fail(CTCASR_ERR_CUDA, "ctc_host: %s", cudaGetErrorString(e));
    return rc;
}

extern "C" int ctcasr_greedy_decode(const float *logits, int T, int B, int V, int blank,
                                    const int32_t *seq_len, int32_t *out_ids, int32_t *out_len, void *stream)
{
    CTCASR_REQUIRE(logits && seq_len && out_ids && out_len && T >= 0 && B >= 1 && V >= 1, "greedy_decode: bad args");
    const int wpb = 4;
    ctc::greedy_decode_kernel<<<ceil_div(B, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(logits, T, B, V, blank, seq_len,
                                                                                       out_ids, out_len);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

extern "C" int ctcasr_edit_distance(const int32_t *hyp, int hyp_stride, const int32_t *hyp_len,
                                    const int32_t *truth, int truth_stride, const int32_t *truth_len,
                                    int B, int max_hyp_len, int normalize, float *out, void *stream)
{
    CTCASR_REQUIRE(hyp && hyp_len && truth && truth_len && out && B >= 1 && max_hyp_len >= 0, "edit_distance: bad args");
    const size_t smem = (size_t)3 * (max_hyp_len + 2) * sizeof(int);
    if (smem > 200 * 1024) return fail(CTCASR_ERR_UNSUPPORTED, "edit_distance: hypothesis too long (%d)", max_hyp_len);
    static size_t smem_set = 0;
    if (smem > smem_set && smem > 48 * 1024) {
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(ctc::edit_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    ctc::edit_distance_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(hyp, hyp_stride, hyp_len, truth, truth_stride, truth_len,
                                                                    normalize, out);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}
