// conv.cu — the 2-D convolution layers of the 'ds2' front-end (asr/util/tf_contrib.py:64-146):
// tf.layers.conv2d(padding='SAME', activation=relu) + tf.minimum(., relu_cutoff), forward and backward.
//
// Formulation: a convolution layer is a dense layer over the patch matrix.  `im2col_kernel` gathers,
// for every output position (to, b, fo), its kt x kf x C input patch into one row of `col`
// [To*B*Fo, Kp] (HBM-bound gather, coalesced along the patch: a (jf, c) run is contiguous in x);
// the product with the kernel [Kp, N] (TF's HWIO layout flattened: row = (it*kf + jf)*C + c) then
// runs on the tcgen05 GEMM with the bias + clipped-ReLU epilogue of the dense layers.  Backward:
// dW = col^T dz and dcol = dz W^T on the same GEMM, and `col2im_kernel` folds dcol back onto the
// input as a GATHER (every input element sums the taps that touched it, fixed order: deterministic,
// no atomics).  The patch matrix is re-built in the backward pass instead of being kept (9.5 GB for
// the second layer at B=32 x 10 s): one buffer serves col and then dcol.
//
// Layout: activations [T, B, F, pitch] time-major (row = (t*B + b)*F + f), `pitch` >= channels so
// that the GEMM's N (>= 64, multiple of 8) can be wider than the layer's filter count: the pad
// channels of y are exact zeros (zero kernel columns and bias) and are skipped by the next im2col.
#include "gemm.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace ctcasr {
namespace conv {

struct Geom {
    int T, B, F, C, xpitch;
    int kt, kf, st, sf;
    int To, Fo, pt, pf;
    int K, Kp;
};

static void same_pad(int in, int k, int s, int *out, int *pad0)
{
    *out = (in + s - 1) / s;
    int total = (*out - 1) * s + k - in;
    if (total < 0) total = 0;
    *pad0 = total / 2;          // TF 'SAME': the odd unit goes after
}

static int make_geom(int T, int B, int F, int C, int xpitch, int kt, int kf, int st, int sf, Geom *g)
{
    CTCASR_REQUIRE(T >= 1 && B >= 1 && F >= 1 && C >= 1 && xpitch >= C, "conv2d: bad dims T=%d B=%d F=%d C=%d pitch=%d", T, B, F, C, xpitch);
    CTCASR_REQUIRE(kt >= 1 && kf >= 1 && st >= 1 && sf >= 1, "conv2d: bad kernel %dx%d / stride %dx%d", kt, kf, st, sf);
    g->T = T; g->B = B; g->F = F; g->C = C; g->xpitch = xpitch;
    g->kt = kt; g->kf = kf; g->st = st; g->sf = sf;
    same_pad(T, kt, st, &g->To, &g->pt);
    same_pad(F, kf, sf, &g->Fo, &g->pf);
    g->K = kt * kf * C;
    g->Kp = (g->K + 7) / 8 * 8;
    return CTCASR_OK;
}

// exact n / d for n, d < 2^16 (one IMAD.HI): mul = ceil(2^32 / d)
struct FastDiv { uint32_t mul, d; };
static FastDiv make_fastdiv(int d)
{
    FastDiv f;
    f.d = (uint32_t)d;
    f.mul = d <= 1 ? 0u : (uint32_t)(0xffffffffull / (uint64_t)d) + 1u;
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv f) { return f.d <= 1 ? n : __umulhi(n, f.mul); }

// One warp per output position (grid-stride), lanes along the patch in units of four columns: one 16-B
// store per lane and iteration.  VEC (C and the input pitch multiples of 4): the four columns are four
// channels of one tap -> one 16-B load; otherwise four scalar gathers.
// NP = 0: fp32 patch matrix `col`.  NP = 1..3: the patch matrix straight as the NP bf16 pieces the tcgen05 GEMM reads
// (piece 1 = bf16(v), piece 2 = bf16(v - piece 1), ...; `pieces` [NP][rows][Kp]) — the fp32 matrix (9.5 GB for the
// second layer at B = 32 x 10 s) is then never written, re-read and split.
template <bool VEC, int NP>
__global__ void __launch_bounds__(256) im2col_kernel(const float *__restrict__ x, float *__restrict__ col, __nv_bfloat16 *__restrict__ pieces,
                                                     const Geom g, size_t rows, const FastDiv dC, const FastDiv dkf)
{
    const int lane = threadIdx.x & 31;
    const size_t warp0 = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    const int kq = g.Kp >> 2;
    for (size_t r = warp0; r < rows; r += nwarps) {
        const int fo = (int)(r % g.Fo);
        const size_t q = r / g.Fo;
        const int b = (int)(q % g.B), to = (int)(q / g.B);
        const int t0 = to * g.st - g.pt, f0 = fo * g.sf - g.pf;
        float4 *crow = reinterpret_cast<float4 *>(col + r * g.Kp);
        for (int j = lane; j < kq; j += 32) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t k = 4u * j;
            if (VEC) {
                if (k < (uint32_t)g.K) {
                    const uint32_t q2 = fdiv(k, dC), c = k - q2 * g.C;
                    const uint32_t it = fdiv(q2, dkf), jf = q2 - it * g.kf;
                    const int t = t0 + (int)it, f = f0 + (int)jf;
                    if (t >= 0 && t < g.T && f >= 0 && f < g.F) {
                        const float4 w = __ldg(reinterpret_cast<const float4 *>(x + (((size_t)t * g.B + b) * g.F + f) * g.xpitch + c));
                        v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t ke = k + e;
                    if (ke < (uint32_t)g.K) {
                        const uint32_t q2 = fdiv(ke, dC), c = ke - q2 * g.C;
                        const uint32_t it = fdiv(q2, dkf), jf = q2 - it * g.kf;
                        const int t = t0 + (int)it, f = f0 + (int)jf;
                        if (t >= 0 && t < g.T && f >= 0 && f < g.F) v[e] = __ldg(x + (((size_t)t * g.B + b) * g.F + f) * g.xpitch + c);
                    }
                }
            }
            if (NP == 0) {
                crow[j] = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int pc = 0; pc < (NP ? NP : 1); ++pc) {
                    __nv_bfloat16 h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { h[e] = __float2bfloat16_rn(v[e]); v[e] -= __bfloat162float(h[e]); }
                    uint2 pk;
                    pk.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
                    pk.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
                    *reinterpret_cast<uint2 *>(pieces + ((size_t)pc * rows + r) * g.Kp + 4 * j) = pk;
                }
            }
        }
    }
}

// One thread per W consecutive channels of an input pixel (t, b, f) over the whole pitch (pad channels
// get 0): it sums, in a fixed order, the taps (it, jf) whose output position (to, fo) exists.
template <int W>
__global__ void __launch_bounds__(256) col2im_kernel(const float *__restrict__ dcol, float *__restrict__ dx, const Geom g, size_t total)
{
    const int cq = g.xpitch / W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cq) * W;
        const size_t r = i / cq;
        float s[W];
#pragma unroll
        for (int e = 0; e < W; ++e) s[e] = 0.f;
        if (c < g.C) {
            const int f = (int)(r % g.F);
            const size_t q = r / g.F;
            const int b = (int)(q % g.B), t = (int)(q / g.B);
            const int tn0 = t + g.pt, fn0 = f + g.pf;
            for (int it = tn0 % g.st; it < g.kt && it <= tn0; it += g.st) {
                const int to = (tn0 - it) / g.st;
                if (to >= g.To) continue;
                const int jf0 = fn0 % g.sf;
                int fo = (fn0 - jf0) / g.sf;
                const float *base = dcol + ((size_t)to * g.B + b) * g.Fo * g.Kp + (size_t)it * g.kf * g.C + c;
                for (int jf = jf0; jf < g.kf && fo >= 0; jf += g.sf, --fo) {
                    if (fo >= g.Fo) continue;
                    const float *p = base + (size_t)fo * g.Kp + jf * g.C;
                    if (W == 4) {
                        const float4 w = __ldg(reinterpret_cast<const float4 *>(p));
                        s[0] += w.x; s[1] += w.y; s[2] += w.z; s[3] += w.w;
                    } else {
                        s[0] += __ldg(p);
                    }
                }
            }
        }
        if (W == 4) *reinterpret_cast<float4 *>(dx + r * g.xpitch + c) = make_float4(s[0], s[1], s[2], s[3]);
        else dx[r * g.xpitch + c] = s[0];
    }
}

template <int NP>
static int launch_im2col_np(const float *x, float *col, __nv_bfloat16 *pieces, const Geom &g, size_t rows, cudaStream_t stream)
{
    const size_t blocks = (rows + 7) / 8;
    const int grid = (int)(blocks < (size_t)148 * 16 ? blocks : (size_t)148 * 16);
    const FastDiv dC = make_fastdiv(g.C), dkf = make_fastdiv(g.kf);
    if (g.C % 4 == 0 && g.xpitch % 4 == 0 && ((uintptr_t)x & 15) == 0)
        im2col_kernel<true, NP><<<grid, 256, 0, stream>>>(x, col, pieces, g, rows, dC, dkf);
    else
        im2col_kernel<false, NP><<<grid, 256, 0, stream>>>(x, col, pieces, g, rows, dC, dkf);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// The patch matrix as the GEMM operand `col` [rows, Kp].  In the bf16 compute modes (inside an open split scope) the
// pieces are produced directly and registered under the address of `col`, which is then only a key: the GEMM finds
// the split in the cache.  np = pieces the GEMM that follows will ask for (3: six-product forward through the ReLU
// kink, 2: three-product, 1: compute = 'bf16'); 0 = the fp32 matrix.
static int launch_im2col(const float *x, float *col, const Geom &g, size_t rows, int np, cudaStream_t stream)
{
    CTCASR_REQUIRE(g.Kp < 65536, "conv2d: patch of %d elements", g.K);
    if (np == 0) return launch_im2col_np<0>(x, col, nullptr, g, rows, stream);
    __nv_bfloat16 *pieces = nullptr;
    if (int rc = split_reserve(col, (int)rows, g.Kp, g.Kp, np, &pieces)) return rc;
    ProfScope prof_split(PROF_SPLIT, stream);
    if (np == 1) return launch_im2col_np<1>(x, col, pieces, g, rows, stream);
    if (np == 2) return launch_im2col_np<2>(x, col, pieces, g, rows, stream);
    return launch_im2col_np<3>(x, col, pieces, g, rows, stream);
}
// pieces per operand of the tcgen05 GEMM this product will run as (0: the fp32 / tf32 / SIMT paths read the fp32 matrix)
static int gemm_pieces(const GemmArgs &a, int compute)
{
    if (!gemm_tc_eligible(a)) return 0;
    if (compute == CTCASR_COMPUTE_BF16) return 1;
    if (compute == CTCASR_COMPUTE_BF16X3) return a.precise ? 3 : 2;
    return 0;
}

static int launch_col2im(const float *dcol, float *dx, const Geom &g, cudaStream_t stream)
{
    const bool vec = g.C % 4 == 0 && g.xpitch % 4 == 0 && ((uintptr_t)dx & 15) == 0;
    const size_t total = (size_t)g.T * g.B * g.F * (g.xpitch / (vec ? 4 : 1));
    const size_t blocks = (total + 255) / 256;
    const int grid = (int)(blocks < (size_t)148 * 32 ? (blocks ? blocks : 1) : (size_t)148 * 32);
    if (vec) col2im_kernel<4><<<grid, 256, 0, stream>>>(dcol, dx, g, total);
    else col2im_kernel<1><<<grid, 256, 0, stream>>>(dcol, dx, g, total);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

}  // namespace conv
}  // namespace ctcasr

using namespace ctcasr;

extern "C" int ctcasr_conv2d_out_dims(int T, int F, int kt, int kf, int st, int sf, int *To, int *Fo)
{
    CTCASR_REQUIRE(To && Fo && T >= 1 && F >= 1 && kt >= 1 && kf >= 1 && st >= 1 && sf >= 1, "conv2d_out_dims: bad args");
    int p;
    conv::same_pad(T, kt, st, To, &p);
    conv::same_pad(F, kf, sf, Fo, &p);
    return CTCASR_OK;
}

extern "C" size_t ctcasr_conv2d_workspace_bytes(int T, int B, int F, int C, int kt, int kf, int st, int sf)
{
    conv::Geom g{};
    if (conv::make_geom(T, B, F, C, C, kt, kf, st, sf, &g) != CTCASR_OK) return 0;
    return align_up((size_t)g.To * B * g.Fo * g.Kp * sizeof(float), 256);
}

extern "C" int ctcasr_conv2d_fwd(const float *x, int x_pitch, const float *w, const float *bias, float *y,
                                 int T, int B, int F, int C, int kt, int kf, int st, int sf, int N,
                                 int act, float cutoff, float drop_rate, uint32_t seed, int compute,
                                 void *ws, size_t ws_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(x && w && y && N >= 1, "conv2d_fwd: bad args");
    CTCASR_REQUIRE(drop_rate >= 0.f && drop_rate < 1.f, "conv2d_fwd: drop_rate %g", drop_rate);
    conv::Geom g{};
    if (int rc = conv::make_geom(T, B, F, C, x_pitch, kt, kf, st, sf, &g)) return rc;
    const size_t rows = (size_t)g.To * B * g.Fo;
    CTCASR_REQUIRE(rows <= 0x7fffffff, "conv2d_fwd: %zu output positions", rows);
    const size_t need = align_up(rows * g.Kp * sizeof(float), 256);
    if (!ws || ws_bytes < need) return fail(CTCASR_ERR_WORKSPACE, "conv2d_fwd: workspace %zu < %zu", ws_bytes, need);
    if (conv_tc_eligible(compute, T, B, F, C, kt, kf, st, sf, N) && ((uintptr_t)x & 15) == 0 && x_pitch % 4 == 0 && N % 8 == 0) {
        // implicit GEMM (conv_tc.cu): the TMA unit gathers the taps from the layer input, no patch matrix
        const int np = compute == CTCASR_COMPUTE_BF16 ? 1 : (act != 0 ? 3 : 2);
        const size_t elems[2] = {(size_t)T * B * F * C, (size_t)g.Kp * N};
        SplitScope scope;
        if (int rc = split_scope_begin(compute, elems, 2)) return rc;
        scope.open = true;
        return conv_tc_fwd(x, x_pitch, w, N, bias, y, N, T, B, F, C, kt, kf, st, sf, g.To, g.Fo, g.pt, g.pf, N, np, act, cutoff,
                           drop_rate, seed, stream);
    }
    if (int rcs = gemm_scratch_check(compute, 1, (int)rows, N, g.Kp)) return rcs;
    float *col = reinterpret_cast<float *>(ws);
    GemmArgs a;
    a.A[0] = col; a.B[0] = w; a.C[0] = y; a.M = (int)rows; a.N = N; a.K = g.Kp; a.lda = g.Kp; a.ldb = N; a.ldc = N;
    a.epi.mode = EPI_BIAS_ACT; a.epi.bias = bias; a.epi.act = act; a.epi.cutoff = cutoff;
    a.epi.drop_rate = drop_rate; a.epi.seed = seed;
    a.precise = act != 0;
    const int np = conv::gemm_pieces(a, compute);
    SplitScope scope;
    if (np) {
        const size_t elems[2] = {rows * (size_t)g.Kp, (size_t)g.Kp * N};
        if (int rc = split_scope_begin(compute, elems, 2)) return rc;
        scope.open = true;
    }
    if (int rc = conv::launch_im2col(x, col, g, rows, np, stream)) return rc;
    return gemm(a, compute, stream);
}

extern "C" int ctcasr_conv2d_bwd(const float *x, int x_pitch, const float *w, const float *y, float *dy,
                                 float *dx, float *dw, float *db,
                                 int T, int B, int F, int C, int kt, int kf, int st, int sf, int N,
                                 int act, float cutoff, float drop_rate, uint32_t seed, int compute,
                                 void *ws, size_t ws_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(x && w && dy && dw && db && N >= 1, "conv2d_bwd: bad args");
    CTCASR_REQUIRE(drop_rate >= 0.f && drop_rate < 1.f, "conv2d_bwd: drop_rate %g", drop_rate);
    CTCASR_REQUIRE(act == 0 || y, "conv2d_bwd: activation mask needs the forward output y");
    conv::Geom g{};
    if (int rc = conv::make_geom(T, B, F, C, x_pitch, kt, kf, st, sf, &g)) return rc;
    const size_t rows = (size_t)g.To * B * g.Fo;
    CTCASR_REQUIRE(rows <= 0x7fffffff, "conv2d_bwd: %zu output positions", rows);
    const size_t need = align_up(rows * g.Kp * sizeof(float), 256);
    if (!ws || ws_bytes < need) return fail(CTCASR_ERR_WORKSPACE, "conv2d_bwd: workspace %zu < %zu", ws_bytes, need);
    static const bool implicit_wgrad = !(getenv("CTCASR_CONV_IMPLICIT_WGRAD") && atoi(getenv("CTCASR_CONV_IMPLICIT_WGRAD")) == 0);
    const bool use_implicit = implicit_wgrad && conv_tc_eligible(compute, T, B, F, C, kt, kf, st, sf, N) && ((uintptr_t)x & 15) == 0 &&
                              x_pitch % 4 == 0 && g.K == g.Kp;
    if (!use_implicit) if (int rcs = gemm_scratch_check(compute, 1, g.Kp, N, (int)rows)) return rcs;
    if (dx) if (int rcs = gemm_scratch_check(compute, 1, (int)rows, g.Kp, N)) return rcs;
    float *col = reinterpret_cast<float *>(ws);
    int rc;
    const bool dgrad_implicit = use_implicit && dx && conv_tc_dgrad_eligible(C, kf, st, sf, x_pitch);
    // all consumers of dz read its bf16 pieces: mask, bias gradient and pieces in one pass, no fp32 dz (pointwise.cu)
    GemmArgs dcol_args;         // dcol[rows,Kp] = dz W^T of the patch-matrix input gradient
    dcol_args.A[0] = dy; dcol_args.B[0] = w; dcol_args.C[0] = col; dcol_args.tb = 1; dcol_args.M = (int)rows; dcol_args.N = g.Kp; dcol_args.K = N;
    dcol_args.lda = N; dcol_args.ldb = N; dcol_args.ldc = g.Kp;
    const bool fused_dz = use_implicit && N <= 16384;          // (independent of dx: the bias gradient's summation order must not depend on it)
    const bool fp32_dz = !fused_dz || (dx && !dgrad_implicit && !gemm_tc_eligible(dcol_args));    // someone reads dz itself
    if (!fused_dz) {
        rc = mask_inplace(dy, y, rows, N, act, cutoff, drop_rate, seed, stream);       // dy -> dz
        if (rc != CTCASR_OK) return rc;
        rc = colsum(dy, (int)rows, N, N, db, stream);
        if (rc != CTCASR_OK) return rc;
    }
    if (use_implicit) {
        // implicit GEMMs (conv_tc.cu): dW = col^T dz with the patches gathered by the TMA unit, and dx = conv_transpose(dz, W)
        // without dcol; dz is split once for both
        const int np = compute == CTCASR_COMPUTE_BF16 ? 1 : 2;
        const size_t elems[3] = {(size_t)T * B * F * C, rows * (size_t)N, (size_t)g.Kp * N};
        SplitScope scope;
        if ((rc = split_scope_begin(compute, elems, 3)) != CTCASR_OK) return rc;
        scope.open = true;
        if (fused_dz) {
            rc = mask_colsum_split(dy, y, (int)rows, N, act, cutoff, drop_rate, seed, np, db, stream);
            if (rc != CTCASR_OK) return rc;
            if (fp32_dz && (rc = mask_inplace(dy, y, rows, N, act, cutoff, drop_rate, seed, stream)) != CTCASR_OK) return rc;
        }
        rc = conv_tc_wgrad(x, x_pitch, dy, N, dw, N, T, B, F, C, kt, kf, st, sf, g.To, g.Fo, g.pt, g.pf, np, stream);
        if (rc != CTCASR_OK) return rc;
        if (dgrad_implicit)
            return conv_tc_dgrad(dy, N, w, N, dx, x_pitch, T, B, F, C, kt, kf, st, sf, g.To, g.Fo, g.pt, g.pf, np, stream);
        if (dx) {   // patch-matrix input gradient (time stride > 1), still inside the scope: its GEMM reads dz's pieces
            if ((rc = gemm(dcol_args, compute, stream)) != CTCASR_OK) return rc;
            return conv::launch_col2im(col, dx, g, stream);
        }
        return CTCASR_OK;
    } else {   // dW[Kp,N] = col^T dz   (rows K..Kp of col^T are zero -> the pad rows of dW are zero)
        GemmArgs a;
        a.A[0] = col; a.B[0] = dy; a.C[0] = dw; a.ta = 1; a.M = g.Kp; a.N = N; a.K = (int)rows; a.lda = g.Kp; a.ldb = N; a.ldc = N;
        const int np = conv::gemm_pieces(a, compute);
        SplitScope scope;
        if (np) {
            const size_t elems[2] = {rows * (size_t)g.Kp, rows * (size_t)N};
            if ((rc = split_scope_begin(compute, elems, 2)) != CTCASR_OK) return rc;
            scope.open = true;
        }
        if ((rc = conv::launch_im2col(x, col, g, rows, np, stream)) != CTCASR_OK) return rc;
        rc = gemm(a, compute, stream);
        if (rc != CTCASR_OK) return rc;
    }
    if (dx) {   // dcol[rows,Kp] = dz W^T into the buffer col occupied (stream order: dW has consumed it)
        GemmArgs a;
        a.A[0] = dy; a.B[0] = w; a.C[0] = col; a.tb = 1; a.M = (int)rows; a.N = g.Kp; a.K = N; a.lda = N; a.ldb = N; a.ldc = g.Kp;
        rc = gemm(a, compute, stream);
        if (rc != CTCASR_OK) return rc;
        if ((rc = conv::launch_col2im(col, dx, g, stream)) != CTCASR_OK) return rc;
    }
    return CTCASR_OK;
}
