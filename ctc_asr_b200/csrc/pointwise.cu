// pointwise.cu — HBM-bound helpers around the GEMMs: layout transpose, bias gradient (column
// sums), activation mask, recurrent cell math for the stepwise (SIMT) recurrent path, Adam.
#include "gemm.cuh"
#include "rnn.cuh"

#include <cuda_bf16.h>

namespace ctcasr {

// ---- [A,B,C] -> [B,A,C] ---------------------------------------------------------------------
__global__ void transpose01_kernel(const float *__restrict__ in, float *__restrict__ out, int A, int B, int C)
{
    const size_t total = (size_t)A * B * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t r = i / C;            // output row = b*A + a
        const int a = (int)(r % A), b = (int)(r / A);
        out[i] = in[((size_t)a * B + b) * C + c];
    }
}

// ---- column sums: out[n] = sum_m x[m*ld + n]  (bias gradients) --------------------------------
// Two deterministic stages: the rows are cut into `parts` equal slabs (grid.y), every CTA sums its slab
// for 32 columns, and the slab sums are added in slab order.  One stage (parts = 1) for short matrices.
constexpr int kColsumMaxParts = 64, kColsumMaxN = 16384;
__device__ float g_colsum_partial[kColsumMaxParts * kColsumMaxN];

__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ x, int M, int N, int ld, float *__restrict__ out,
                                                     int rows_per_part, int to_partial)
{
    __shared__ float red[8][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = threadIdx.x >> 5;
    const int m0 = blockIdx.y * rows_per_part, m1 = min(M, m0 + rows_per_part);
    float acc = 0.f;
    if (n < N)
        for (int m = m0 + r; m < m1; m += 8) acc += x[(size_t)m * ld + n];
    red[r][threadIdx.x & 31] = acc;
    __syncthreads();
    if (r == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        if (to_partial) g_colsum_partial[(size_t)blockIdx.y * N + n] = s;
        else out[n] = s;
    }
}

__global__ void colsum_finish_kernel(float *__restrict__ out, int N, int parts)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int p = 0; p < parts; ++p) s += g_colsum_partial[(size_t)p * N + n];
    out[n] = s;
}

int colsum(const float *x, int M, int N, int ld, float *out, cudaStream_t stream)
{
    const int nb = ceil_div(N, 32);
    int parts = ceil_div(2 * 148, nb);
    if (parts > kColsumMaxParts) parts = kColsumMaxParts;
    if (parts > M / 256) parts = M / 256;
    if (parts < 2 || N > kColsumMaxN) {
        colsum_kernel<<<dim3(nb, 1), 256, 0, stream>>>(x, M, N, ld, out, M, 0);
        CTCASR_LAUNCH_CHECK();
        return CTCASR_OK;
    }
    const int rpp = ceil_div(M, parts);
    colsum_kernel<<<dim3(nb, parts), 256, 0, stream>>>(x, M, N, ld, out, rpp, 1);
    CTCASR_LAUNCH_CHECK();
    colsum_finish_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(out, N, parts);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// ---- dz = dy * act'(y) * dropout mask, in place ----------------------------------------------
__global__ void mask_inplace_kernel(float *__restrict__ dy, const float *__restrict__ y, size_t total, int N,
                                    int act, float cutoff, float drop_rate, uint32_t seed)
{
    const float inv_keep = drop_rate > 0.f ? 1.f / (1.f - drop_rate) : 1.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        bool pass = drop_rate > 0.f ? drop_keep(seed, i, drop_rate) : true;
        if (act == 1) { const float v = y[i]; pass = pass && v > 0.f && v < cutoff * inv_keep; }
        dy[i] = pass ? dy[i] * inv_keep : 0.f;
    }
}

int mask_inplace(float *dy, const float *y, size_t M, int N, int act, float cutoff, float drop_rate,
                 uint32_t seed, cudaStream_t stream)
{
    if (act == 0 && drop_rate <= 0.f) return CTCASR_OK;
    const size_t total = M * (size_t)N;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    mask_inplace_kernel<<<blocks, 256, 0, stream>>>(dy, y, total, N, act, cutoff, drop_rate, seed);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// ---- dense / conv backward prologue in one pass: dz = dy * act'(y) * dropout mask, its column sums (the bias gradient)
// and its bf16 pieces for the tcgen05 GEMMs that read dz (dW = x^T dz, dx = dz W^T).  Replaces mask_inplace + colsum + two
// split_bf16 passes (24 B per element of HBM traffic) by one pass of 8 B in + 2 NP B out; the fp32 dz is not written (only
// the GEMMs read it, through the pieces).  Column sums: rows cut into slabs (grid.y), every CTA sums its slab for 64
// columns (8 row phases added in order), the slabs are added in slab order: deterministic.
template <int NP>
__global__ void __launch_bounds__(256) mask_colsum_split_kernel(const float *__restrict__ dy, const float *__restrict__ y, int M, int N,
                                                                int act, float cutoff, float drop_rate, uint32_t seed,
                                                                __nv_bfloat16 *__restrict__ pieces, float *__restrict__ out,
                                                                int rows_per_part, int to_partial)
{
    __shared__ float red[8][66];
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int n = blockIdx.x * 64 + 2 * lane;
    const int m0 = blockIdx.y * rows_per_part, m1 = min(M, m0 + rows_per_part);
    const float inv_keep = drop_rate > 0.f ? 1.f / (1.f - drop_rate) : 1.f;
    const size_t piece = (size_t)M * N;
    float a0 = 0.f, a1 = 0.f;
    if (n < N)
        for (int m = m0 + r; m < m1; m += 8) {
            const size_t i = (size_t)m * N + n;
            const float2 g = *reinterpret_cast<const float2 *>(dy + i);
            bool p0 = drop_rate > 0.f ? drop_keep(seed, i, drop_rate) : true;
            bool p1 = drop_rate > 0.f ? drop_keep(seed, i + 1, drop_rate) : true;
            if (act == 1) {
                const float2 v = *reinterpret_cast<const float2 *>(y + i);
                p0 = p0 && v.x > 0.f && v.x < cutoff * inv_keep;
                p1 = p1 && v.y > 0.f && v.y < cutoff * inv_keep;
            }
            float d0 = p0 ? g.x * inv_keep : 0.f, d1 = p1 ? g.y * inv_keep : 0.f;
            a0 += d0; a1 += d1;
#pragma unroll
            for (int pc = 0; pc < NP; ++pc) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(d0), h1 = __float2bfloat16_rn(d1);
                d0 -= __bfloat162float(h0); d1 -= __bfloat162float(h1);
                *reinterpret_cast<uint32_t *>(pieces + pc * piece + i) =
                    (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            }
        }
    red[r][2 * lane] = a0; red[r][2 * lane + 1] = a1;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int nn = blockIdx.x * 64 + threadIdx.x;
        if (nn < N) {
            float s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s2 += red[i][threadIdx.x];
            if (to_partial) g_colsum_partial[(size_t)blockIdx.y * N + nn] = s2;
            else out[nn] = s2;
        }
    }
}

// dz pieces registered in the open split scope under the address of dy (gemm.cuh, split_reserve); db = column sums of dz
int mask_colsum_split(const float *dy, const float *y, int M, int N, int act, float cutoff, float drop_rate, uint32_t seed,
                      int np, float *db, cudaStream_t stream)
{
    if (N % 8 || N > kColsumMaxN || np < 1 || np > 3) return fail(CTCASR_ERR_UNSUPPORTED, "mask_colsum_split: N = %d, %d pieces", N, np);
    __nv_bfloat16 *pieces = nullptr;
    if (int rc = split_reserve(dy, M, N, N, np, &pieces)) return rc;
    const int nb = ceil_div(N, 64);
    int parts = ceil_div(4 * 148, nb);
    if (parts > kColsumMaxParts) parts = kColsumMaxParts;
    if (parts > M / 64) parts = M / 64;
    if (parts < 1) parts = 1;
    const int rpp = ceil_div(M, parts);
    const int to_partial = parts > 1;
    ProfScope prof(PROF_SPLIT, stream);
    if (np == 1) mask_colsum_split_kernel<1><<<dim3(nb, parts), 256, 0, stream>>>(dy, y, M, N, act, cutoff, drop_rate, seed, pieces, db, rpp, to_partial);
    else if (np == 2) mask_colsum_split_kernel<2><<<dim3(nb, parts), 256, 0, stream>>>(dy, y, M, N, act, cutoff, drop_rate, seed, pieces, db, rpp, to_partial);
    else mask_colsum_split_kernel<3><<<dim3(nb, parts), 256, 0, stream>>>(dy, y, M, N, act, cutoff, drop_rate, seed, pieces, db, rpp, to_partial);
    CTCASR_LAUNCH_CHECK();
    if (to_partial) {
        colsum_finish_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(db, N, parts);
        CTCASR_LAUNCH_CHECK();
    }
    return CTCASR_OK;
}

// ---- the 29-class layer on the tensor cores: its operands widened to 64 columns (zeros), its results narrowed back ----
__global__ void pad_cols64_kernel(const float *__restrict__ src, size_t rows, int n, float *__restrict__ dst)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * 64; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 63);
        dst[i] = c < n ? src[(i >> 6) * n + c] : 0.f;
    }
}
__global__ void compact_cols64_kernel(const float *__restrict__ src, size_t rows, int n, const float *__restrict__ bias, float *__restrict__ dst)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / n;
        const int c = (int)(i - r * n);
        dst[i] = src[r * 64 + c] + (bias ? bias[c] : 0.f);
    }
}
int pad_cols64(const float *src, int rows, int n, float *dst, cudaStream_t stream)
{
    const size_t blocks = ((size_t)rows * 64 + 255) / 256;
    pad_cols64_kernel<<<(int)(blocks < (size_t)148 * 8 ? blocks : (size_t)148 * 8), 256, 0, stream>>>(src, (size_t)rows, n, dst);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}
int compact_cols64(const float *src, int rows, int n, const float *bias, float *dst, cudaStream_t stream)
{
    const size_t blocks = ((size_t)rows * n + 255) / 256;
    compact_cols64_kernel<<<(int)(blocks < (size_t)148 * 8 ? blocks : (size_t)148 * 8), 256, 0, stream>>>(src, (size_t)rows, n, bias, dst);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// ---- y = dropout(x): the keep-mask of the dense epilogue over the flat index (RNN input / output / inter-layer dropout) ----
__global__ void dropout_kernel(const float *__restrict__ x, float *__restrict__ y, size_t total, float drop_rate, uint32_t seed)
{
    const float inv_keep = 1.f / (1.f - drop_rate);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        y[i] = drop_keep(seed, i, drop_rate) ? x[i] * inv_keep : 0.f;
}

// ---- recurrent cell math, stepwise path --------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

// gates: [T*B, 2*G*H] holding z = x Wx + b + h_prev Wh for frame t; replaced by activations.
// grid (ceil(B*H/256), 2 directions); step index `i`: fw frame t = i, bw frame t = T-1-i.
__global__ void rnn_cell_fwd_kernel(RnnStep s, int i)
{
    const int d = blockIdx.y;
    const int t = d == 0 ? i : s.T - 1 - i;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.B * s.H) return;
    const int b = idx / s.H, u = idx % s.H;
    const int H = s.H, GH = s.G * H;
    float *g = s.gates + ((size_t)t * s.B + b) * 2 * GH + (size_t)d * GH;
    float *yo = s.y + ((size_t)t * s.B + b) * 2 * H + (size_t)d * H + u;
    const bool live = !s.use_len || t < s.seq_len[b];
    const int tp = d == 0 ? t - 1 : t + 1;
    const bool has_prev = i > 0;
    if (s.cell == CTCASR_CELL_GRU) {
        float *qt = s.cstate + ((size_t)t * s.B + b) * 2 * H + (size_t)d * H + u;
        if (!live) { *yo = 0.f; *qt = 0.f; g[u] = 0.f; g[H + u] = 0.f; g[2 * H + u] = 0.f; return; }
        const float *rh = s.rh + ((size_t)d * s.B + b) * 3 * H;
        const float rr = has_prev ? rh[u] : 0.f, rz = has_prev ? rh[H + u] : 0.f, rn = has_prev ? rh[2 * H + u] : 0.f;
        const float hp = has_prev ? s.y[((size_t)tp * s.B + b) * 2 * H + (size_t)d * H + u] : 0.f;
        const float q = rn + s.bias_rn[d * H + u];
        const float gr = sigmoidf_(g[u] + rr), gz = sigmoidf_(g[H + u] + rz);
        const float gn = tanhf(g[2 * H + u] + gr * q);
        g[u] = gr; g[H + u] = gz; g[2 * H + u] = gn;
        *qt = q;
        *yo = (1.f - gz) * gn + gz * hp;
        return;
    }
    if (s.cell == CTCASR_CELL_LSTM) {
        float *ct = s.cstate + ((size_t)t * s.B + b) * 2 * H + (size_t)d * H + u;
        const float cp = has_prev ? s.cstate[((size_t)tp * s.B + b) * 2 * H + (size_t)d * H + u] : 0.f;
        if (!live) { *yo = 0.f; *ct = cp; g[u] = 0.f; g[H + u] = 0.f; g[2 * H + u] = 0.f; g[3 * H + u] = 0.f; return; }
        const float gi = sigmoidf_(g[u]), gj = tanhf(g[H + u]);
        const float gf = sigmoidf_(g[2 * H + u] + s.forget_bias), go = sigmoidf_(g[3 * H + u]);
        const float c = gf * cp + gi * gj;
        g[u] = gi; g[H + u] = gj; g[2 * H + u] = gf; g[3 * H + u] = go;
        *ct = c;
        *yo = go * tanhf(c);
    } else {
        if (!live) { *yo = 0.f; g[u] = 0.f; return; }
        const float z = g[u];
        const float h = s.cell == CTCASR_CELL_RNN_TANH ? tanhf(z) : fmaxf(z, 0.f);
        g[u] = h;
        *yo = h;
    }
}

// backward cell step: consumes dy[t], dh_rec (recurrent gradient from the step processed before),
// dc carry; replaces the activations in `gates` by dz.  Same grid / step convention, but the
// steps run in reverse processing order: i = T-1 .. 0.
__global__ void rnn_cell_bwd_kernel(RnnStep s, int i)
{
    const int d = blockIdx.y;
    const int t = d == 0 ? i : s.T - 1 - i;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.B * s.H) return;
    const int b = idx / s.H, u = idx % s.H;
    const int H = s.H, GH = s.G * H;
    float *g = s.gates + ((size_t)t * s.B + b) * 2 * GH + (size_t)d * GH;
    const bool live = !s.use_len || t < s.seq_len[b];
    float *dhr = s.dh_rec + ((size_t)d * s.B + b) * H + u;
    const float dh = s.dy[((size_t)t * s.B + b) * 2 * H + (size_t)d * H + u] + (i == s.T - 1 ? 0.f : *dhr);
    const int tp = d == 0 ? t - 1 : t + 1;
    const bool has_prev = i > 0;
    if (s.cell == CTCASR_CELL_GRU) {
        float *dhd = s.dc_carry + ((size_t)d * s.B + b) * H + u;
        float *zr = s.dzr + ((size_t)t * s.B + b) * 2 * GH + (size_t)d * GH;
        if (!live) { g[u] = 0.f; g[H + u] = 0.f; g[2 * H + u] = 0.f; zr[u] = 0.f; zr[H + u] = 0.f; zr[2 * H + u] = 0.f; *dhd = 0.f; return; }
        const float gr = g[u], gz = g[H + u], gn = g[2 * H + u];
        const float q = s.cstate[((size_t)t * s.B + b) * 2 * H + (size_t)d * H + u];
        const float hp = has_prev ? s.y[((size_t)tp * s.B + b) * 2 * H + (size_t)d * H + u] : 0.f;
        const float dht = dh + (i == s.T - 1 ? 0.f : *dhd);
        const float dn_pre = dht * (1.f - gz) * (1.f - gn * gn);
        const float dz_pre = dht * (hp - gn) * gz * (1.f - gz);
        const float dr_pre = dn_pre * q * gr * (1.f - gr);
        g[u] = dr_pre; g[H + u] = dz_pre; g[2 * H + u] = dn_pre;
        zr[u] = dr_pre; zr[H + u] = dz_pre; zr[2 * H + u] = dn_pre * gr;
        *dhd = dht * gz;
        return;
    }
    if (s.cell == CTCASR_CELL_LSTM) {
        float *dcc = s.dc_carry + ((size_t)d * s.B + b) * H + u;
        if (!live) { g[u] = 0.f; g[H + u] = 0.f; g[2 * H + u] = 0.f; g[3 * H + u] = 0.f; *dcc = 0.f; return; }
        const float gi = g[u], gj = g[H + u], gf = g[2 * H + u], go = g[3 * H + u];
        const float c = s.cstate[((size_t)t * s.B + b) * 2 * H + (size_t)d * H + u];
        const float cp = has_prev ? s.cstate[((size_t)tp * s.B + b) * 2 * H + (size_t)d * H + u] : 0.f;
        const float tc = tanhf(c);
        const float dc = dh * go * (1.f - tc * tc) + (i == s.T - 1 ? 0.f : *dcc);
        g[u] = dc * gj * gi * (1.f - gi);
        g[H + u] = dc * gi * (1.f - gj * gj);
        g[2 * H + u] = dc * cp * gf * (1.f - gf);
        g[3 * H + u] = dh * tc * go * (1.f - go);
        *dcc = dc * gf;
    } else {
        if (!live) { g[u] = 0.f; return; }
        const float h = g[u];
        g[u] = s.cell == CTCASR_CELL_RNN_TANH ? dh * (1.f - h * h) : (h > 0.f ? dh : 0.f);
    }
}

int rnn_cell_fwd(const RnnStep &s, int i, cudaStream_t stream)
{
    dim3 grid(ceil_div(s.B * s.H, 256), 2);
    rnn_cell_fwd_kernel<<<grid, 256, 0, stream>>>(s, i);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}
int rnn_cell_bwd(const RnnStep &s, int i, cudaStream_t stream)
{
    dim3 grid(ceil_div(s.B * s.H, 256), 2);
    rnn_cell_bwd_kernel<<<grid, 256, 0, stream>>>(s, i);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// ---- Adam (TF1 formulation), one pass over the flat buffers -----------------------------------
__global__ void adam_kernel(float *__restrict__ p, float *__restrict__ m, float *__restrict__ v,
                            const float *__restrict__ g, size_t n, float lr_t, float b1, float b2, float eps,
                            float gscale)
{
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i];
        float4 vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gg = reinterpret_cast<const float4 *>(g)[i];
        float *pa = &pp.x, *ma = &mm.x, *va = &vv.x;
        const float *ga = &gg.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = ga[j] * gscale;
            ma[j] = b1 * ma[j] + (1.f - b1) * gj;
            va[j] = b2 * va[j] + (1.f - b2) * gj * gj;
            pa[j] -= lr_t * ma[j] / (sqrtf(va[j]) + eps);
        }
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    if (blockIdx.x == 0)
        for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            const float gj = g[i] * gscale;
            m[i] = b1 * m[i] + (1.f - b1) * gj;
            v[i] = b2 * v[i] + (1.f - b2) * gj * gj;
            p[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
        }
}

}  // namespace ctcasr

using namespace ctcasr;

extern "C" int ctcasr_transpose01(const float *in, float *out, int A, int B, int C, void *stream)
{
    CTCASR_REQUIRE(in && out && A >= 0 && B >= 0 && C >= 1, "transpose01: bad args");
    const size_t total = (size_t)A * B * C;
    if (total == 0) return CTCASR_OK;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    transpose01_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, A, B, C);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

extern "C" int ctcasr_adam(float *p, float *m, float *v, const float *g, size_t n, int step,
                           float lr, float beta1, float beta2, float eps, float grad_scale, void *stream)
{
    CTCASR_REQUIRE(p && m && v && g && step >= 1, "adam: bad args");
    CTCASR_REQUIRE(((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) % 16 == 0, "adam: buffers must be 16-B aligned");
    if (n == 0) return CTCASR_OK;
    const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step)));
    const size_t want = (n / 4 + 255) / 256;
    const int blocks = (int)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, m, v, g, n, lr_t, beta1, beta2, eps, grad_scale);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

extern "C" int ctcasr_dropout(const float *x, float *y, size_t n, float drop_rate, uint32_t seed, void *stream_)
{
    using namespace ctcasr;
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(x && y, "dropout: null pointer");
    CTCASR_REQUIRE(drop_rate >= 0.f && drop_rate < 1.f, "dropout: drop_rate %g", drop_rate);
    if (n == 0) return CTCASR_OK;
    if (drop_rate == 0.f) {
        if (x != y) CTCASR_CUDA_CHECK(cudaMemcpyAsync(y, x, n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        return CTCASR_OK;
    }
    const size_t blocks = (n + 255) / 256;
    dropout_kernel<<<(int)(blocks < (size_t)148 * 16 ? blocks : (size_t)148 * 16), 256, 0, stream>>>(x, y, n, drop_rate, seed);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}
