// gemm_simt.cu — generic fp32 FFMA GEMM (CTCASR_COMPUTE_FP32): any shape, any leading dimension,
// both operand orientations, fused epilogue.  This is the exact-fp32 parity path and the shape
// fallback (K = 80 feature columns, N = 29 classes, H = 128 ...) for the tcgen05 kernels in
// gemm_tc.cu; it is device code, not a CPU fallback.
//
// C[M,N] = op(A) op(B);  TA == false: A[m*lda + k], true: A[k*lda + m];
//                        TB == false: B[k*ldb + n], true: B[n*ldb + k].
// gridDim.z selects one of up to two independent problems (the fw / bw direction of a recurrent
// step) that share shapes but not pointers.
#include "gemm.cuh"

namespace ctcasr {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

// Split-K (few output tiles, long K: the [2048 x 29] weight gradient of the logits layer contracts 32000
// frames): slice s of the K range goes to CTA z = s * nz + problem and writes a raw partial tile to
// `part`; splitk_reduce_kernel adds the slices in a fixed order (deterministic) and applies the epilogue.
struct SplitK {
    int S = 1, Kc = 0;
    float *part = nullptr;      // [S][nz][M][N]
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g, const SplitK sk)
{
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int z = blockIdx.z % g.nz, slice = blockIdx.z / g.nz;
    const float *__restrict__ A = g.A[z];
    const float *__restrict__ B = g.B[z];
    float *__restrict__ C = g.C[z];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[TM][TN] = {};
    const int k_begin = sk.S > 1 ? slice * sk.Kc : 0;
    const int k_end = sk.S > 1 ? min(g.K, k_begin + sk.Kc) : g.K;

    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int j = 0; j < (BM * BK) / 256; ++j) {
            const int i = tid + j * 256;
            int m, k;
            if (TA) { k = i / BM; m = i % BM; } else { m = i / BK; k = i % BK; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < g.M && gk < k_end) v = TA ? A[(size_t)gk * g.lda + gm] : A[(size_t)gm * g.lda + gk];
            As[k][m] = v;
        }
#pragma unroll
        for (int j = 0; j < (BN * BK) / 256; ++j) {
            const int i = tid + j * 256;
            int n, k;
            if (TB) { n = i / BK; k = i % BK; } else { k = i / BN; n = i % BN; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < g.N && gk < k_end) v = TB ? B[(size_t)gn * g.ldb + gk] : B[(size_t)gk * g.ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            if (sk.S > 1) {
                sk.part[(((size_t)slice * g.nz + z) * g.M + m) * g.N + n] = acc[i][j];
                continue;
            }
            float *c = C + (size_t)m * g.ldc + n;
            const float old = g.epi.accumulate ? *c : 0.f;
            *c = epilogue_apply(g.epi, acc[i][j], m, n, g.N, old);
        }
    }
}

__global__ void splitk_reduce_kernel(const GemmArgs g, const SplitK sk)
{
    const size_t per = (size_t)g.M * g.N, total = per * g.nz;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(i / per);
        const size_t r = i % per;
        const int m = (int)(r / g.N), n = (int)(r % g.N);
        float v = 0.f;
        for (int s = 0; s < sk.S; ++s) v += sk.part[((size_t)s * g.nz + z) * per + r];
        float *c = g.C[z] + (size_t)m * g.ldc + n;
        const float old = g.epi.accumulate ? *c : 0.f;
        *c = epilogue_apply(g.epi, v, m, n, g.N, old);
    }
}

int gemm_simt(const GemmArgs &g, cudaStream_t stream)
{
    if (g.M <= 0 || g.N <= 0) return CTCASR_OK;
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), g.nz);
    SplitK sk;
    const int ctas = grid.x * grid.y * grid.z;
    if (ctas <= 74 && g.K >= 4096) {        // under half a wave of CTAs on a long contraction
        int S = (2 * 148) / ctas;
        const int max_s = g.K / 1024;       // at least 1024 k per slice
        if (S > max_s) S = max_s;
        if (S > 1) {
            const int Kc = ceil_div(ceil_div(g.K, S), BK) * BK;
            S = ceil_div(g.K, Kc);
            float *part = reinterpret_cast<float *>(scratch_free((size_t)S * g.nz * g.M * g.N * sizeof(float)));
            if (part && S > 1) { sk.S = S; sk.Kc = Kc; sk.part = part; grid.z = g.nz * S; }   // no arena: unsplit (slower, same result class)
        }
    }
    if (g.ta && g.tb) gemm_simt_kernel<true, true><<<grid, 256, 0, stream>>>(g, sk);
    else if (g.ta) gemm_simt_kernel<true, false><<<grid, 256, 0, stream>>>(g, sk);
    else if (g.tb) gemm_simt_kernel<false, true><<<grid, 256, 0, stream>>>(g, sk);
    else gemm_simt_kernel<false, false><<<grid, 256, 0, stream>>>(g, sk);
    CTCASR_LAUNCH_CHECK();
    if (sk.S > 1) return splitk_reduce(g, sk.S, sk.part, stream);
    return CTCASR_OK;
}

int splitk_reduce(const GemmArgs &g, int S, float *part, cudaStream_t stream)
{
    SplitK sk;
    sk.S = S; sk.part = part;
    const size_t total = (size_t)g.M * g.N * g.nz;
    splitk_reduce_kernel<<<(int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, stream>>>(g, sk);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

}  // namespace ctcasr
