// step_gemm.cu — the per-frame recurrent product of the stepwise RNN path (rnn.cu): for the cells without a
// persistent kernel (tanh / ReLU RNN, GRU: the rest of the reference's rnn_cell menu, asr/params.py:48-50)
// every frame needs  C[z][B, N] (+)= A[z][B, K] op(W[z])  for both directions z with B <= 32 rows.
// In the generic SIMT GEMM that shape is a 128-iteration dependent chain of un-prefetched loads (~80 us a
// call); here a CTA owns 32 output columns of one direction, streams its W and A chunks through a
// double-buffered cp.async pipeline (the recurrent weights of a simple cell, 33.5 MB at H = 2048, stay in
// L2 across frames) and keeps a 2 x 2 register tile per thread.  Exact fp32 FFMA, fixed summation order.
#include "gemm.cuh"

namespace ctcasr {
namespace stepg {

constexpr int BM = 32, BN = 32, BKC = 64, LDS_ = BKC + 4;      // 68-float rows: 16-B aligned, conflict-light

__device__ __forceinline__ void cp16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// TB == false: W stored [K][N] (forward: h Wh);  TB == true: W stored [N][K] (backward: dz Wh^T)
template <bool TB>
__global__ void __launch_bounds__(256) step_gemm_kernel(const GemmArgs g)
{
    __shared__ __align__(16) float As[2][BM][LDS_];
    __shared__ __align__(16) float Ws[2][TB ? BN : BKC][TB ? LDS_ : BN + 4];
    const int z = blockIdx.z, m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const float *__restrict__ A = g.A[z] + (size_t)m0 * g.lda;
    const float *__restrict__ W = g.B[z];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int rows = min(BM, g.M - m0);
    // rows >= M of the A tile are never loaded: zero them once
    for (int i = tid; i < 2 * BM * LDS_; i += 256) (&As[0][0][0])[i] = 0.f;
    __syncthreads();

    auto load = [&](int buf, int k0) {
        // A chunk [rows][64]: 16 x 16-B pieces per row
        for (int i = tid; i < BM * (BKC / 4); i += 256) {
            const int r = i / (BKC / 4), c4 = (i % (BKC / 4)) * 4;
            if (r < rows) cp16(&As[buf][r][c4], A + (size_t)r * g.lda + k0 + c4);
        }
        if (TB) {       // W[n][k]: 32 rows of 64 contiguous k
            for (int i = tid; i < BN * (BKC / 4); i += 256) {
                const int r = i / (BKC / 4), c4 = (i % (BKC / 4)) * 4;
                cp16(&Ws[buf][r][c4], W + (size_t)(n0 + r) * g.ldb + k0 + c4);
            }
        } else {        // W[k][n]: 64 rows of 32 contiguous n
            for (int i = tid; i < BKC * (BN / 4); i += 256) {
                const int r = i / (BN / 4), c4 = (i % (BN / 4)) * 4;
                cp16(&Ws[buf][r][c4], W + (size_t)(k0 + r) * g.ldb + n0 + c4);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const int nk = g.K / BKC;
    load(0, 0);
    for (int it = 0; it < nk; ++it) {
        const int buf = it & 1;
        if (it + 1 < nk) {
            load(buf ^ 1, (it + 1) * BKC);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
#pragma unroll 16
        for (int k = 0; k < BKC; ++k) {
            const float a0 = As[buf][ty][k], a1 = As[buf][ty + 16][k];
            const float b0 = TB ? Ws[buf][tx][k] : Ws[buf][k][tx];
            const float b1 = TB ? Ws[buf][tx + 16][k] : Ws[buf][k][tx + 16];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
    float *C = g.C[z];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float *c = C + (size_t)m * g.ldc + n0 + tx + 16 * j;
            *c = g.epi.accumulate ? acc[i][j] + *c : acc[i][j];
        }
    }
}

}  // namespace stepg

// Falls back to the generic SIMT GEMM when the shape is not the recurrence's (plain store / accumulate epilogue,
// A not transposed, N % 32 == 0, K % 64 == 0, 16-B aligned rows).
int step_gemm(const GemmArgs &g, cudaStream_t stream)
{
    bool ok = g.epi.mode == EPI_STORE && !g.ta && g.M >= 1 && g.N % stepg::BN == 0 && g.K % stepg::BKC == 0 && g.K >= stepg::BKC &&
              g.lda % 4 == 0 && g.ldb % 4 == 0;
    for (int z = 0; z < g.nz && ok; ++z)
        ok = (((uintptr_t)g.A[z] | (uintptr_t)g.B[z]) & 15) == 0;
    if (!ok) return gemm_simt(g, stream);
    dim3 grid(g.N / stepg::BN, ceil_div(g.M, stepg::BM), g.nz);
    if (g.tb) stepg::step_gemm_kernel<true><<<grid, 256, 0, stream>>>(g);
    else stepg::step_gemm_kernel<false><<<grid, 256, 0, stream>>>(g);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

}  // namespace ctcasr
