// step_gemm.cu — the per-frame recurrent product of the stepwise RNN path (rnn.cu): for the cells without a
// persistent kernel (tanh / ReLU RNN, GRU: the rest of the reference's rnn_cell menu, asr/params.py:48-50)
// every frame needs  C[z][B, N] (+)= A[z][B, K] op(W[z])  for both directions z with B <= 32 rows.
// In the generic SIMT GEMM that shape is a 128-iteration dependent chain of un-prefetched loads (~80 us a
// call); here a CTA owns 32 output columns of one direction, streams its W and A chunks through a
// six-deep cp.async pipeline (the recurrent weights of a simple cell, 33.5 MB at H = 2048, stay in
// L2 across frames) and keeps a 2 x 2 register tile per thread.  Exact fp32 FFMA, fixed summation order.
#include "gemm.cuh"

#include <cooperative_groups.h>

namespace ctcasr {
namespace stepg {
namespace cg = cooperative_groups;

constexpr int KSPLIT = 4;       // CTAs per cluster: each takes a quarter of K, partial sums meet through DSMEM

constexpr int BM = 32, BN = 64, BKC = 32, LDS_ = BKC + 4;      // 36-float rows: 16-B aligned, conflict-light
constexpr int LDW_ = BN + 4;                                    // W rows in the [k][n] layout
constexpr int NSTAGE = 4;                                       // chunks in flight
constexpr int A_FLOATS = BM * LDS_;                             // A chunk [32 m][32 k]
constexpr int STAGE_FLOATS = A_FLOATS + 64 * LDS_;              // + W chunk: [64 n][32 k] (TB) or [32 k][64 n] (2176 <= 2304)
constexpr int SMEM_BYTES = NSTAGE * STAGE_FLOATS * 4;           // 55,296 B: four CTAs per SM

__device__ __forceinline__ void cp16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// TB == false: W stored [K][N] (forward: h Wh);  TB == true: W stored [N][K] (backward: dz Wh^T)
// MODE 0: C (+)= A op(W).  MODE 1 / 2: the cell math of the one-gate cells (tanh / ReLU RNN) in the epilogue,
// so that a frame of the recurrence is ONE launch:
//   1 forward   C holds the input projection P_t;  h = act(P_t + acc) -> C (saved activation) and y
//   2 backward  C holds the saved activation h_t;  dz = (dy_t + acc) act'(h_t) -> C
// (rows past an utterance's length produce zeros: dynamic_rnn(sequence_length) semantics, rnn.cu)
// Thread = two output columns (lane, lane + 32), all 32 batch rows in registers: per group of four k the warp
// reads the W values of its columns and the 32 A rows as BROADCAST 16-B loads — 34 shared-memory instructions
// per 256 FMAs (a 16-B shared load costs four issue cycles even when broadcast: with a 2 x 2 register tile, or
// with one column per thread, the kernel was bound by the shared-memory pipe at 40-50 us a frame).  The four
// warps of the CTA split every 32-k chunk between them; partial sums are added in a fixed order.
// One SM's worth of this product (a 32 x 32 tile over all of K) is only four warps: to fill the FMA pipes the
// K range is cut over a cluster of four CTAs (4 x as many resident warps per SM to hide the shared-memory
// latency); rank 0 adds the 16 partial tiles (cluster rank, then warp: fixed order) out of its peers'
// shared memory and runs the epilogue.
template <bool TB, int MODE>
__global__ void __cluster_dims__(1, 1, KSPLIT) __launch_bounds__(128) step_gemm_kernel(const GemmArgs g, const StepCell sc)
{
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int krank = (int)cluster.block_rank();
    const int z = blockIdx.z / KSPLIT, m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const float *__restrict__ A = g.A[z] + (size_t)m0 * g.lda;
    const float *__restrict__ W = g.B[z];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = min(BM, g.M - m0);
    // rows >= M of the A chunks are never loaded: zero them once
    for (int i = tid; i < NSTAGE * STAGE_FLOATS; i += 128) smem[i] = 0.f;
    __syncthreads();
    // per chunk every thread moves two 16-B pieces of A (rows lr, lr + 16) and four of W
    const int lr = tid >> 3, lc4 = (tid & 7) * 4;
    const float *a_src = A + (size_t)lr * g.lda + lc4;
    auto load = [&](int chunk) {
        float *st = smem + (chunk % NSTAGE) * STAGE_FLOATS;
        if (lr < rows) cp16(st + lr * LDS_ + lc4, a_src + (size_t)chunk * BKC);
        if (lr + 16 < rows) cp16(st + (lr + 16) * LDS_ + lc4, a_src + (size_t)16 * g.lda + (size_t)chunk * BKC);
        float *ws = st + A_FLOATS;
        if (TB) {       // W[n][k]: 64 rows (n) of 32 k
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = lr + 16 * j;
                cp16(ws + r * LDS_ + lc4, W + (size_t)(n0 + r) * g.ldb + (size_t)chunk * BKC + lc4);
            }
        } else {        // W[k][n]: 32 rows (k) of 64 n: 16 pieces per row
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = tid + 128 * j, r = i >> 4, c4 = (i & 15) * 4;
                cp16(ws + r * LDW_ + c4, W + ((size_t)chunk * BKC + r) * g.ldb + n0 + c4);
            }
        }
    };

    float acc0[BM], acc1[BM];                                   // columns lane and lane + 32
#pragma unroll
    for (int m = 0; m < BM; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }
    const int nk = g.K / BKC / KSPLIT, c0 = krank * nk;           // this CTA's chunks: [c0, c0 + nk)
    for (int c = 0; c < NSTAGE - 1; ++c) {
        if (c < nk) load(c0 + c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int it = 0; it < nk; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NSTAGE - 2) : "memory");       // chunk `it` has landed
        __syncthreads();                                                           // ... for everyone; chunk it-1 is consumed
        if (it + NSTAGE - 1 < nk) load(c0 + it + NSTAGE - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float *As = smem + ((c0 + it) % NSTAGE) * STAGE_FLOATS, *Ws = As + A_FLOATS;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k0 = (2 * warp + q) * 4;
            float4 w0, w1;
            if (TB) {
                w0 = *reinterpret_cast<const float4 *>(Ws + lane * LDS_ + k0);
                w1 = *reinterpret_cast<const float4 *>(Ws + (lane + 32) * LDS_ + k0);
            } else {
                w0 = make_float4(Ws[k0 * LDW_ + lane], Ws[(k0 + 1) * LDW_ + lane], Ws[(k0 + 2) * LDW_ + lane], Ws[(k0 + 3) * LDW_ + lane]);
                w1 = make_float4(Ws[k0 * LDW_ + lane + 32], Ws[(k0 + 1) * LDW_ + lane + 32], Ws[(k0 + 2) * LDW_ + lane + 32],
                                 Ws[(k0 + 3) * LDW_ + lane + 32]);
            }
#pragma unroll
            for (int m = 0; m < BM; ++m) {
                const float4 a = *reinterpret_cast<const float4 *>(As + m * LDS_ + k0);      // same address in every lane
                acc0[m] = fmaf(a.x, w0.x, acc0[m]); acc1[m] = fmaf(a.x, w1.x, acc1[m]);
                acc0[m] = fmaf(a.y, w0.y, acc0[m]); acc1[m] = fmaf(a.y, w1.y, acc1[m]);
                acc0[m] = fmaf(a.z, w0.z, acc0[m]); acc1[m] = fmaf(a.z, w1.z, acc1[m]);
                acc0[m] = fmaf(a.w, w0.w, acc0[m]); acc1[m] = fmaf(a.w, w1.w, acc1[m]);
            }
        }
    }
    // ---- partial tiles of the four warps into this CTA's shared memory; rank 0 adds all 16 (cluster rank, then warp)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float *red = smem;                                          // [4 warps][32 m][65]
    constexpr int RLD = BN + 1;
#pragma unroll
    for (int m = 0; m < BM; ++m) {
        red[(warp * BM + m) * RLD + lane] = acc0[m];
        red[(warp * BM + m) * RLD + lane + 32] = acc1[m];
    }
    cluster.sync();
    if (krank != 0) { cluster.sync(); return; }                 // peers keep their tiles alive until rank 0 has read them
    const float *peer[KSPLIT];
#pragma unroll
    for (int r = 0; r < KSPLIT; ++r) peer[r] = cluster.map_shared_rank(red, r);
    float *C = g.C[z];
    float vsum[16];                                             // rows warp*8 .. +7, columns lane and lane + 32
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int ml = warp * 8 + (j >> 1), nl = lane + 32 * (j & 1);
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < KSPLIT; ++r)
#pragma unroll
            for (int w = 0; w < 4; ++w) v += peer[r][(w * BM + ml) * RLD + nl];
        vsum[j] = v;
    }
    cluster.sync();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int ml = warp * 8 + (j >> 1), m = m0 + ml, n = n0 + lane + 32 * (j & 1);
        if (m >= g.M) break;
        const float v = vsum[j];
        const bool live = MODE == 0 || !sc.use_len || sc.t[z] < sc.seq_len[m];
        float *c = C + (size_t)m * g.ldc + n;
        if (MODE == 0) {
            *c = g.epi.accumulate ? v + *c : v;
        } else if (MODE == 1) {
            const float zz = *c + v;
            const float h = !live ? 0.f : (sc.cell == CTCASR_CELL_RNN_TANH ? tanhf(zz) : fmaxf(zz, 0.f));
            *c = h;
            sc.y[z][(size_t)m * sc.ldy + n] = h;
        } else {
            const float h = *c;
            const float dh = sc.dy[z][(size_t)m * sc.ldy + n] + v;
            *c = !live ? 0.f : (sc.cell == CTCASR_CELL_RNN_TANH ? dh * (1.f - h * h) : (h > 0.f ? dh : 0.f));
        }
    }
}

}  // namespace stepg

static int set_smem_once()
{
    static bool done = false;
    if (done) return CTCASR_OK;
    CTCASR_CUDA_CHECK(cudaFuncSetAttribute(stepg::step_gemm_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, stepg::SMEM_BYTES));
    CTCASR_CUDA_CHECK(cudaFuncSetAttribute(stepg::step_gemm_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, stepg::SMEM_BYTES));
    CTCASR_CUDA_CHECK(cudaFuncSetAttribute(stepg::step_gemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, stepg::SMEM_BYTES));
    CTCASR_CUDA_CHECK(cudaFuncSetAttribute(stepg::step_gemm_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, stepg::SMEM_BYTES));
    done = true;
    return CTCASR_OK;
}

static bool eligible(const GemmArgs &g)
{
    bool ok = g.epi.mode == EPI_STORE && !g.ta && g.M >= 1 && g.N % stepg::BN == 0 && g.K % (32 * stepg::KSPLIT) == 0 &&
              g.lda % 4 == 0 && g.ldb % 4 == 0;
    for (int z = 0; z < g.nz && ok; ++z)
        ok = (((uintptr_t)g.A[z] | (uintptr_t)g.B[z]) & 15) == 0;
    return ok;
}

// Falls back to the generic SIMT GEMM when the shape is not the recurrence's (plain store / accumulate epilogue,
// A not transposed, N % 32 == 0, K % 64 == 0, 16-B aligned rows).
int step_gemm(const GemmArgs &g, cudaStream_t stream)
{
    if (!eligible(g)) return gemm_simt(g, stream);
    if (int rc = set_smem_once()) return rc;
    dim3 grid(g.N / stepg::BN, ceil_div(g.M, stepg::BM), g.nz * stepg::KSPLIT);
    const StepCell none{};
    if (g.tb) stepg::step_gemm_kernel<true, 0><<<grid, 128, stepg::SMEM_BYTES, stream>>>(g, none);
    else stepg::step_gemm_kernel<false, 0><<<grid, 128, stepg::SMEM_BYTES, stream>>>(g, none);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// One frame of a one-gate cell, product + cell math in one launch (sc.mode 1 forward: g.tb == 0, 2 backward:
// g.tb == 1).  CTCASR_ERR_UNSUPPORTED (nothing launched) when the shape is not eligible: the caller then
// runs the two-launch path.
int step_gemm_cell(const GemmArgs &g, const StepCell &sc, cudaStream_t stream)
{
    if (!eligible(g) || g.nz != 2 || (sc.mode != 1 && sc.mode != 2) || (sc.mode == 1) == (g.tb != 0)) return CTCASR_ERR_UNSUPPORTED;
    if (int rc = set_smem_once()) return rc;
    dim3 grid(g.N / stepg::BN, ceil_div(g.M, stepg::BM), g.nz * stepg::KSPLIT);
    if (sc.mode == 1) stepg::step_gemm_kernel<false, 1><<<grid, 128, stepg::SMEM_BYTES, stream>>>(g, sc);
    else stepg::step_gemm_kernel<true, 2><<<grid, 128, stepg::SMEM_BYTES, stream>>>(g, sc);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

}  // namespace ctcasr
