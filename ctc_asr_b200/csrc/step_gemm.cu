// step_gemm.cu — the per-frame recurrent product of the stepwise RNN path (rnn.cu): for the cells without a
// persistent kernel (tanh / ReLU RNN, GRU: the rest of the reference's rnn_cell menu, asr/params.py:48-50)
// every frame needs  C[z][B, N] (+)= A[z][B, K] op(W[z])  for both directions z with B <= 32 rows.
// In the generic SIMT GEMM that shape is a 128-iteration dependent chain of un-prefetched loads (~80 us a
// call); here a CTA owns 32 output columns of one direction, streams its W and A chunks through a
// six-deep cp.async pipeline (the recurrent weights of a simple cell, 33.5 MB at H = 2048, stay in
// L2 across frames) and keeps a 2 x 2 register tile per thread.  Exact fp32 FFMA, fixed summation order.
#include "gemm.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>

namespace ctcasr {
namespace stepg {
namespace cg = cooperative_groups;

constexpr int BM = 32, BKC = 32, LDS_ = BKC + 4;               // 36-float rows: 16-B aligned, conflict-light

// NW warps per CTA, NC output columns per lane (tile = 32 rows x 32 NC columns), KS CTAs per cluster (K split)
template <int NW, int NC, int KS>
struct Cfg {
    static constexpr int NT = 32 * NW, BN = 32 * NC, LDW = BN + 4;
    static constexpr int A_FLOATS = BM * LDS_;
    static constexpr int W_FLOATS = BN * LDS_ > BKC * LDW ? BN * LDS_ : BKC * LDW;   // [BN n][32 k] (TB) or [32 k][BN n]
    static constexpr int STAGE_FLOATS = A_FLOATS + W_FLOATS;
    static constexpr int NSTAGE = NC == 1 ? 6 : 4;              // 55,296 B either way
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_FLOATS * 4;
    static constexpr int QW = 8 / NW;                           // groups of four k per warp and chunk
    static constexpr int RLD = BN + 1;
    static_assert(NW * BM * RLD * 4 <= SMEM_BYTES, "reduction tile must fit in the pipeline buffers");
};

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
//
// A lane owns NC output columns and keeps all 32 batch rows in registers: per group of four k the warp reads
// the W values of its columns and the 32 A rows as BROADCAST 16-B loads.  The warps of a CTA split every 32-k
// chunk between them, the K range is cut over a cluster of KS CTAs (more resident warps per SM to hide the
// shared-memory latency), and rank 0 adds the KS x NW partial tiles out of its peers' shared memory in a
// fixed order (deterministic) before the epilogue.
template <bool TB, int MODE, int NW, int NC, int KS>
__global__ void __cluster_dims__(1, 1, KS) __launch_bounds__(32 * NW) step_gemm_kernel(const GemmArgs g, const StepCell sc)
{
    using C_ = Cfg<NW, NC, KS>;
    constexpr int NT = C_::NT, BN = C_::BN, LDW = C_::LDW, NSTAGE = C_::NSTAGE, STAGE_FLOATS = C_::STAGE_FLOATS;
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int krank = (int)cluster.block_rank();
    const int z = blockIdx.z / KS, m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const float *__restrict__ A = g.A[z] + (size_t)m0 * g.lda;
    const float *__restrict__ W = g.B[z];
    const int lda = g.lda, ldb = g.ldb;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = min(BM, g.M - m0);
    // rows >= M of the A chunks are never loaded: zero them once
    for (int i = tid; i < NSTAGE * STAGE_FLOATS; i += NT) smem[i] = 0.f;
    __syncthreads();
    auto load = [&](int chunk) {
        float *st = smem + (chunk % NSTAGE) * STAGE_FLOATS;
        float *ws = st + C_::A_FLOATS;
#pragma unroll
        for (int i = tid; i < BM * 8; i += NT) {
            const int r = i >> 3, c4 = (i & 7) * 4;
            if (r < rows) cp16(st + r * LDS_ + c4, A + (size_t)r * lda + (size_t)chunk * BKC + c4);
        }
        if (TB) {       // W[n][k]: BN rows (n) of 32 k
#pragma unroll
            for (int i = tid; i < BN * 8; i += NT) {
                const int r = i >> 3, c4 = (i & 7) * 4;
                cp16(ws + r * LDS_ + c4, W + (size_t)(n0 + r) * ldb + (size_t)chunk * BKC + c4);
            }
        } else {        // W[k][n]: 32 rows (k) of BN n
#pragma unroll
            for (int i = tid; i < BKC * (BN / 4); i += NT) {
                const int r = i / (BN / 4), c4 = (i % (BN / 4)) * 4;
                cp16(ws + r * LDW + c4, W + ((size_t)chunk * BKC + r) * ldb + n0 + c4);
            }
        }
    };

    float acc[NC][BM];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int m = 0; m < BM; ++m) acc[c][m] = 0.f;
    const int nk = g.K / BKC / KS, c0 = krank * nk;              // this CTA's chunks: [c0, c0 + nk)
    for (int c = 0; c < NSTAGE - 1; ++c) {
        if (c < nk) load(c0 + c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int it = 0; it < nk; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NSTAGE - 2) : "memory");       // chunk `it` has landed
        __syncthreads();                                                           // ... for everyone; chunk it-1 is consumed
        if (it + NSTAGE - 1 < nk) load(c0 + it + NSTAGE - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float *As = smem + ((c0 + it) % NSTAGE) * STAGE_FLOATS, *Ws = As + C_::A_FLOATS;
#pragma unroll
        for (int q = 0; q < C_::QW; ++q) {
            const int k0 = (warp * C_::QW + q) * 4;
            float4 w[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int nl = lane + 32 * c;
                if (TB) w[c] = *reinterpret_cast<const float4 *>(Ws + nl * LDS_ + k0);
                else w[c] = make_float4(Ws[k0 * LDW + nl], Ws[(k0 + 1) * LDW + nl], Ws[(k0 + 2) * LDW + nl], Ws[(k0 + 3) * LDW + nl]);
            }
            // eight rows at a time, one k at a time: consecutive FMAs belong to different accumulators, so the
            // four-deep dependent chain of an accumulator is spread over 8 NC independent instructions
#pragma unroll
            for (int mg = 0; mg < BM; mg += 8) {
                float4 a[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(As + (mg + i) * LDS_ + k0);   // broadcast loads
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < NC; ++c) acc[c][mg + i] = fmaf(a[i].x, w[c].x, acc[c][mg + i]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < NC; ++c) acc[c][mg + i] = fmaf(a[i].y, w[c].y, acc[c][mg + i]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < NC; ++c) acc[c][mg + i] = fmaf(a[i].z, w[c].z, acc[c][mg + i]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < NC; ++c) acc[c][mg + i] = fmaf(a[i].w, w[c].w, acc[c][mg + i]);
            }
        }
    }
    // ---- partial tiles of the warps into this CTA's shared memory; rank 0 adds all KS x NW (cluster rank, then warp)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float *red = smem;                                          // [NW][32 m][RLD]
    constexpr int RLD = C_::RLD;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int m = 0; m < BM; ++m) red[(warp * BM + m) * RLD + lane + 32 * c] = acc[c][m];
    cluster.sync();
    if (krank != 0) { cluster.sync(); return; }                 // peers keep their tiles alive until rank 0 has read them
    const float *peer[KS];
#pragma unroll
    for (int r = 0; r < KS; ++r) peer[r] = cluster.map_shared_rank(red, r);
    constexpr int OUT = BM * BN / NT;                           // outputs per thread: o = tid + j * NT -> (o / BN, o % BN)
    float vsum[OUT];
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
        const int o = tid + j * NT, ml = o / BN, nl = o % BN;
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < KS; ++r)
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) v += peer[r][(w2 * BM + ml) * RLD + nl];
        vsum[j] = v;
    }
    cluster.sync();
    float *C = g.C[z];
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
        const int o = tid + j * NT, ml = o / BN, m = m0 + ml, n = n0 + o % BN;
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

// variant 0: 4 warps, 2 columns per lane, K over 4 CTAs;  variant 1: 8 warps, 1 column per lane, K over 2 CTAs
// (twice the resident warps per SM).  CTCASR_STEP_VARIANT selects (measurement aid).
static int variant()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CTCASR_STEP_VARIANT"); v = e ? atoi(e) : 0; if (v != 1) v = 0; }
    return v;
}

template <bool TB, int MODE, int NW, int NC, int KS>
static int launch(const GemmArgs &g, const StepCell &sc, cudaStream_t stream)
{
    using C_ = Cfg<NW, NC, KS>;
    static bool attr_set = false;
    if (!attr_set) {
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(step_gemm_kernel<TB, MODE, NW, NC, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(g.N / C_::BN, ceil_div(g.M, BM), g.nz * KS);
    step_gemm_kernel<TB, MODE, NW, NC, KS><<<grid, C_::NT, C_::SMEM_BYTES, stream>>>(g, sc);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

template <bool TB, int MODE>
static int launch_variant(const GemmArgs &g, const StepCell &sc, cudaStream_t stream)
{
    return variant() == 1 ? launch<TB, MODE, 8, 1, 2>(g, sc, stream) : launch<TB, MODE, 4, 2, 4>(g, sc, stream);
}

}  // namespace stepg

static bool eligible(const GemmArgs &g)
{
    bool ok = g.epi.mode == EPI_STORE && !g.ta && g.M >= 1 && g.N % 64 == 0 && g.K % 128 == 0 && g.lda % 4 == 0 && g.ldb % 4 == 0;
    for (int z = 0; z < g.nz && ok; ++z)
        ok = (((uintptr_t)g.A[z] | (uintptr_t)g.B[z]) & 15) == 0;
    return ok;
}

// Falls back to the generic SIMT GEMM when the shape is not the recurrence's (plain store / accumulate epilogue,
// A not transposed, N % 64 == 0, K % 128 == 0, 16-B aligned rows).
int step_gemm(const GemmArgs &g, cudaStream_t stream)
{
    if (!eligible(g)) return gemm_simt(g, stream);
    const StepCell none{};
    return g.tb ? stepg::launch_variant<true, 0>(g, none, stream) : stepg::launch_variant<false, 0>(g, none, stream);
}

// One frame of a one-gate cell, product + cell math in one launch (sc.mode 1 forward: g.tb == 0, 2 backward:
// g.tb == 1).  CTCASR_ERR_UNSUPPORTED (nothing launched) when the shape is not eligible: the caller then
// runs the two-launch path.
int step_gemm_cell(const GemmArgs &g, const StepCell &sc, cudaStream_t stream)
{
    if (!eligible(g) || g.nz != 2 || (sc.mode != 1 && sc.mode != 2) || (sc.mode == 1) == (g.tb != 0)) return CTCASR_ERR_UNSUPPORTED;
    return sc.mode == 1 ? stepg::launch_variant<false, 1>(g, sc, stream) : stepg::launch_variant<true, 2>(g, sc, stream);
}

}  // namespace ctcasr
