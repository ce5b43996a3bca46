// conv_tc.cu — the 'ds2' conv layers as implicit GEMMs on tcgen05 (sm_100a): tf.layers.conv2d(SAME) + relu + tf.minimum +
// tf.layers.dropout (asr/util/tf_contrib.py:123-135), forward pass, weight gradient and input gradient, WITHOUT a patch
// matrix (three kernels: conv_fwd_kernel here at the top, conv_wgrad_kernel and conv_dgrad_kernel further down).
//
// conv.cu writes every output position's kt x kf x C patch to HBM (14.2 GB of bf16 pieces for the second layer at
// B = 32 x 10 s) and multiplies that matrix by the kernel.  Here the A operand of the same GEMM is gathered by the TMA
// unit straight from the layer input: the input's bf16 pieces are viewed as a 5-D tensor [piece][T][B][F][C] and one
// k-block of the contraction is ONE TAP (it, jf) x 32 channels, i.e. the box
//     { 32 channels, FOB output frequencies (traversal stride sf), BB utterances, TB output frames (stride st), 1 piece }
// at coordinate (c0, fo0*sf + jf - pf, b0, to0*st + it - pt, piece) — FOB*BB*TB = 128 rows of 64 B, exactly the K-major
// SWIZZLE_64B tile the MMA reads.  'SAME' padding is the TMA's out-of-range zero fill (coordinates may be negative), so
// neither the input nor the patches are ever copied; the input pieces (181 MB for that layer) stay in L2.  The B operand
// is the kernel [Kp, N] in its HWIO row order (row = (it*kf + jf)*C + c = k-block * 32 + c), split as in gemm_tc.cu.
// Arithmetic, pipeline and epilogue are those of gemm_tc_kernel: NP = 3 pieces / 6 products (bf16x3 through the ReLU
// kink), NP = 2 / 3 products (the gradients) or NP = 1 (compute = 'bf16'), fp32 accumulation in tensor memory, TMA
// producer warp / one MMA-issuing thread / four epilogue warps, two accumulators; the epilogue maps a tile row back to
// its output position (to, b, fo).
#include "gemm.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace ctcasr {
namespace convtc {

constexpr int BM = 128, BK = 32, NACC = 2, ACC_COLS = 128, NTHREADS = 192;
constexpr int A_PIECE = BM * BK * 2;            // 8 KB: 128 rows x 32 bf16
constexpr int B_BOX = BK * 128;                 // 4 KB: 32 k-rows x 64 filters (MN-major, SWIZZLE_128B)
constexpr int STG_LD = 36, STG_BYTES = 4 * 32 * STG_LD * 4;

template <int NP> struct Cfg {
    static constexpr int NPROD = NP == 3 ? 6 : (NP == 2 ? 3 : 1);
    static constexpr int STAGE = NP * (A_PIECE + 2 * B_BOX);                    // up to 128 filters: two boxes
    static constexpr int NSTAGE = NP == 3 ? 4 : (NP == 2 ? 6 : 8);              // 192 / 192 / 128 KB
    static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256 + STG_BYTES;
};

struct Params {
    int To, B, Fo, N, ldc;              // output [To, B, Fo, ldc], N real GEMM columns (multiple of 16, <= 128)
    int fob, bb, tb;                    // rows of a tile: tb x bb x fob = 128 (fo fastest)
    int fo_groups, b_groups, to_groups, num_tiles;
    int kt, kf, cchunks;                // taps and 32-channel chunks per tap: kt * kf * cchunks k-blocks
    int st, sf, pt, pf;
    int nbx;                            // 64-filter boxes of B per k-block (1 or 2)
    float *y;
    Epilogue epi;
};

__device__ __forceinline__ void decode_tile(const Params &p, int u, int &to0, int &b0, int &fo0)
{
    const int fg = u % p.fo_groups;
    const int r = u / p.fo_groups;
    const int bg = r % p.b_groups, tg = r / p.b_groups;
    to0 = tg * p.tb; b0 = bg * p.bb; fo0 = fg * p.fob;
}

template <int NP>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const Params p)
{
    using C_ = Cfg<NP>;
    constexpr int NSTAGE = C_::NSTAGE, STAGE = C_::STAGE;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + NSTAGE * STAGE;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + NACC + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 2 * NACC);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
    float *stg = reinterpret_cast<float *>(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (threadIdx.x >> 5 & 3) * 32 * STG_LD;
    auto a_addr = [&](int stage, int pc) { return smem_base + stage * STAGE + pc * A_PIECE; };
    auto b_addr = [&](int stage, int pc) { return smem_base + stage * STAGE + NP * A_PIECE + pc * 2 * B_BOX; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.kt * p.kf * p.cchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 4); }
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { ptx::tma_prefetch_desc(&mapX); ptx::tma_prefetch_desc(&mapW); }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, NACC * ACC_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        // ===== TMA producer: one tap x 32 channels of the 128 output positions + the kernel rows of that tap =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)NP * (A_PIECE + p.nbx * B_BOX);
            for (int u = blockIdx.x; u < p.num_tiles; u += gridDim.x) {
                int to0, b0, fo0;
                decode_tile(p, u, to0, b0, fo0);
                const int t0 = to0 * p.st - p.pt, f0 = fo0 * p.sf - p.pf;
                int kb = 0;
                for (int it = 0; it < p.kt; ++it)
                    for (int jf = 0; jf < p.kf; ++jf)
                        for (int cc = 0; cc < p.cchunks; ++cc, ++kb) {
                            ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                            ptx::mbar_expect_tx(full_bar(stage), bytes);
#pragma unroll
                            for (int pc = 0; pc < NP; ++pc) {
                                ptx::tma_load_5d(a_addr(stage, pc), &mapX, cc * BK, f0 + jf, b0, t0 + it, pc, full_bar(stage));
                                for (int j = 0; j < p.nbx; ++j)
                                    ptx::tma_load_3d(b_addr(stage, pc) + j * B_BOX, &mapW, j * 64, kb * BK, pc, full_bar(stage));
                            }
                            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr int PA[6] = {0, 0, 1, 0, 1, 2};
            constexpr int PB[6] = {0, 1, 0, 2, 1, 0};
            const uint32_t idesc = ptx::make_idesc_bf16(BM, p.N, 0, 1);        // A K-major, B MN-major
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < p.num_tiles; u += gridDim.x) {
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < kblocks; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int q = 0; q < C_::NPROD; ++q) {
                        const uint64_t adesc = ptx::make_smem_desc(a_addr(stage, PA[q]), 16, 512, 4);          // K-major, SWIZZLE_64B
                        const uint64_t bdesc = ptx::make_smem_desc(b_addr(stage, PB[q]), B_BOX, 1024, 2);      // MN-major, SWIZZLE_128B
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j)
                            ptx::mma_bf16(tmem_d, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(128 * j), idesc, (kb | q | j) != 0);
                    }
                    ptx::mma_commit(empty_bar(stage));
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                ptx::mma_commit(tfull_bar(acc));
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: tile row r = (tl * bb + bl) * fob + fol  ->  output row ((to0 + tl) * B + b0 + bl) * Fo + fo0 + fol =====
        int acc = 0; uint32_t acc_phase = 0;
        const int sub_n = 4 * (lane & 7), sub_r = lane >> 3;
        for (int u = blockIdx.x; u < p.num_tiles; u += gridDim.x) {
            int to0, b0, fo0;
            decode_tile(p, u, to0, b0, fo0);
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * ACC_COLS;
            // the output rows of the eight tile rows this lane stores (-1: beyond the tensor)
            long long orow[8];
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
                const int r = warp * 32 + itr * 4 + sub_r;
                const int fol = r % p.fob, q = r / p.fob;
                const int bl = q % p.bb, tl = q / p.bb;
                const int to = to0 + tl, b = b0 + bl, fo = fo0 + fol;
                orow[itr] = (to < p.To && b < p.B && fo < p.Fo) ? ((long long)to * p.B + b) * p.Fo + fo : -1;
            }
            const int nchunks = (p.N + 31) / 32;
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                uint32_t r[32];
                ptx::tmem_ld32(taddr + c * 32, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4 *>(stg + lane * STG_LD + 4 * q) =
                        make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                    __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                __syncwarp();
                const int n = c * 32 + sub_n;
                if (n < p.N) {
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        if (orow[itr] < 0) continue;
                        const int rr = itr * 4 + sub_r;
                        const float4 v = *reinterpret_cast<const float4 *>(stg + rr * STG_LD + sub_n);
                        const int m = (int)orow[itr];
                        float4 o;
                        o.x = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.x, m, n + 0, p.ldc, 0.f);
                        o.y = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.y, m, n + 1, p.ldc, 0.f);
                        o.z = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.z, m, n + 2, p.ldc, 0.f);
                        o.w = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.w, m, n + 3, p.ldc, 0.f);
                        *reinterpret_cast<float4 *>(p.y + (size_t)orow[itr] * p.ldc + n) = o;
                    }
                }
                __syncwarp();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, NACC * ACC_COLS);
}

// pad columns N .. ldc of the output (GEMM width > filter count: the first two layers): exact zeros, as conv.cu's GEMM
// writes them (zero kernel columns and bias)
__global__ void zero_pad_columns_kernel(float *y, size_t rows, int n0, int ldc)
{
    const int w = ldc - n0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * (size_t)w; i += (size_t)gridDim.x * blockDim.x)
        y[(i / w) * ldc + n0 + i % w] = 0.f;
}

template <int NP>
static int launch(const CUtensorMap &mx, const CUtensorMap &mw, const Params &p, cudaStream_t stream)
{
    using C_ = Cfg<NP>;
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        int dev = 0;
        CTCASR_CUDA_CHECK(cudaGetDevice(&dev));
        CTCASR_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(conv_fwd_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
        attr_set = true;
    }
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    ProfScope prof(PROF_GEMM_TC, stream);
    conv_fwd_kernel<NP><<<grid, NTHREADS, C_::SMEM, stream>>>(mx, mw, p);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}


// ================================ weight gradient: dW[K, N] = col^T dz without col ================================
// GEMM rows = patch elements k = (tap * C + c), columns = filters, contraction over the output positions.  An M tile is
// four consecutive 32-row (tap, channel-chunk) units; a k-block is 32 output positions (fob x bb x tb, fo fastest):
//   A  four boxes {32 channels, the 32 positions} of the input pieces, one per unit, each at its tap's offset — MN-major
//      A operand (a row of shared memory = one position, 64 B = 32 consecutive k), SWIZZLE_64B;
//   B  dz pieces viewed as [piece][To][B][Fo][N]: boxes {64 filters, fob, bb, tb} = the same 32 positions in the same
//      order, MN-major, SWIZZLE_128B.
// Few output tiles (58 for an 11 x 21 x 32 kernel) against 10^4 k-blocks: the positions are split into slices whose raw
// partial tiles splitk_reduce adds in slice order (deterministic).
constexpr int WG_BKP = 32;                          // positions per k-block
constexpr int WG_A_BOX = WG_BKP * 64;               // 2 KB: 32 positions x 32 channels
template <int NP> struct WCfg {
    static constexpr int NPROD = NP == 3 ? 6 : (NP == 2 ? 3 : 1);
    static constexpr int STAGE = NP * (4 * WG_A_BOX + 2 * B_BOX);               // 16 KB per piece
    static constexpr int NSTAGE = NP == 3 ? 4 : (NP == 2 ? 6 : 10);
    static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256 + STG_BYTES;
};
struct WParams {
    int N, K;                           // GEMM columns (multiple of 16, <= 128), rows (= kt*kf*C)
    int fob, bb, tb, fo_groups, b_groups, to_groups, ngroups;
    int kf, cchunks, nunits;            // nunits = kt*kf*cchunks 32-row units
    int st, sf, pt, pf;
    int nbx, mtiles, splits, groups_per_split;
    int groups_per_chunk;               // chained accumulation (gemm_tc.cu): position groups per accumulator
    float *out;                         // dW [K, ldo] (splits == 1) or the partial tiles [splits][K][ldo]
    int ldo;
};

template <int NP>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ, const WParams p)
{
    using C_ = WCfg<NP>;
    constexpr int NSTAGE = C_::NSTAGE, STAGE = C_::STAGE;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + NSTAGE * STAGE;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + NACC + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 2 * NACC);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
    float *stg = reinterpret_cast<float *>(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (threadIdx.x >> 5 & 3) * 32 * STG_LD;
    auto a_addr = [&](int stage, int pc) { return smem_base + stage * STAGE + pc * 4 * WG_A_BOX; };
    auto b_addr = [&](int stage, int pc) { return smem_base + stage * STAGE + NP * 4 * WG_A_BOX + pc * 2 * B_BOX; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwork = p.mtiles * p.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 4); }
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { ptx::tma_prefetch_desc(&mapX); ptx::tma_prefetch_desc(&mapZ); }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, NACC * ACC_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)NP * (4 * WG_A_BOX + p.nbx * B_BOX);
            for (int u = blockIdx.x; u < nwork; u += gridDim.x) {
                const int slice = u / p.mtiles, mt = u - slice * p.mtiles;
                // the four units of this tile: channel offset and tap offsets (a unit beyond the kernel reads far out of range: zeros)
                int c0[4], dt[4], df[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int un = mt * 4 + j;
                    const int tap = un / p.cchunks;
                    c0[j] = (un - tap * p.cchunks) * 32;
                    const int it = tap / p.kf;
                    dt[j] = un < p.nunits ? it - p.pt : (1 << 24);
                    df[j] = (tap - it * p.kf) - p.pf;
                }
                const int g0 = slice * p.groups_per_split, g1 = min(p.ngroups, g0 + p.groups_per_split);
                for (int g = g0; g < g1; ++g) {
                    const int fg = g % p.fo_groups, r = g / p.fo_groups;
                    const int bg = r % p.b_groups, tg = r / p.b_groups;
                    const int to0 = tg * p.tb, b0 = bg * p.bb, fo0 = fg * p.fob;
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    ptx::mbar_expect_tx(full_bar(stage), bytes);
#pragma unroll
                    for (int pc = 0; pc < NP; ++pc) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            ptx::tma_load_5d(a_addr(stage, pc) + j * WG_A_BOX, &mapX, c0[j], fo0 * p.sf + df[j], b0, to0 * p.st + dt[j], pc, full_bar(stage));
                        for (int j = 0; j < p.nbx; ++j)
                            ptx::tma_load_5d(b_addr(stage, pc) + j * B_BOX, &mapZ, j * 64, fo0, b0, to0, pc, full_bar(stage));
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr int PA[6] = {0, 0, 1, 0, 1, 2};
            constexpr int PB[6] = {0, 1, 0, 2, 1, 0};
            const uint32_t idesc = ptx::make_idesc_bf16(BM, p.N, 1, 1);        // both operands MN-major
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < nwork; u += gridDim.x) {
                const int slice = u / p.mtiles;
                const int g0 = slice * p.groups_per_split, g1 = min(p.ngroups, g0 + p.groups_per_split);
                for (int c0 = g0; c0 < g1; c0 += p.groups_per_chunk) {     // one accumulator per chunk of the chain
                    const int c1 = min(g1, c0 + p.groups_per_chunk);
                    ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                    for (int g = c0; g < c1; ++g) {
                        ptx::mbar_wait(full_bar(stage), phase);
                        ptx::tc_fence_after();
#pragma unroll
                        for (int q = 0; q < C_::NPROD; ++q) {
                            // A: MN-major, SWIZZLE_64B: 32-element atoms WG_A_BOX apart along M, 8-position groups 512 B apart
                            const uint64_t adesc = ptx::make_smem_desc(a_addr(stage, PA[q]), WG_A_BOX, 512, 4);
                            const uint64_t bdesc = ptx::make_smem_desc(b_addr(stage, PB[q]), B_BOX, 1024, 2);
#pragma unroll
                            for (int j = 0; j < WG_BKP / 16; ++j)       // 16 positions per MMA: 16 rows of 64 B (A) / 128 B (B)
                                ptx::mma_bf16(tmem_d, adesc + (uint64_t)(64 * j), bdesc + (uint64_t)(128 * j), idesc, ((g - c0) | q | j) != 0);
                        }
                        ptx::mma_commit(empty_bar(stage));
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    ptx::mma_commit(tfull_bar(acc));
                    if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        int acc = 0; uint32_t acc_phase = 0;
        const int sub_n = 4 * (lane & 7), sub_r = lane >> 3;
        for (int u = blockIdx.x; u < nwork; u += gridDim.x) {
            const int slice = u / p.mtiles, mt = u - slice * p.mtiles;
            float *out = p.out + (size_t)slice * p.K * p.ldo;
            const int g0 = slice * p.groups_per_split, g1 = min(p.ngroups, g0 + p.groups_per_split);
            // chunks of the chain: the first stores, the others add to what this thread stored (fp32, round to nearest)
            for (int c0 = g0; c0 < g1; c0 += p.groups_per_chunk) {
                const bool first = c0 == g0;
                ptx::mbar_wait(tfull_bar(acc), acc_phase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * ACC_COLS;
                const int m0 = mt * BM + warp * 32;
                const int nchunks = (p.N + 31) / 32;
#pragma unroll 1
                for (int c = 0; c < nchunks; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld32(taddr + c * 32, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4 *>(stg + lane * STG_LD + 4 * q) =
                            make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                        __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                    __syncwarp();
                    const int n = c * 32 + sub_n;
                    if (n < p.N) {
                        float4 old[8];                  // (all loads in flight before the first store, as in gemm_tc.cu)
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const int m = m0 + itr * 4 + sub_r;
                            old[itr] = (!first && m < p.K) ? *reinterpret_cast<const float4 *>(out + (size_t)m * p.ldo + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr) {
                            const int rr = itr * 4 + sub_r, m = m0 + rr;
                            if (m < p.K) {
                                const float4 v = *reinterpret_cast<const float4 *>(stg + rr * STG_LD + sub_n);
                                *reinterpret_cast<float4 *>(out + (size_t)m * p.ldo + n) =
                                    make_float4(v.x + old[itr].x, v.y + old[itr].y, v.z + old[itr].z, v.w + old[itr].w);
                            }
                        }
                    }
                    __syncwarp();
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, NACC * ACC_COLS);
}

template <int NP>
static int launch_wgrad(const CUtensorMap &mx, const CUtensorMap &mz, const WParams &p, cudaStream_t stream)
{
    using C_ = WCfg<NP>;
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        int dev = 0;
        CTCASR_CUDA_CHECK(cudaGetDevice(&dev));
        CTCASR_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
        attr_set = true;
    }
    const int nwork = p.mtiles * p.splits;
    const int grid = nwork < num_sms ? nwork : num_sms;
    ProfScope prof(PROF_GEMM_TC, stream);
    conv_wgrad_kernel<NP><<<grid, NTHREADS, C_::SMEM, stream>>>(mx, mz, p);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// ================================ input gradient: dx = conv_transpose(dz, W) without dcol ================================
// dx[t, b, f, c] = sum over taps (it, jf) and filters n of dz[t + pt - it, b, (f + pf - jf) / sf, n] W[(it*kf + jf)*C + c, n]
// (time stride 1), the frequency term only when sf divides f + pf - jf.  Input positions are processed per residue class
// rho = f mod sf: f = rho + sf*i sees the taps jf = j0 + sf*j' (j0 = (rho + pf) mod sf) at fo = i + q0 - j', consecutive in i.
// GEMM rows = 128 input positions of one class (ib x bb x tb, i fastest), columns = the C input channels, one k-block =
// one tap x 32 filters:
//   A  dz pieces [piece][To][B][Fo][N]: box {32 filters, ib, bb, tb} at (n0, i0 + q0 - j', b0, t0 + pt - it): K-major, SWIZZLE_64B;
//      positions whose (to, fo) fall outside the layer's output read zeros (TMA out-of-range fill)
//   B  kernel pieces [piece][Kp][N]: box {32 filters, the tap's C rows}: K-major, SWIZZLE_64B.
struct DParams {
    int T, B, F, C, ldx;                // dx [T, B, F, ldx], C real channels (multiple of 32, <= 128)
    int ib, bb, tb, i_groups, b_groups, t_groups, num_tiles;   // tiles: class x t x b x i
    int kt, kf, sf, pt, pf, nchunks;    // nchunks = 32-filter chunks of the GEMM's contraction per tap
    float *dx;
};
template <int NP> struct DCfg {
    static constexpr int NPROD = NP == 3 ? 6 : (NP == 2 ? 3 : 1);
    static constexpr int STAGE = NP * (A_PIECE + A_PIECE);                      // A 8 KB + B up to 128 rows x 64 B
    static constexpr int NSTAGE = NP == 3 ? 4 : (NP == 2 ? 6 : 10);
    static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256 + STG_BYTES;
};
__device__ __forceinline__ void dgrad_tile(const DParams &p, int u, int &rho, int &t0, int &b0, int &i0, int &j0, int &q0, int &nj)
{
    const int ig = u % p.i_groups; int r = u / p.i_groups;
    const int bg = r % p.b_groups; r /= p.b_groups;
    const int tg = r % p.t_groups; rho = r / p.t_groups;
    t0 = tg * p.tb; b0 = bg * p.bb; i0 = ig * p.ib;
    j0 = (rho + p.pf) % p.sf;
    q0 = (rho + p.pf - j0) / p.sf;
    nj = j0 < p.kf ? (p.kf - 1 - j0) / p.sf + 1 : 0;
}

template <int NP>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_dgrad_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapW, const DParams p)
{
    using C_ = DCfg<NP>;
    constexpr int NSTAGE = C_::NSTAGE, STAGE = C_::STAGE;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + NSTAGE * STAGE;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + NACC + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 2 * NACC);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
    float *stg = reinterpret_cast<float *>(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (threadIdx.x >> 5 & 3) * 32 * STG_LD;
    auto a_addr = [&](int stage, int pc) { return smem_base + stage * STAGE + pc * A_PIECE; };
    auto b_addr = [&](int stage, int pc) { return smem_base + stage * STAGE + NP * A_PIECE + pc * A_PIECE; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 4); }
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { ptx::tma_prefetch_desc(&mapZ); ptx::tma_prefetch_desc(&mapW); }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, NACC * ACC_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)NP * (A_PIECE + p.C * 64);
            for (int u = blockIdx.x; u < p.num_tiles; u += gridDim.x) {
                int rho, t0, b0, i0, j0, q0, nj;
                dgrad_tile(p, u, rho, t0, b0, i0, j0, q0, nj);
                for (int it = 0; it < p.kt; ++it)
                    for (int jj = 0; jj < nj; ++jj) {
                        const int jf = j0 + p.sf * jj;
                        for (int nc = 0; nc < p.nchunks; ++nc) {
                            ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                            ptx::mbar_expect_tx(full_bar(stage), bytes);
#pragma unroll
                            for (int pc = 0; pc < NP; ++pc) {
                                ptx::tma_load_5d(a_addr(stage, pc), &mapZ, nc * 32, i0 + q0 - jj, b0, t0 + p.pt - it, pc, full_bar(stage));
                                ptx::tma_load_3d(b_addr(stage, pc), &mapW, nc * 32, (it * p.kf + jf) * p.C, pc, full_bar(stage));
                            }
                            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        }
                    }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr int PA[6] = {0, 0, 1, 0, 1, 2};
            constexpr int PB[6] = {0, 1, 0, 2, 1, 0};
            const uint32_t idesc = ptx::make_idesc_bf16(BM, p.C, 0, 0);        // both operands K-major
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < p.num_tiles; u += gridDim.x) {
                int rho, t0, b0, i0, j0, q0, nj;
                dgrad_tile(p, u, rho, t0, b0, i0, j0, q0, nj);
                const int nk = p.kt * nj * p.nchunks;
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < nk; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int q = 0; q < C_::NPROD; ++q) {
                        const uint64_t adesc = ptx::make_smem_desc(a_addr(stage, PA[q]), 16, 512, 4);
                        const uint64_t bdesc = ptx::make_smem_desc(b_addr(stage, PB[q]), 16, 512, 4);
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j)
                            ptx::mma_bf16(tmem_d, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), idesc, (kb | q | j) != 0);
                    }
                    ptx::mma_commit(empty_bar(stage));
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                ptx::mma_commit(tfull_bar(acc));
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        int acc = 0; uint32_t acc_phase = 0;
        const int sub_n = 4 * (lane & 7), sub_r = lane >> 3;
        for (int u = blockIdx.x; u < p.num_tiles; u += gridDim.x) {
            int rho, t0, b0, i0, j0, q0, nj;
            dgrad_tile(p, u, rho, t0, b0, i0, j0, q0, nj);
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * ACC_COLS;
            long long orow[8];
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
                const int r = warp * 32 + itr * 4 + sub_r;
                const int il = r % p.ib, q = r / p.ib;
                const int bl = q % p.bb, tl = q / p.bb;
                const int t = t0 + tl, b = b0 + bl, f = rho + p.sf * (i0 + il);
                orow[itr] = (t < p.T && b < p.B && f < p.F) ? ((long long)t * p.B + b) * p.F + f : -1;
            }
            const int nchunks = (p.ldx + 31) / 32;              // the pad channels get their zeros here
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                const bool real = c * 32 < p.C;
                if (real) {
                    uint32_t r[32];
                    ptx::tmem_ld32(taddr + c * 32, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4 *>(stg + lane * STG_LD + 4 * q) =
                            make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                        __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                }
                __syncwarp();
                const int n = c * 32 + sub_n;
                if (n < p.ldx) {
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        if (orow[itr] < 0) continue;
                        const int rr = itr * 4 + sub_r;
                        const float4 v = (real && nj > 0) ? *reinterpret_cast<const float4 *>(stg + rr * STG_LD + sub_n) : make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4 *>(p.dx + (size_t)orow[itr] * p.ldx + n) = v;
                    }
                }
                __syncwarp();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, NACC * ACC_COLS);
}

template <int NP>
static int launch_dgrad(const CUtensorMap &mz, const CUtensorMap &mw, const DParams &p, cudaStream_t stream)
{
    using C_ = DCfg<NP>;
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        int dev = 0;
        CTCASR_CUDA_CHECK(cudaGetDevice(&dev));
        CTCASR_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(conv_dgrad_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
        attr_set = true;
    }
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    ProfScope prof(PROF_GEMM_TC, stream);
    conv_dgrad_kernel<NP><<<grid, NTHREADS, C_::SMEM, stream>>>(mz, mw, p);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

// the 128 = tb x bb x fob (fwd) / 32 (wgrad) rows of a box: powers of two wasting the fewest rows on partial boxes
static void pick_boxes(int rows, int To, int B, int Fo, int st, int sf, int *fob_, int *bb_, int *tb_)
{
    double best = 1e30;
    for (int fob = 1; fob <= rows; fob *= 2)
        for (int bb = 1; fob * bb <= rows; bb *= 2) {
            const int tb = rows / (fob * bb);
            if (fob * sf > 256 || tb * st > 256) continue;                      // TMA box extents
            const double waste = (double)((Fo + fob - 1) / fob * fob) / Fo * ((B + bb - 1) / bb * bb) / B * ((To + tb - 1) / tb * tb) / To;
            if (waste < best - 1e-9) { best = waste; *fob_ = fob; *bb_ = bb; *tb_ = tb; }
        }
}

}  // namespace convtc

// Which layers take the implicit kernel: a bf16 compute mode, whole 32-channel chunks, a GEMM width the accumulator and
// the two B boxes hold, an activation layout whose rows are 16-B aligned for the TMA unit.  CTCASR_CONV_IMPLICIT=0 keeps
// the patch-matrix path (conv.cu) for comparisons.
bool conv_tc_eligible(int compute, int T, int B, int F, int C, int kt, int kf, int st, int sf, int N)
{
    static const bool enabled = !(getenv("CTCASR_CONV_IMPLICIT") && atoi(getenv("CTCASR_CONV_IMPLICIT")) == 0);
    if (!enabled) return false;
    if (compute != CTCASR_COMPUTE_BF16X3 && compute != CTCASR_COMPUTE_BF16) return false;
    if (C < 32 || C % 32) return false;
    if (N < 16 || N > 128 || N % 16) return false;
    if (st < 1 || st > 8 || sf < 1 || sf > 8) return false;                     // TMA traversal strides
    if (kt > 64 || kf > 64) return false;
    return (double)T * B * F * kt * kf * C * N / (st * sf) >= 4.0e6;
}

// y [To, B, Fo, ldc] = dropout(act(conv(x, w) + bias)); x [T, B, F, x_pitch] fp32 (C real channels), w [Kp, ldw] fp32 with
// N_real = the GEMM columns actually computed.  Must be called inside an open split scope (gemm.cuh).
int conv_tc_fwd(const float *x, int x_pitch, const float *w, int ldw, const float *bias, float *y, int ldc,
                int T, int B, int F, int C, int kt, int kf, int st, int sf, int To, int Fo, int pt, int pf, int N_real,
                int np, int act, float cutoff, float drop_rate, uint32_t seed, cudaStream_t stream)
{
    using namespace convtc;
    const int K = kt * kf * C, Kp = (K + 7) / 8 * 8;
    // operand pieces: the input with its pad channels dropped ([T*B*F rows][C]), the kernel as the dense layers split theirs
    const __nv_bfloat16 *xs = nullptr, *ws = nullptr;
    int ldx = 0, ldws = 0;
    size_t xpiece = 0, wpiece = 0;
    if (int rc = gemm_tc_pieces(x, T * B * F, C, x_pitch, np, stream, &xs, &ldx, &xpiece)) return rc;
    if (int rc = gemm_tc_pieces(w, Kp, ldw, ldw, np, stream, &ws, &ldws, &wpiece)) return rc;

    Params p{};
    p.To = To; p.B = B; p.Fo = Fo; p.N = (N_real + 15) / 16 * 16; p.ldc = ldc;
    pick_boxes(128, To, B, Fo, st, sf, &p.fob, &p.bb, &p.tb);       // tile rows = tb x bb x fob
    p.fo_groups = (Fo + p.fob - 1) / p.fob; p.b_groups = (B + p.bb - 1) / p.bb; p.to_groups = (To + p.tb - 1) / p.tb;
    p.num_tiles = p.fo_groups * p.b_groups * p.to_groups;
    p.kt = kt; p.kf = kf; p.cchunks = C / 32; p.st = st; p.sf = sf; p.pt = pt; p.pf = pf;
    p.nbx = (p.N + 63) / 64;
    p.y = y;
    p.epi.mode = EPI_BIAS_ACT; p.epi.bias = bias; p.epi.act = act; p.epi.cutoff = cutoff; p.epi.drop_rate = drop_rate; p.epi.seed = seed;

    CUtensorMap mx, mw;
    {   // input pieces [np][T][B][F][ldx]: box {32 c, fob (stride sf), bb, tb (stride st), 1}
        const unsigned long long dims[5] = {(unsigned long long)C, (unsigned long long)F, (unsigned long long)B, (unsigned long long)T, (unsigned long long)np};
        const unsigned long long strides[4] = {(unsigned long long)ldx * 2, (unsigned long long)F * ldx * 2, (unsigned long long)B * F * ldx * 2, (unsigned long long)xpiece * 2};
        const unsigned box[5] = {32u, (unsigned)(p.fob * sf), (unsigned)p.bb, (unsigned)(p.tb * st), 1u};
        const unsigned estr[5] = {1u, (unsigned)sf, 1u, (unsigned)st, 1u};
        if (int rc = tma_encode_bf16(&mx, xs, 5, dims, strides, box, estr, 64)) return rc;
    }
    {   // kernel pieces [np][Kp][ldws]: MN-major boxes {64 filters, 32 k-rows, 1}
        const unsigned long long dims[3] = {(unsigned long long)ldw, (unsigned long long)Kp, (unsigned long long)np};
        const unsigned long long strides[2] = {(unsigned long long)ldws * 2, (unsigned long long)wpiece * 2};
        const unsigned box[3] = {64u, 32u, 1u};
        const unsigned estr[3] = {1u, 1u, 1u};
        if (int rc = tma_encode_bf16(&mw, ws, 3, dims, strides, box, estr, 128)) return rc;
    }
    int rc;
    if (np == 3) rc = launch<3>(mx, mw, p, stream);
    else if (np == 2) rc = launch<2>(mx, mw, p, stream);
    else rc = launch<1>(mx, mw, p, stream);
    if (rc != CTCASR_OK) return rc;
    if (p.N < ldc) {
        const size_t rows = (size_t)To * B * Fo;
        zero_pad_columns_kernel<<<148 * 4, 256, 0, stream>>>(y, rows, p.N, ldc);
        CTCASR_LAUNCH_CHECK();
    }
    return CTCASR_OK;
}

// dW [K = kt*kf*C, ldw] = col^T dz for dz [To*B*Fo, ldz] fp32 (the layer's masked output gradient); inside an open split scope
int conv_tc_wgrad(const float *x, int x_pitch, const float *dz, int ldz, float *dw, int ldw,
                  int T, int B, int F, int C, int kt, int kf, int st, int sf, int To, int Fo, int pt, int pf,
                  int np, cudaStream_t stream)
{
    using namespace convtc;
    const int K = kt * kf * C;
    const size_t rows = (size_t)To * B * Fo;
    const __nv_bfloat16 *xs = nullptr, *zs = nullptr;
    int ldx = 0, ldzs = 0;
    size_t xpiece = 0, zpiece = 0;
    if (int rc = gemm_tc_pieces(x, T * B * F, C, x_pitch, np, stream, &xs, &ldx, &xpiece)) return rc;
    if (int rc = gemm_tc_pieces(dz, (int)rows, ldz, ldz, np, stream, &zs, &ldzs, &zpiece)) return rc;

    WParams p{};
    p.N = (ldz + 15) / 16 * 16; p.K = K; p.ldo = ldw;
    pick_boxes(WG_BKP, To, B, Fo, st, sf, &p.fob, &p.bb, &p.tb);
    p.fo_groups = (Fo + p.fob - 1) / p.fob; p.b_groups = (B + p.bb - 1) / p.bb; p.to_groups = (To + p.tb - 1) / p.tb;
    p.ngroups = p.fo_groups * p.b_groups * p.to_groups;
    p.kf = kf; p.cchunks = C / 32; p.nunits = kt * kf * p.cchunks;
    p.st = st; p.sf = sf; p.pt = pt; p.pf = pf;
    p.nbx = (p.N + 63) / 64;
    p.mtiles = (p.nunits + 3) / 4;
    // ~2 waves of (slice, tile) units, at least 64 position groups per slice; the partial tiles go behind the cached splits
    int S = (2 * 148) / p.mtiles;
    if (S > p.ngroups / 64) S = p.ngroups / 64;
    float *part = nullptr;
    if (S > 1) {
        const int per = (p.ngroups + S - 1) / S;
        S = (p.ngroups + per - 1) / per;
        part = S > 1 ? reinterpret_cast<float *>(scratch_free((size_t)S * K * ldw * sizeof(float))) : nullptr;
        if (part) { p.splits = S; p.groups_per_split = per; }
    }
    if (!part) { p.splits = 1; p.groups_per_split = p.ngroups; }
    p.groups_per_chunk = np >= 2 ? chain_chunk_kblocks(p.groups_per_split, WG_BKP) : p.groups_per_split;   // (bf16x3 arithmetic only)
    p.out = part ? part : dw;

    CUtensorMap mx, mz;
    {
        const unsigned long long dims[5] = {(unsigned long long)C, (unsigned long long)F, (unsigned long long)B, (unsigned long long)T, (unsigned long long)np};
        const unsigned long long strides[4] = {(unsigned long long)ldx * 2, (unsigned long long)F * ldx * 2, (unsigned long long)B * F * ldx * 2, (unsigned long long)xpiece * 2};
        const unsigned box[5] = {32u, (unsigned)(p.fob * sf), (unsigned)p.bb, (unsigned)(p.tb * st), 1u};
        const unsigned estr[5] = {1u, (unsigned)sf, 1u, (unsigned)st, 1u};
        if (int rc = tma_encode_bf16(&mx, xs, 5, dims, strides, box, estr, 64)) return rc;
    }
    {   // dz pieces [np][To][B][Fo][ldzs]: boxes {64 filters, fob, bb, tb, 1}
        const unsigned long long dims[5] = {(unsigned long long)ldz, (unsigned long long)Fo, (unsigned long long)B, (unsigned long long)To, (unsigned long long)np};
        const unsigned long long strides[4] = {(unsigned long long)ldzs * 2, (unsigned long long)Fo * ldzs * 2, (unsigned long long)B * Fo * ldzs * 2, (unsigned long long)zpiece * 2};
        const unsigned box[5] = {64u, (unsigned)p.fob, (unsigned)p.bb, (unsigned)p.tb, 1u};
        const unsigned estr[5] = {1u, 1u, 1u, 1u, 1u};
        if (int rc = tma_encode_bf16(&mz, zs, 5, dims, strides, box, estr, 128)) return rc;
    }
    int rc;
    if (np == 3) rc = launch_wgrad<3>(mx, mz, p, stream);
    else if (np == 2) rc = launch_wgrad<2>(mx, mz, p, stream);
    else rc = launch_wgrad<1>(mx, mz, p, stream);
    if (rc != CTCASR_OK) return rc;
    if (p.splits > 1) {
        GemmArgs g;
        g.M = K; g.N = ldw; g.C[0] = dw; g.ldc = ldw;
        return splitk_reduce(g, p.splits, part, stream);
    }
    return CTCASR_OK;
}

// dx [T, B, F, x_pitch] = conv_transpose(dz [To, B, Fo, ldz], w [Kp, ldw]) for time stride 1; inside an open split scope
bool conv_tc_dgrad_eligible(int C, int kf, int st, int sf, int x_pitch)
{
    static const bool enabled = !(getenv("CTCASR_CONV_IMPLICIT_DGRAD") && atoi(getenv("CTCASR_CONV_IMPLICIT_DGRAD")) == 0);
    return enabled && st == 1 && kf >= sf && C <= 128 && x_pitch % 4 == 0;
}
int conv_tc_dgrad(const float *dz, int ldz, const float *w, int ldw, float *dx, int x_pitch,
                  int T, int B, int F, int C, int kt, int kf, int st, int sf, int To, int Fo, int pt, int pf,
                  int np, cudaStream_t stream)
{
    using namespace convtc;
    (void)st;
    const int K = kt * kf * C, Kp = (K + 7) / 8 * 8;
    const size_t rows = (size_t)To * B * Fo;
    const __nv_bfloat16 *zs = nullptr, *ws = nullptr;
    int ldzs = 0, ldws = 0;
    size_t zpiece = 0, wpiece = 0;
    if (int rc = gemm_tc_pieces(dz, (int)rows, ldz, ldz, np, stream, &zs, &ldzs, &zpiece)) return rc;
    if (int rc = gemm_tc_pieces(w, Kp, ldw, ldw, np, stream, &ws, &ldws, &wpiece)) return rc;

    DParams p{};
    p.T = T; p.B = B; p.F = F; p.C = C; p.ldx = x_pitch; p.dx = dx;
    const int imax = (F + sf - 1) / sf;                         // positions of a residue class (the longest one)
    pick_boxes(128, T, B, imax, 1, 1, &p.ib, &p.bb, &p.tb);
    p.i_groups = (imax + p.ib - 1) / p.ib; p.b_groups = (B + p.bb - 1) / p.bb; p.t_groups = (T + p.tb - 1) / p.tb;
    p.num_tiles = p.i_groups * p.b_groups * p.t_groups * sf;
    p.kt = kt; p.kf = kf; p.sf = sf; p.pt = pt; p.pf = pf; p.nchunks = (ldz + 31) / 32;

    CUtensorMap mz, mw;
    {   // dz pieces [np][To][B][Fo][ldzs]: boxes {32 filters, ib, bb, tb, 1}
        const unsigned long long dims[5] = {(unsigned long long)ldz, (unsigned long long)Fo, (unsigned long long)B, (unsigned long long)To, (unsigned long long)np};
        const unsigned long long strides[4] = {(unsigned long long)ldzs * 2, (unsigned long long)Fo * ldzs * 2, (unsigned long long)B * Fo * ldzs * 2, (unsigned long long)zpiece * 2};
        const unsigned box[5] = {32u, (unsigned)p.ib, (unsigned)p.bb, (unsigned)p.tb, 1u};
        const unsigned estr[5] = {1u, 1u, 1u, 1u, 1u};
        if (int rc = tma_encode_bf16(&mz, zs, 5, dims, strides, box, estr, 64)) return rc;
    }
    {   // kernel pieces [np][Kp][ldws]: K-major boxes {32 filters, C rows of one tap, 1}
        const unsigned long long dims[3] = {(unsigned long long)ldw, (unsigned long long)Kp, (unsigned long long)np};
        const unsigned long long strides[2] = {(unsigned long long)ldws * 2, (unsigned long long)wpiece * 2};
        const unsigned box[3] = {32u, (unsigned)C, 1u};
        const unsigned estr[3] = {1u, 1u, 1u};
        if (int rc = tma_encode_bf16(&mw, ws, 3, dims, strides, box, estr, 64)) return rc;
    }
    if (np == 3) return launch_dgrad<3>(mz, mw, p, stream);
    if (np == 2) return launch_dgrad<2>(mz, mw, p, stream);
    return launch_dgrad<1>(mz, mw, p, stream);
}

}  // namespace ctcasr
