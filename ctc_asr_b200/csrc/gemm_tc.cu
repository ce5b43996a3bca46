// gemm_tc.cu — tcgen05 / TMA / TMEM GEMM for sm_100a (CTCASR_COMPUTE_TF32, CTCASR_COMPUTE_BF16X3, CTCASR_COMPUTE_BF16):
// the single-CTA kernel gemm_tc_kernel and its CTA-pair (cta_group::2) variant gemm_tc_pair_kernel.
//
// C[M,N] = op(A) op(B), fp32 in HBM, fp32 accumulation in tensor memory.  Used for every GEMM-shaped
// piece of the path: the dense layers (asr/util/tf_contrib.py:52-58, asr/model.py:220-232), the
// hoisted RNN input projection and the backward GEMMs (dgrad, wgrad) — in the reference these are
// cuBLAS sgemm calls issued by TensorFlow.
//
// Arithmetic modes (template parameter MODE):
//   TF32    tcgen05.mma kind::tf32 straight on the fp32 operands (the TMA unit rounds fp32 -> tf32 on
//           the way into shared memory).  Fastest; 2^-11 operand rounding (3e-4 of max per GEMM), which
//           flips ReLU masks in the dense stack and costs up to 3e-2 of max-norm gradient error.
//   BF16X3  fp32-accurate emulation: a pre-pass splits each operand into bf16 pieces a = a1 + a2
//           (+ a3); the kernel issues kind::f16 MMAs for a1b1 + a1b2 + a2b1 (3 products, error
//           ~2^-16) or, for layers whose output goes through a ReLU kink (`precise`), the 6 products
//           of a 3-piece split (error ~2^-23) — all accumulated in the same fp32 TMEM tile.
//           Same bytes per element as fp32 (2 x 2 B), 1.5x the tensor time of TF32.
//   BF16    one bf16 piece per operand, one product (CTCASR_COMPUTE_BF16): plain bf16 tensor-core arithmetic
//           with fp32 accumulation — BASELINE cfg3's "bf16" arithmetic, 2^-9 operand rounding, 3x fewer MMAs.
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 4      TMA producer: cp.async.bulk.tensor boxes into a shared-memory ring
//               (per stage and piece: A 128 x 32, B 256 x 32), mbarrier expect_tx
//   warp 5      MMA issuer: one elected lane issues the tcgen05.mma's of a stage (128 x 256 x 8|16
//               each); tcgen05.commit releases the stage / publishes the accumulator
//   warps 0-3   epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / clipped ReLU / dropout /
//               activation mask / accumulate -> 128-bit global stores.  Two 256-column TMEM
//               accumulators, so the epilogue of tile i overlaps the main loop of tile i+1.
// Both operand orientations are handled in the descriptors, not by transposing data:
//   K-major  (A[m][k], B[n][k]): one box [rows x 32 k] per piece
//              tf32: SWIZZLE_128B, SBO 1024          bf16: SWIZZLE_64B, SBO 512
//   MN-major (A[k][m], B[k][n]): boxes [32 k x 128 B of m|n], 4 KB apart (LBO)
//              tf32: TMA 128B_ATOM_32B / desc SWIZZLE_128B_BASE32B, SBO 512 (the only legal layout
//                    for 32-bit MN-major operands)   bf16: SWIZZLE_128B, SBO 1024
// Tiles are walked in groups of 16 row-tiles x all column-tiles so that the ~148 tiles in flight
// share A and B panels through L2.
#include "gemm.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <mutex>
#include <stdlib.h>

namespace ctcasr {
namespace tc {

constexpr int BM = 128, BN = 256;
constexpr int NACC = 2;
constexpr int GROUP_M = 16;
constexpr int NTHREADS = 192;
constexpr int STG_LD = 36;                                      // epilogue staging tile: 32 rows x 36 floats per warp
constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;

enum { MODE_TF32 = 1, MODE_BF16X3 = 2, MODE_BF16X6 = 3, MODE_BF16X1 = 4 };   // X3 / X6: value = pieces per operand

template <int MODE>
struct Cfg {
    static constexpr bool kBf16 = MODE != MODE_TF32;
    static constexpr int NP = (kBf16 && MODE != MODE_BF16X1) ? MODE : 1;     // pieces per operand
    static constexpr int ESZ = kBf16 ? 2 : 4;
    // k-block of a pipeline stage.  One product per operand pair (BF16X1) takes 64 k (128-B rows, SWIZZLE_128B): with
    // 32 k a stage is two 128-cycle MMAs, and the issuing thread's barrier wait + commit and the producer's TMA
    // instructions (up to 6 boxes of 4 KB per stage) cost as much as those MMAs run — measured 830-1100 TFLOP/s
    // depending only on the number of boxes per stage.  The multi-product modes do 3-6x the MMAs per stage.
    static constexpr int BK = MODE == MODE_BF16X1 ? 64 : 32;
    static constexpr int A_PIECE = BM * BK * ESZ, B_PIECE = BN * BK * ESZ;
    static constexpr int STAGE_BYTES = NP * (A_PIECE + B_PIECE);
    static constexpr int NSTAGE = MODE == MODE_BF16X6 ? 2 : 4;      // 192 KB (BF16X6: 144 KB)
    static constexpr int UMMA_K = kBf16 ? 16 : 8;
    static constexpr int KSTEPS = BK / UMMA_K;
    static constexpr int MN_BOX = 128 / ESZ;                    // m|n elements per 128-B row
    static constexpr int MN_BOX_BYTES = BK * 128;               // one MN-major box: BK k-rows x 128 B
    static constexpr int NPROD = (MODE == MODE_TF32 || MODE == MODE_BF16X1) ? 1 : (MODE == MODE_BF16X3 ? 3 : 6);
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + STG_BYTES;
    // per k-step advance of the descriptor start address (bytes)
    static constexpr int KMAJ_STEP = UMMA_K * ESZ;              // inside the swizzled row
    static constexpr int MNMAJ_STEP = UMMA_K * 128;             // UMMA_K rows of 128 B
};

struct Params {
    int M, N, K, nz, ta, tb, ldc;
    float *C[2];
    Epilogue epi;
    int tiles_m, tiles_n, kblocks, num_tiles;
    // split-K (few output tiles, long contraction): work unit u = slice * num_tiles + tile; slice s covers
    // k-blocks [s * kb_per_split, ...) and stores its raw partial tile to part[s][z][M][N]
    int splits, kb_per_split;
    float *part;
    // chained accumulation: the k-blocks of a unit are summed in chunks of kb_per_chunk, one tensor-memory accumulator per
    // chunk, and the epilogue adds the chunks in fp32 (round to nearest) into the output it wrote for the first one.  The
    // tensor core truncates every accumulation (measured bias -2.4e-8 of the accumulator per MMA, profiles/
    // r2_accum_error.json): the error of ONE chain grows linearly with its length (1.2e-4 at K = 32,000), the chunks bound it
    int kb_per_chunk;
};

__device__ __forceinline__ void decode_tile(const Params &p, int t, int &z, int &mb, int &nb)
{
    const int per_z = p.tiles_m * p.tiles_n;
    z = t / per_z;
    t -= z * per_z;
    const int per_group = GROUP_M * p.tiles_n;
    const int g = t / per_group;
    const int first_m = g * GROUP_M;
    const int gm = min(GROUP_M, p.tiles_m - first_m);
    const int r = t - g * per_group;
    mb = first_m + r % gm;
    nb = r / gm;
}

template <int MODE>
__device__ __forceinline__ uint64_t operand_desc(uint32_t addr, bool mn_major)
{
    if (MODE == MODE_TF32)
        return mn_major ? ptx::make_smem_desc(addr, Cfg<MODE>::MN_BOX_BYTES, 512, 1)     // SWIZZLE_128B_BASE32B
                        : ptx::make_smem_desc(addr, 16, 1024, 2);                        // SWIZZLE_128B
    if (!mn_major && Cfg<MODE>::BK == 64) return ptx::make_smem_desc(addr, 16, 1024, 2);   // bf16, 128-B rows: SWIZZLE_128B
    return mn_major ? ptx::make_smem_desc(addr, Cfg<MODE>::MN_BOX_BYTES, 1024, 2)        // SWIZZLE_128B
                    : ptx::make_smem_desc(addr, 16, 512, 4);                             // SWIZZLE_64B
}

// Epilogue store of one 32 x 32 accumulator block from its staging tile: this lane owns columns
// n .. n+3 of rows sub_r, sub_r + 4, ...
enum { EK_RAW = 0, EK_ACC = 1, EK_BIAS_ACT = 2, EK_MASK = 3 };
template <int EK>
__device__ __forceinline__ void store_rows(const Params &p, const float *stg, float *base, int ld, int m0, int n, int sub_r, int sub_n)
{
    if (EK == EK_ACC) {
        // all eight loads of the old values in flight before the first store: with the load inside the store loop they
        // serialise behind each other (the compiler cannot reorder them across stores that may alias), eight L2 / HBM round
        // trips per 32-column block — longer than the MMAs of a chunk of a chained accumulation
        float4 old[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int m = m0 + it * 4 + sub_r;
            old[it] = m < p.M ? *reinterpret_cast<const float4 *>(base + (size_t)m * ld + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + sub_r, m = m0 + rr;
            if (m >= p.M) break;
            const float4 v = *reinterpret_cast<const float4 *>(stg + rr * STG_LD + sub_n);
            *reinterpret_cast<float4 *>(base + (size_t)m * ld + n) = make_float4(v.x + old[it].x, v.y + old[it].y, v.z + old[it].z, v.w + old[it].w);
        }
    }
#pragma unroll
    for (int it = 0; it < (EK == EK_ACC ? 0 : 8); ++it) {
        const int rr = it * 4 + sub_r, m = m0 + rr;
        if (m >= p.M) break;
        const float4 v = *reinterpret_cast<const float4 *>(stg + rr * STG_LD + sub_n);
        float *cp = base + (size_t)m * ld + n;
        float4 o = v;
        if (EK == EK_BIAS_ACT) {
            o.x = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.x, m, n + 0, p.N, 0.f);
            o.y = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.y, m, n + 1, p.N, 0.f);
            o.z = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.z, m, n + 2, p.N, 0.f);
            o.w = epilogue_apply_m<EPI_BIAS_ACT>(p.epi, v.w, m, n + 3, p.N, 0.f);
        } else if (EK == EK_MASK) {
            o.x = epilogue_apply_m<EPI_MASK>(p.epi, v.x, m, n + 0, p.N, 0.f);
            o.y = epilogue_apply_m<EPI_MASK>(p.epi, v.y, m, n + 1, p.N, 0.f);
            o.z = epilogue_apply_m<EPI_MASK>(p.epi, v.z, m, n + 2, p.N, 0.f);
            o.w = epilogue_apply_m<EPI_MASK>(p.epi, v.w, m, n + 3, p.N, 0.f);
        }
        *reinterpret_cast<float4 *>(cp) = o;
    }
}

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const Params p)
{
    using C_ = Cfg<MODE>;
    constexpr int NSTAGE = C_::NSTAGE, STAGE_BYTES = C_::STAGE_BYTES, NP = C_::NP;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle atoms: 1024-B aligned
    const uint32_t bar_base = smem_base + NSTAGE * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + NACC + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 2 * NACC);
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
    float *stg = reinterpret_cast<float *>(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (threadIdx.x >> 5 & 3) * 32 * STG_LD;
    // stage layout: A pieces, then B pieces
    auto a_addr = [&](int stage, int piece) { return smem_base + stage * STAGE_BYTES + piece * C_::A_PIECE; };
    auto b_addr = [&](int stage, int piece) { return smem_base + stage * STAGE_BYTES + NP * C_::A_PIECE + piece * C_::B_PIECE; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 4); }
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) {
        ptx::tma_prefetch_desc(&mapA0); ptx::tma_prefetch_desc(&mapB0);
        if (p.nz > 1) { ptx::tma_prefetch_desc(&mapA1); ptx::tma_prefetch_desc(&mapB1); }
    }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, NACC * BN);       // 512 columns: two fp32 128x256 accumulators
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        // ===================================== TMA producer ======================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < p.num_tiles * p.splits; u += gridDim.x) {
                int z, mb, nb;
                const int slice = u / p.num_tiles;
                decode_tile(p, u - slice * p.num_tiles, z, mb, nb);
                const CUtensorMap *ma = z ? &mapA1 : &mapA0;
                const CUtensorMap *mbp = z ? &mapB1 : &mapB0;
                const int kb0 = slice * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    ptx::mbar_expect_tx(full_bar(stage), STAGE_BYTES);
#pragma unroll
                    for (int pc = 0; pc < NP; ++pc) {
                        if (!p.ta) {
                            ptx::tma_load_3d(a_addr(stage, pc), ma, kb * C_::BK, mb * BM, pc, full_bar(stage));
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / C_::MN_BOX; ++j)
                                ptx::tma_load_3d(a_addr(stage, pc) + j * C_::MN_BOX_BYTES, ma, mb * BM + j * C_::MN_BOX,
                                                 kb * C_::BK, pc, full_bar(stage));
                        }
                        if (p.tb) {
                            ptx::tma_load_3d(b_addr(stage, pc), mbp, kb * C_::BK, nb * BN, pc, full_bar(stage));
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / C_::MN_BOX; ++j)
                                ptx::tma_load_3d(b_addr(stage, pc) + j * C_::MN_BOX_BYTES, mbp, nb * BN + j * C_::MN_BOX,
                                                 kb * C_::BK, pc, full_bar(stage));
                        }
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ====================================== MMA issuer =======================================
        if (lane == 0) {
            const uint32_t a_step = (p.ta ? C_::MNMAJ_STEP : C_::KMAJ_STEP) >> 4;
            const uint32_t b_step = (p.tb ? C_::KMAJ_STEP : C_::MNMAJ_STEP) >> 4;
            // products of the split operands, largest first: a1b1, a1b2, a2b1, (a1b3, a2b2, a3b1)
            constexpr int PA[6] = {0, 0, 1, 0, 1, 2};
            constexpr int PB[6] = {0, 1, 0, 2, 1, 0};
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < p.num_tiles * p.splits; u += gridDim.x) {
                int z, mb, nb;
                const int slice = u / p.num_tiles;
                decode_tile(p, u - slice * p.num_tiles, z, mb, nb);
                // instruction N = the columns this tile really has (multiple of 16): a narrow output
                // (conv layers: 64 / 96 filters) does not pay for 256 columns of tensor time
                const int n_eff = min(BN, (p.N - nb * BN + 15) & ~15);
                const uint32_t idesc = C_::kBf16 ? ptx::make_idesc_bf16(BM, n_eff, p.ta ? 1 : 0, p.tb ? 0 : 1)
                                                 : ptx::make_idesc_tf32(BM, n_eff, p.ta ? 1 : 0, p.tb ? 0 : 1);
                const int kb0 = slice * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
                for (int c0 = kb0; c0 < kb1; c0 += p.kb_per_chunk) {        // one accumulator per chunk of the chain
                    const int c1 = min(kb1, c0 + p.kb_per_chunk);
                    ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kb = c0; kb < c1; ++kb) {
                        ptx::mbar_wait(full_bar(stage), phase);
                        ptx::tc_fence_after();
#pragma unroll
                        for (int q = 0; q < C_::NPROD; ++q) {
                            const uint64_t adesc = operand_desc<MODE>(a_addr(stage, PA[q]), p.ta != 0);
                            const uint64_t bdesc = operand_desc<MODE>(b_addr(stage, PB[q]), p.tb == 0);
#pragma unroll
                            for (int j = 0; j < C_::KSTEPS; ++j) {
                                const uint32_t accum = ((kb - c0) | q | j) != 0;
                                if (C_::kBf16) ptx::mma_bf16(tmem_d, adesc + (uint64_t)(a_step * j), bdesc + (uint64_t)(b_step * j), idesc, accum);
                                else           ptx::mma_tf32(tmem_d, adesc + (uint64_t)(a_step * j), bdesc + (uint64_t)(b_step * j), idesc, accum);
                            }
                        }
                        ptx::mma_commit(empty_bar(stage));          // stage reusable once these MMAs retire
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    ptx::mma_commit(tfull_bar(acc));                // accumulator complete
                    if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ======================================= epilogue ========================================
        int acc = 0; uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < p.num_tiles * p.splits; u += gridDim.x) {
            int z, mb, nb;
            const int slice = u / p.num_tiles;
            decode_tile(p, u - slice * p.num_tiles, z, mb, nb);
            float *C = p.C[z];
            const int kb0 = slice * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
            // chunks of a chained accumulation: the first one stores, the others add to what THIS thread stored (same
            // rows and columns every time: program order is all the ordering the read-modify-write needs)
            for (int c0 = kb0; c0 < kb1; c0 += p.kb_per_chunk) {
                const bool first = c0 == kb0;
                ptx::mbar_wait(tfull_bar(acc), acc_phase);
                ptx::tc_fence_after();
                const int m0 = mb * BM + warp * 32;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * BN;
                const int nchunks = min(BN / 32, (p.N - nb * BN + 31) / 32);      // the columns the MMAs wrote
                // tcgen05.ld hands lane l the 32 columns of ROW l; storing that straight out would touch 32
                // different rows per instruction.  The 32 x 32 block is turned through a padded shared-memory
                // tile so that every store (and mask / accumulate load) instruction covers four whole 128-B rows.
                const int sub_n = 4 * (lane & 7), sub_r = lane >> 3;
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
                    const int n = nb * BN + c * 32 + sub_n;
                    if (n < p.N) {                // N % 8 == 0 (eligibility), so a float4 is all-in or all-out
                        // one straight-line variant per epilogue kind (a single epilogue warp per scheduler has
                        // nobody to hide its latency behind: instruction count is what the K = 64 GEMMs pay for)
                        if (p.splits > 1) { // raw partial sums; splitk_reduce adds the slices and applies the epilogue
                            float *part = p.part + ((size_t)slice * p.nz + z) * p.M * p.N;
                            if (first) store_rows<EK_RAW>(p, stg, part, p.N, m0, n, sub_r, sub_n);
                            else store_rows<EK_ACC>(p, stg, part, p.N, m0, n, sub_r, sub_n);
                        }
                        else if (p.epi.mode == EPI_BIAS_ACT) store_rows<EK_BIAS_ACT>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);   // (never chunked)
                        else if (p.epi.mode == EPI_MASK) store_rows<EK_MASK>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);
                        else if (p.epi.accumulate || !first) store_rows<EK_ACC>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);
                        else store_rows<EK_RAW>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);
                    }
                    __syncwarp();
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));   // 4 arrivals free the accumulator
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, NACC * BN);
}

// ================================== CTA-pair variant (cta_group::2) ================================
// The bf16 modes on a 2-CTA cluster (the two SMs of a TPC): one 256 x 256 output tile per pair, tcgen05.mma
// cta_group::2 with M = 256 issued by the leader CTA.  CTA r loads ITS 128 rows of A and ITS 128 columns of B; the
// MMA reads both halves of B across the pair, so a CTA pulls 16 KB (A) + 16 KB (half of B) per 64 k of a 128 x 256
// output instead of 16 + 32 KB: with one product per operand pair (MODE_BF16X1) the single-CTA kernel needs 94 B
// per clock and SM from L2, more than the ~61 B the L2 -> SM path delivers (tools/ubench/l2_peak.cu), and ran at
// 0.64 of the tensor peak; the pair needs 62.5.
// Barriers: full[s] lives in the leader (its producer announces the bytes of BOTH CTAs, both producers' TMA boxes
// complete on it); empty[s] and tfull[a] exist in both CTAs and are signalled by multicast commits; tempty[a] of the
// leader collects the four epilogue warps of both CTAs.
template <int MODE>
struct Cfg2 {
    using C1 = Cfg<MODE>;
    static_assert(C1::kBf16, "CTA-pair kernel: bf16 modes");
    static constexpr int NP = C1::NP;
    static constexpr int BNH = BN / 2;                              // B columns held by one CTA
    static constexpr int BK = C1::BK;
    static constexpr int A_PIECE = BM * BK * 2, B_PIECE = BNH * BK * 2;
    static constexpr int STAGE_BYTES = NP * (A_PIECE + B_PIECE);    // per CTA
    static constexpr int NSTAGE = 6;                                // 192 KB
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 + 256 + STG_BYTES;
};

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
                    const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
                    const Params p)
{
    using C_ = Cfg<MODE>;
    using C2 = Cfg2<MODE>;
    constexpr int NSTAGE = C2::NSTAGE, STAGE_BYTES = C2::STAGE_BYTES, NP = C2::NP;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + NSTAGE * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + NACC + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 2 * NACC);
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
    float *stg = reinterpret_cast<float *>(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (threadIdx.x >> 5 & 3) * 32 * STG_LD;
    auto a_addr = [&](int stage, int piece) { return smem_base + stage * STAGE_BYTES + piece * C2::A_PIECE; };
    auto b_addr = [&](int stage, int piece) { return smem_base + stage * STAGE_BYTES + NP * C2::A_PIECE + piece * C2::B_PIECE; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 8); }
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) {
        ptx::tma_prefetch_desc(&mapA0); ptx::tma_prefetch_desc(&mapB0);
        if (p.nz > 1) { ptx::tma_prefetch_desc(&mapA1); ptx::tma_prefetch_desc(&mapB1); }
    }
    if (warp == 5) ptx::tmem_alloc_pair(tmem_slot, NACC * BN);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();                 // the peer's barriers exist before anything is signalled across the pair
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        // ============ TMA producer (both CTAs): my rows of A, my columns of B, bytes on the leader's barrier ============
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = pair; u < p.num_tiles; u += npairs) {
                int z, mb, nb;
                decode_tile(p, u, z, mb, nb);
                const CUtensorMap *ma = z ? &mapA1 : &mapA0;
                const CUtensorMap *mbp = z ? &mapB1 : &mapB0;
                const int m0 = mb * 2 * BM + (int)rank * BM, n0 = nb * BN + (int)rank * C2::BNH;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    if (leader) ptx::mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
#pragma unroll
                    for (int pc = 0; pc < NP; ++pc) {
                        if (!p.ta) {
                            ptx::tma_load_3d_pair(a_addr(stage, pc), ma, kb * C_::BK, m0, pc, full_bar(stage));
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / C_::MN_BOX; ++j)
                                ptx::tma_load_3d_pair(a_addr(stage, pc) + j * C_::MN_BOX_BYTES, ma, m0 + j * C_::MN_BOX,
                                                      kb * C_::BK, pc, full_bar(stage));
                        }
                        if (p.tb) {
                            ptx::tma_load_3d_pair(b_addr(stage, pc), mbp, kb * C_::BK, n0, pc, full_bar(stage));
                        } else {
#pragma unroll
                            for (int j = 0; j < C2::BNH / C_::MN_BOX; ++j)
                                ptx::tma_load_3d_pair(b_addr(stage, pc) + j * C_::MN_BOX_BYTES, mbp, n0 + j * C_::MN_BOX,
                                                      kb * C_::BK, pc, full_bar(stage));
                        }
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ================================ MMA issuer (leader CTA only) ================================
        if (lane == 0 && leader) {
            const uint32_t a_step = (p.ta ? C_::MNMAJ_STEP : C_::KMAJ_STEP) >> 4;
            const uint32_t b_step = (p.tb ? C_::KMAJ_STEP : C_::MNMAJ_STEP) >> 4;
            constexpr int PA[6] = {0, 0, 1, 0, 1, 2};
            constexpr int PB[6] = {0, 1, 0, 2, 1, 0};
            const uint32_t idesc = ptx::make_idesc_bf16(2 * BM, BN, p.ta ? 1 : 0, p.tb ? 0 : 1);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = pair; u < p.num_tiles; u += npairs) {
                for (int c0 = 0; c0 < p.kblocks; c0 += p.kb_per_chunk) {   // one accumulator per chunk of the chain
                    const int c1 = min(p.kblocks, c0 + p.kb_per_chunk);
                    ptx::mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1);     // epilogue warps of both CTAs
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kb = c0; kb < c1; ++kb) {
                        ptx::mbar_wait(full_bar(stage), phase);
                        ptx::tc_fence_after();
#pragma unroll
                        for (int q = 0; q < C_::NPROD; ++q) {
                            const uint64_t adesc = operand_desc<MODE>(a_addr(stage, PA[q]), p.ta != 0);
                            const uint64_t bdesc = operand_desc<MODE>(b_addr(stage, PB[q]), p.tb == 0);
#pragma unroll
                            for (int j = 0; j < C_::KSTEPS; ++j)
                                ptx::mma_bf16_pair(tmem_d, adesc + (uint64_t)(a_step * j), bdesc + (uint64_t)(b_step * j), idesc,
                                                   ((kb - c0) | q | j) != 0);
                        }
                        ptx::mma_commit_pair(empty_bar(stage));         // frees the stage in both CTAs
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    ptx::mma_commit_pair(tfull_bar(acc));
                    if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ================== epilogue (both CTAs): my 128 rows of the pair's accumulator ==================
        int acc = 0; uint32_t acc_phase = 0;
        const uint32_t tempty_leader = ptx::mapa(tempty_bar(0), 0);
        for (int u = pair; u < p.num_tiles; u += npairs) {
            int z, mb, nb;
            decode_tile(p, u, z, mb, nb);
            float *C = p.C[z];
            for (int c0 = 0; c0 < p.kblocks; c0 += p.kb_per_chunk) {       // chunks of a chained accumulation, as in gemm_tc_kernel
                const bool first = c0 == 0;
                ptx::mbar_wait(tfull_bar(acc), acc_phase);
                ptx::tc_fence_after();
                const int m0 = mb * 2 * BM + (int)rank * BM + warp * 32;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * BN;
                const int sub_n = 4 * (lane & 7), sub_r = lane >> 3;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld32(taddr + c * 32, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4 *>(stg + lane * STG_LD + 4 * q) =
                            make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                        __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                    __syncwarp();
                    const int n = nb * BN + c * 32 + sub_n;
                    if (p.epi.mode == EPI_BIAS_ACT) store_rows<EK_BIAS_ACT>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);
                    else if (p.epi.mode == EPI_MASK) store_rows<EK_MASK>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);
                    else if (p.epi.accumulate || !first) store_rows<EK_ACC>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);
                    else store_rows<EK_RAW>(p, stg, C, p.ldc, m0, n, sub_r, sub_n);
                    __syncwarp();
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_remote(tempty_leader + 8u * acc);      // (the leader's own window for rank 0)
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();                 // nobody leaves while the pair's MMAs / remote arrivals can still touch it
    if (warp == 5) ptx::tmem_dealloc_pair(tmem_base, NACC * BN);
}

// ---- operand split pre-pass: fp32 [rows][ld] -> NP bf16 matrices [NP][rows][ldo] ---------------------
// piece 1 = bf16(a), piece 2 = bf16(a - piece 1), piece 3 = bf16(a - piece 1 - piece 2); the
// subtractions are exact in fp32.  HBM-bound: 4 B in, 2*NP B out per element.
template <int NP>
__global__ void split_bf16_kernel(const float *__restrict__ x, int rows, int cols, int ld,
                                  __nv_bfloat16 *__restrict__ out, int ldo)
{
    const int c4 = cols / 4;
    const size_t total = (size_t)rows * c4;
    const size_t piece = (size_t)rows * ldo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / c4), c = (int)(i % c4) * 4;
        const float4 v = *reinterpret_cast<const float4 *>(x + (size_t)r * ld + c);
        float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int pc = 0; pc < NP; ++pc) {
            __nv_bfloat16 h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { h[j] = __float2bfloat16_rn(f[j]); f[j] -= __bfloat162float(h[j]); }
            uint2 pk;
            pk.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
            pk.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
            *reinterpret_cast<uint2 *>(out + pc * piece + (size_t)r * ldo + c) = pk;
        }
    }
}

// ---- host side: tensor maps ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// 3-D tensor [piece][outer][inner] (row pitch ld elements, piece pitch `piece_elems`), box =
// [1][box_outer][box_inner].  The piece dimension keeps boxes at the K / M / N tails from running
// into the next piece: out-of-range rows are zero-filled per dimension.
static int encode_map(CUtensorMap *map, const void *base, bool bf16, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint64_t npiece, uint64_t piece_elems, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(CTCASR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const uint64_t esz = bf16 ? 2 : 4;
    cuuint64_t dims[3] = {inner, outer, npiece};
    cuuint64_t strides[2] = {ld * esz, (npiece > 1 ? piece_elems : outer * ld) * esz};
    cuuint32_t box[3] = {box_inner, box_outer, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    // fp32 operands use element type TFLOAT32: the TMA unit converts fp32 -> tf32 (round to nearest)
    // while copying, so the tensor cores do not truncate the low 13 mantissa bits themselves
    // (truncation measured as a systematic -7e-4 relative bias per GEMM).  CTCASR_TMA_F32=1 disables.
    static const bool plain_f32 = getenv("CTCASR_TMA_F32") != nullptr;
    const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                        : (plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32);
    CUresult r = fn(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CTCASR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                                      (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
    return CTCASR_OK;
}

// ---- context: the caller's scratch arena for the split operands (the library allocates nothing) and the split cache
// (see split_scope_begin in gemm.cuh).  One per ctcasr_handle_t; a thread works on the context it bound with
// ctcasr_use() (the process-default one otherwise), so two host threads driving two streams do not share an arena.
struct SplitEntry {
    const float *base; int rows, cols, ld, np;      // the fp32 matrix that was split
    __nv_bfloat16 *out; int ldo; size_t piece;      // its pieces in the arena
};
struct Context {
    char *scratch = nullptr;
    size_t scratch_bytes = 0, scratch_needed = 0;
    SplitEntry split[16];
    int nsplit = 0;
    bool scope = false;
    size_t cursor = 0;                              // arena bytes in use (reset per GEMM outside a scope)
};
static Context g_default_ctx;
static thread_local Context *g_ctx = nullptr;
static inline Context &ctx() { return g_ctx ? *g_ctx : g_default_ctx; }
#define g_scratch (ctx().scratch)
#define g_scratch_bytes (ctx().scratch_bytes)
#define g_scratch_needed (ctx().scratch_needed)
#define g_split (ctx().split)
#define g_nsplit (ctx().nsplit)
#define g_scope (ctx().scope)
#define g_cursor (ctx().cursor)

// bf16 pieces of the fp32 matrix x [rows][cols] (pitch ld): from the cache when x lies inside a matrix that
// was split in this scope, otherwise split now.  Returns the pointer to piece 0, its pitch and the piece stride.
static const SplitEntry *find_split(const float *x, int rows, int cols, int ld, int np, size_t *r0, size_t *c0)
{
    if (!g_scope) return nullptr;
    for (int i = 0; i < g_nsplit; ++i) {
        const SplitEntry &e = g_split[i];
        if (e.np != np || e.ld != ld || x < e.base) continue;
        const size_t off = (size_t)(x - e.base);
        *r0 = off / (size_t)ld; *c0 = off % (size_t)ld;
        if (*r0 + rows > (size_t)e.rows || *c0 + cols > (size_t)e.cols || (*c0 & 7)) continue;
        return &e;
    }
    return nullptr;
}
// arena bytes a split of this operand takes (0 when it is served from the cache)
static size_t split_bytes(const float *x, int rows, int cols, int ld, int np)
{
    size_t r0, c0;
    if (find_split(x, rows, cols, ld, np, &r0, &c0)) return 0;
    return align_up((size_t)rows * ((cols + 7) / 8 * 8) * np * 2, 1024);
}

template <int NP>
static int acquire_split(const float *x, int rows, int cols, int ld, cudaStream_t stream,
                         const __nv_bfloat16 **out, int *ldo, size_t *piece)
{
    size_t r0, c0;
    if (const SplitEntry *e = find_split(x, rows, cols, ld, NP, &r0, &c0)) {
        *out = e->out + r0 * e->ldo + c0; *ldo = e->ldo; *piece = e->piece;
        return CTCASR_OK;
    }
    const int lo = (cols + 7) / 8 * 8;
    const size_t pc = (size_t)rows * lo;
    const size_t bytes = align_up(pc * NP * 2, 1024);
    if (g_cursor + bytes > g_scratch_bytes) {       // launch() checks the whole GEMM first: not reached from there
        g_scratch_needed = g_cursor + bytes;
        return fail(CTCASR_ERR_WORKSPACE, "gemm_tc: split-operand scratch %zu B < %zu B needed (ctcasr_set_scratch)",
                    g_scratch_bytes, g_scratch_needed);
    }
    __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(g_scratch + g_cursor);
    g_cursor += bytes;
    const size_t n4 = (size_t)rows * (cols / 4);
    const int grid = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    {
        ProfScope prof_split(PROF_SPLIT, stream);
        split_bf16_kernel<NP><<<grid, 256, 0, stream>>>(x, rows, cols, ld, dst, lo);
        CTCASR_LAUNCH_CHECK();
    }
    if (g_scope && g_nsplit < 16) g_split[g_nsplit++] = SplitEntry{x, rows, cols, ld, NP, dst, lo, pc};
    *out = dst; *ldo = lo; *piece = pc;
    return CTCASR_OK;
}

// k-blocks per chunk of a chained accumulation (Params::kb_per_chunk) for a unit of `kblocks` k-blocks of `bk` elements:
// chains up to 12288 elements stay whole (every forward product of the path: K <= 4096), longer ones (the weight gradients:
// K = frames of the batch; input gradients through 8H = 16384 gate columns) are cut into equal chunks of <= 8192: error
// <= 3e-5 whatever K, +0.3 % on a lone K = 32,000 weight-gradient GEMM and ~1 % of the cfg2 step (profiles/r2_chain_ab.json;
// chunks of 4096: 1.6e-5, +2.7 % on that GEMM).  Not for the
// bias + activation and mask epilogues of an unsplit GEMM, which are applied once to the complete sum (split-K slices
// store raw partial sums whatever the epilogue).
template <int MODE>
static int chunk_kblocks(const Epilogue &epi, int splits, int kblocks, int bk)
{
    // one product per operand pair (bf16, tf32): the operand rounding (2e-3 / 3e-4) is far above the accumulator's error
    if (MODE != MODE_BF16X3 && MODE != MODE_BF16X6) return kblocks;
    if (splits == 1 && epi.mode != EPI_STORE) return kblocks;
    return chain_chunk_kblocks(kblocks, bk);
}

template <int MODE>
static int launch_pair(const CUtensorMap *maps, Params p, int num_sms, cudaStream_t stream)
{
    if (MODE != MODE_BF16X1 && MODE != MODE_BF16X3) return fail(CTCASR_ERR_INVALID, "gemm_tc: CTA-pair kernel: bf16 modes only");
    constexpr int MP = (MODE == MODE_BF16X1 || MODE == MODE_BF16X3) ? MODE : MODE_BF16X1;     // (never instantiated for the others)
    using C2 = Cfg2<MP>;
    static bool attr_set = false;
    if (!attr_set) {
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_pair_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::SMEM_BYTES));
        attr_set = true;
    }
    p.tiles_m = ceil_div(p.M, 2 * BM);                          // tiles of the pair: 256 rows
    p.num_tiles = p.tiles_m * p.tiles_n * p.nz;
    p.splits = 1; p.kb_per_split = p.kblocks; p.part = nullptr;
    p.kb_per_chunk = chunk_kblocks<MODE>(p.epi, 1, p.kblocks, C2::BK);
    const int pairs = p.num_tiles < num_sms / 2 ? p.num_tiles : num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = C2::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ProfScope prof_gemm(PROF_GEMM_TC, stream);
    CTCASR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<MP>, maps[0], maps[1], maps[2], maps[3], p));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return CTCASR_OK;
}

template <int MODE>
static int launch(const GemmArgs &g, cudaStream_t stream)
{
    using C_ = Cfg<MODE>;
    constexpr int NP = C_::NP;
    const void *Abase[2] = {g.A[0], g.nz > 1 ? g.A[1] : g.A[0]};
    const void *Bbase[2] = {g.B[0], g.nz > 1 ? g.B[1] : g.B[0]};
    // stored shapes: A is [M][K] (or [K][M] when ta), B is [K][N] (or [N][K] when tb)
    const int a_rows = g.ta ? g.K : g.M, a_cols = g.ta ? g.M : g.K;
    const int b_rows = g.tb ? g.N : g.K, b_cols = g.tb ? g.K : g.N;
    int lda = g.lda, ldb = g.ldb;
    size_t a_piece = 0, b_piece = 0;
    if (C_::kBf16) {
        if (!g_scope) g_cursor = 0;
        size_t need = g_cursor;                      // everything this GEMM will split, checked before the first launch
        for (int z = 0; z < g.nz; ++z)
            need += split_bytes(g.A[z], a_rows, a_cols, g.lda, NP) + split_bytes(g.B[z], b_rows, b_cols, g.ldb, NP);
        if (need > g_scratch_bytes) {
            g_scratch_needed = need;
            return fail(CTCASR_ERR_WORKSPACE, "gemm_tc: split-operand scratch %zu B < %zu B needed (ctcasr_set_scratch)",
                        g_scratch_bytes, need);
        }
        for (int z = 0; z < g.nz; ++z) {
            const __nv_bfloat16 *sa = nullptr, *sb = nullptr;
            int la = 0, lb = 0;
            size_t pa = 0, pb = 0;
            if (int rc = acquire_split<NP>(g.A[z], a_rows, a_cols, g.lda, stream, &sa, &la, &pa)) return rc;
            if (int rc = acquire_split<NP>(g.B[z], b_rows, b_cols, g.ldb, stream, &sb, &lb, &pb)) return rc;
            // the two problems of a batched GEMM share one tensor-map geometry
            if (z > 0 && (la != lda || lb != ldb || pa != a_piece || pb != b_piece))
                return fail(CTCASR_ERR_INVALID, "gemm_tc: batched operands with different split layouts");
            Abase[z] = sa; Bbase[z] = sb; lda = la; ldb = lb; a_piece = pa; b_piece = pb;
        }
        if (g.nz == 1) { Abase[1] = Abase[0]; Bbase[1] = Bbase[0]; }
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        CTCASR_CUDA_CHECK(cudaGetDevice(&dev));
        CTCASR_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // CTA-pair kernel (gemm_tc_pair_kernel): whole 256-column tiles, at least one 256 x 256 tile per pair, no split-K
    const int pair_tiles = ceil_div(g.M, 2 * BM) * (g.N / BN) * g.nz;
    static const bool pair_enabled = !(getenv("CTCASR_GEMM_PAIR") && atoi(getenv("CTCASR_GEMM_PAIR")) == 0);
    // (three products per operand pair are tensor-bound on one CTA already; measured: the pair gains 6 % with both operands
    // as stored by the forward pass and loses 4 % in the dgrad / wgrad orientations)
    const bool use_pair = pair_enabled && (MODE == MODE_BF16X1 || (MODE == MODE_BF16X3 && !g.ta && !g.tb)) && g.N % BN == 0 &&
                          pair_tiles >= num_sms / 2;
    CUtensorMap maps[4];
    const CUtensorMapSwizzle sw_k = (C_::kBf16 && C_::BK == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;   // 64-B / 128-B rows
    const CUtensorMapSwizzle sw_mn = C_::kBf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    for (int z = 0; z < 2; ++z) {
        int rc;
        if (!g.ta) rc = encode_map(&maps[2 * z], Abase[z], C_::kBf16, g.K, g.M, lda, NP, a_piece, C_::BK, BM, sw_k);             // A[m][k]
        else       rc = encode_map(&maps[2 * z], Abase[z], C_::kBf16, g.M, g.K, lda, NP, a_piece, C_::MN_BOX, C_::BK, sw_mn);    // A[k][m]
        if (rc != CTCASR_OK) return rc;
        if (g.tb)  rc = encode_map(&maps[2 * z + 1], Bbase[z], C_::kBf16, g.K, g.N, ldb, NP, b_piece, C_::BK, use_pair ? BN / 2 : BN, sw_k);   // B[n][k]
        else       rc = encode_map(&maps[2 * z + 1], Bbase[z], C_::kBf16, g.N, g.K, ldb, NP, b_piece, C_::MN_BOX, C_::BK, sw_mn); // B[k][n]
        if (rc != CTCASR_OK) return rc;
    }
    Params p;
    p.M = g.M; p.N = g.N; p.K = g.K; p.nz = g.nz; p.ta = g.ta; p.tb = g.tb; p.ldc = g.ldc;
    p.C[0] = g.C[0]; p.C[1] = g.nz > 1 ? g.C[1] : g.C[0];
    p.epi = g.epi;
    p.tiles_m = ceil_div(g.M, BM); p.tiles_n = ceil_div(g.N, BN); p.kblocks = ceil_div(g.K, C_::BK);
    p.num_tiles = p.tiles_m * p.tiles_n * g.nz;
    if (use_pair) return launch_pair<MODE>(maps, p, num_sms, stream);
    static bool attr_set = false;
    if (!attr_set) {
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
        attr_set = true;
    }
    // split-K: under half a wave of output tiles on a long contraction (conv weight gradients: [Kp, 64]
    // outputs contracting 10^5..10^6 patch rows) -> ~2 waves of (slice, tile) units, >= 32 k-blocks per slice
    p.splits = 1; p.kb_per_split = p.kblocks; p.part = nullptr;
    constexpr int KB32 = C_::BK / 32;                           // thresholds in 32-wide k-blocks whatever the stage depth
    if (p.num_tiles <= num_sms / 2 && p.kblocks * KB32 >= 128) {
        int S = (2 * num_sms) / p.num_tiles;
        if (S > p.kblocks * KB32 / 32) S = p.kblocks * KB32 / 32;
        if (S > 1) {
            const int per = ceil_div(p.kblocks, S);
            S = ceil_div(p.kblocks, per);
            const size_t bytes = (size_t)S * g.nz * g.M * g.N * sizeof(float);
            const size_t used = align_up(C_::kBf16 || g_scope ? g_cursor : 0, 1024);
            if (S > 1 && g_scratch && used + bytes <= g_scratch_bytes) {      // no room: unsplit (slower, same result class)
                p.splits = S; p.kb_per_split = per; p.part = reinterpret_cast<float *>(g_scratch + used);
            }
        }
    }
    p.kb_per_chunk = chunk_kblocks<MODE>(p.epi, p.splits, p.kb_per_split, C_::BK);
    const int units = p.num_tiles * p.splits;
    const int grid = units < num_sms ? units : num_sms;
    {
        ProfScope prof_gemm(PROF_GEMM_TC, stream);
        gemm_tc_kernel<MODE><<<grid, NTHREADS, C_::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
        CTCASR_LAUNCH_CHECK();
    }
    if (p.splits > 1) return splitk_reduce(g, p.splits, p.part, stream);
    return CTCASR_OK;
}

}  // namespace tc

int chain_chunk_kblocks(int kblocks, int bk)
{
    static const int chain = getenv("CTCASR_GEMM_CHAIN") ? atoi(getenv("CTCASR_GEMM_CHAIN")) : 8192;
    if (chain <= 0) return kblocks;
    const int target = chain / bk > 0 ? chain / bk : 1;
    if (kblocks <= target + target / 2) return kblocks;
    const int n = ceil_div(kblocks, target);
    return ceil_div(kblocks, n);
}

bool gemm_tc_eligible(const GemmArgs &g)
{
    if (g.M < 1 || g.N < 64 || g.K < 8) return false;
    if ((g.M % 8) || (g.N % 8) || (g.K % 8)) return false;                         // 16-B rows of the bf16 pieces
    if ((g.lda % 4) || (g.ldb % 4) || (g.ldc % 4)) return false;                   // TMA 16-B pitches, float4 stores
    if ((double)g.M * g.N * g.K < 4.0e6) return false;                             // not worth a persistent launch
    for (int z = 0; z < g.nz; ++z)
        if (((uintptr_t)g.A[z] | (uintptr_t)g.B[z] | (uintptr_t)g.C[z]) & 15) return false;
    if (g.epi.mode == EPI_MASK && ((g.epi.ldm % 4) || ((uintptr_t)g.epi.mask_y & 15))) return false;
    return true;
}

// Upper bound of the split-operand scratch a GEMM of this shape can need (3 pieces, padded rows).
// Composite entry points check it BEFORE launching anything, so a too-small arena never leaves
// half-updated buffers behind.
int gemm_scratch_check(int compute, int nz, int M, int N, int K)
{
    if (compute != CTCASR_COMPUTE_BF16X3 && compute != CTCASR_COMPUTE_BF16) return CTCASR_OK;
    auto pad = [](int v) { return (size_t)((v + 7) / 8 * 8); };
    const size_t need = (size_t)nz * (align_up(6 * pad(M) * pad(K), 1024) + align_up(6 * pad(K) * pad(N), 1024));
    if (need > tc::ctx().scratch_bytes) {
        if (need > tc::ctx().scratch_needed) tc::ctx().scratch_needed = need;
        return fail(CTCASR_ERR_WORKSPACE, "split-operand scratch %zu B < %zu B needed (ctcasr_set_scratch)", tc::ctx().scratch_bytes, need);
    }
    return CTCASR_OK;
}

int split_scope_begin(int compute, const size_t *elems, int n)
{
    if (compute != CTCASR_COMPUTE_BF16X3 && compute != CTCASR_COMPUTE_BF16) return CTCASR_OK;
    size_t need = 0;
    for (int i = 0; i < n; ++i) need += align_up(6 * elems[i], 1024);
    if (need > tc::ctx().scratch_bytes) {
        if (need > tc::ctx().scratch_needed) tc::ctx().scratch_needed = need;
        return fail(CTCASR_ERR_WORKSPACE, "split-operand scratch %zu B < %zu B needed (ctcasr_set_scratch)", tc::ctx().scratch_bytes, need);
    }
    tc::ctx().scope = true; tc::ctx().nsplit = 0; tc::ctx().cursor = 0;
    return CTCASR_OK;
}
void split_scope_end() { tc::ctx().scope = false; tc::ctx().nsplit = 0; tc::ctx().cursor = 0; }

int gemm_tc_pieces(const float *x, int rows, int cols, int ld, int np, cudaStream_t stream,
                   const __nv_bfloat16 **out, int *ldo, size_t *piece)
{
    if (!tc::ctx().scope) return fail(CTCASR_ERR_INVALID, "gemm_tc_pieces: needs an open split scope");
    if (np == 1) return tc::acquire_split<1>(x, rows, cols, ld, stream, out, ldo, piece);
    if (np == 2) return tc::acquire_split<2>(x, rows, cols, ld, stream, out, ldo, piece);
    if (np == 3) return tc::acquire_split<3>(x, rows, cols, ld, stream, out, ldo, piece);
    return fail(CTCASR_ERR_INVALID, "gemm_tc_pieces: %d pieces", np);
}

int tma_encode_bf16(void *map, const void *base, int rank, const unsigned long long *dims, const unsigned long long *strides,
                    const unsigned *box, const unsigned *estr, int swizzle_bytes)
{
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) return fail(CTCASR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i];
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), d, st, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CTCASR_ERR_CUDA, "cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)r);
    return CTCASR_OK;
}

int split_reserve(const float *key, int rows, int cols, int ld, int np, __nv_bfloat16 **pieces)
{
    using namespace tc;
    if (!g_scope || g_nsplit >= 16 || (cols & 7) || np < 1 || np > 3)
        return fail(CTCASR_ERR_INVALID, "split_reserve: needs an open split scope and 16-B piece rows");
    const size_t pc = (size_t)rows * cols;
    const size_t bytes = align_up(pc * np * 2, 1024);
    if (g_cursor + bytes > g_scratch_bytes) {
        g_scratch_needed = g_cursor + bytes;
        return fail(CTCASR_ERR_WORKSPACE, "split-operand scratch %zu B < %zu B needed (ctcasr_set_scratch)", g_scratch_bytes, g_scratch_needed);
    }
    __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(g_scratch + g_cursor);
    g_cursor += bytes;
    g_split[g_nsplit++] = SplitEntry{key, rows, cols, ld, np, dst, cols, pc};
    *pieces = dst;
    return CTCASR_OK;
}

void *scratch_alloc(size_t bytes)
{
    tc::Context &c = tc::ctx();
    bytes = align_up(bytes, 1024);
    if (!c.scope) { fail(CTCASR_ERR_INVALID, "scratch_alloc: needs an open split scope"); return nullptr; }
    if (!c.scratch || c.cursor + bytes > c.scratch_bytes) {
        c.scratch_needed = c.cursor + bytes;
        fail(CTCASR_ERR_WORKSPACE, "split-operand scratch %zu B < %zu B needed (ctcasr_set_scratch)", c.scratch_bytes, c.scratch_needed);
        return nullptr;
    }
    void *p = c.scratch + c.cursor;
    c.cursor += bytes;
    return p;
}

void *scratch_free(size_t bytes)
{
    const size_t used = tc::ctx().scope ? tc::ctx().cursor : 0;
    if (!tc::ctx().scratch || used + bytes > tc::ctx().scratch_bytes) return nullptr;
    return tc::ctx().scratch + used;
}

int gemm_tc(const GemmArgs &g, int compute, cudaStream_t stream)
{
    if (!gemm_tc_eligible(g)) return fail(CTCASR_ERR_UNSUPPORTED, "gemm_tc: shape not eligible");
    if (compute == CTCASR_COMPUTE_TF32) return tc::launch<tc::MODE_TF32>(g, stream);
    if (compute == CTCASR_COMPUTE_BF16) return tc::launch<tc::MODE_BF16X1>(g, stream);
    if (compute == CTCASR_COMPUTE_BF16X3)
        return g.precise ? tc::launch<tc::MODE_BF16X6>(g, stream) : tc::launch<tc::MODE_BF16X3>(g, stream);
    return fail(CTCASR_ERR_INVALID, "gemm_tc: compute mode %d", compute);
}

}  // namespace ctcasr

extern "C" int ctcasr_set_scratch(void *ptr, size_t bytes)
{
    if (bytes && (!ptr || ((uintptr_t)ptr & 1023))) return ctcasr::fail(CTCASR_ERR_INVALID, "set_scratch: need a 1024-B aligned device pointer");
    ctcasr::tc::ctx().scratch = reinterpret_cast<char *>(ptr);
    ctcasr::tc::ctx().scratch_bytes = bytes;
    return CTCASR_OK;
}
extern "C" size_t ctcasr_scratch_needed(void) { return ctcasr::tc::ctx().scratch_needed; }
extern "C" size_t ctcasr_scratch_bytes(void) { return ctcasr::tc::ctx().scratch_bytes; }

extern "C" int ctcasr_create(ctcasr_handle_t *out)
{
    if (!out) return ctcasr::fail(CTCASR_ERR_INVALID, "create: null pointer");
    *out = reinterpret_cast<ctcasr_handle_t>(new ctcasr::tc::Context());
    return CTCASR_OK;
}
extern "C" int ctcasr_use(ctcasr_handle_t h)
{
    ctcasr::tc::g_ctx = reinterpret_cast<ctcasr::tc::Context *>(h);      // NULL: the process-default context
    return CTCASR_OK;
}
extern "C" int ctcasr_destroy(ctcasr_handle_t h)
{
    if (!h) return CTCASR_OK;
    ctcasr::tc::Context *c = reinterpret_cast<ctcasr::tc::Context *>(h);
    if (c->scope) return ctcasr::fail(CTCASR_ERR_INVALID, "destroy: a split scope is open on this context");
    if (ctcasr::tc::g_ctx == c) ctcasr::tc::g_ctx = nullptr;
    delete c;
    return CTCASR_OK;
}
