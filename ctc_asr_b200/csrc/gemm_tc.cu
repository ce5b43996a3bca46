// gemm_tc.cu — tcgen05 / TMA / TMEM GEMM for sm_100a (CTCASR_COMPUTE_TF32).
//
// C[M,N] = op(A) op(B) with fp32 operands read as TF32 by the tensor cores and fp32 accumulation
// in tensor memory.  Used for every GEMM-shaped piece of the path: the dense layers
// (asr/util/tf_contrib.py:52-58, asr/model.py:220-232), the hoisted RNN input projection and the
// three backward GEMMs per layer (dgrad, wgrad) — in the reference these are cuBLAS sgemm calls
// issued by TensorFlow.
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 4      TMA producer: cp.async.bulk.tensor 2-D boxes of 32 fp32 (128 B, SWIZZLE_128B) into a
//               4-stage shared-memory ring (A 128x32, B 256x32 per stage = 48 KB), mbarrier expect_tx
//   warp 5      MMA issuer: one elected lane issues 4 x tcgen05.mma.kind::tf32 (128 x 256 x 8) per
//               stage; tcgen05.commit releases the stage / publishes the accumulator
//   warps 0-3   epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / clipped ReLU / dropout /
//               activation mask / accumulate -> 128-bit global stores.  Two 256-column TMEM
//               accumulators, so the epilogue of tile i overlaps the main loop of tile i+1.
// Both operand orientations are handled in the descriptors, not by transposing data:
//   K-major  (A[m][k], B[n][k]): one box [rows x 32 k];        smem desc SBO = 1024 B
//   MN-major (A[k][m], B[k][n]): boxes [32 k x 32 m|n] 4 KB apart, TMA swizzle 128B_ATOM_32B;
//                                smem desc SWIZZLE_128B_BASE32B, LBO = 4096 B, SBO = 512 B
// Tiles are walked in groups of 16 row-tiles x all column-tiles so that the ~148 tiles in flight
// share A and B panels through L2.
#include "gemm.cuh"
#include "ptx.cuh"

#include <mutex>
#include <stdlib.h>

namespace ctcasr {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int NSTAGE = 4, NACC = 2;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GROUP_M = 16;
constexpr int NTHREADS = 192;

struct Params {
    int M, N, K, nz, ta, tb, ldc;
    float *C[2];
    Epilogue epi;
    int tiles_m, tiles_n, kblocks, num_tiles;
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

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const Params p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B: 1024-B aligned
    const uint32_t bar_base = smem_base + NSTAGE * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + NACC + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 2 * NACC);
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

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
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                int z, mb, nb;
                decode_tile(p, t, z, mb, nb);
                const CUtensorMap *ma = z ? &mapA1 : &mapA0;
                const CUtensorMap *mbp = z ? &mapB1 : &mapB0;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    ptx::mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    if (!p.ta) {
                        ptx::tma_load_2d(sa, ma, kb * BK, mb * BM, full_bar(stage));
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 32; ++j)
                            ptx::tma_load_2d(sa + j * 4096, ma, mb * BM + j * 32, kb * BK, full_bar(stage));
                    }
                    if (p.tb) {
                        ptx::tma_load_2d(sb, mbp, kb * BK, nb * BN, full_bar(stage));
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j)
                            ptx::tma_load_2d(sb + j * 4096, mbp, nb * BN + j * 32, kb * BK, full_bar(stage));
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ====================================== MMA issuer =======================================
        if (lane == 0) {
            const uint32_t idesc = ptx::make_idesc_tf32(BM, BN, p.ta ? 1 : 0, p.tb ? 0 : 1);
            // per k-step (8 tf32 = 32 B along K): K-major advances 32 B inside the swizzle row,
            // MN-major advances one 8-row group (1024 B)
            const uint32_t a_step = p.ta ? (1024u >> 4) : (32u >> 4);
            const uint32_t b_step = p.tb ? (32u >> 4) : (1024u >> 4);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint64_t adesc = p.ta ? ptx::make_desc_mnmajor(sa, 4096) : ptx::make_desc_kmajor(sa);
                    const uint64_t bdesc = p.tb ? ptx::make_desc_kmajor(sb) : ptx::make_desc_mnmajor(sb, 4096);
#pragma unroll
                    for (int j = 0; j < BK / 8; ++j)
                        ptx::mma_tf32(tmem_d, adesc + (uint64_t)(a_step * j), bdesc + (uint64_t)(b_step * j), idesc,
                                      (kb | j) != 0);
                    ptx::mma_commit(empty_bar(stage));          // stage reusable once these MMAs retire
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                ptx::mma_commit(tfull_bar(acc));                // accumulator complete
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ======================================= epilogue ========================================
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            int z, mb, nb;
            decode_tile(p, t, z, mb, nb);
            float *C = p.C[z];
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const int m = mb * BM + warp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                ptx::tmem_ld32(taddr + c * 32, r);
                ptx::tmem_ld_wait();
                const int n0 = nb * BN + c * 32;
                if (m < p.M && n0 < p.N) {
                    float *crow = C + (size_t)m * p.ldc + n0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int n = n0 + 4 * q;
                        if (n < p.N) {        // N % 4 == 0 (eligibility), so a float4 is all-in or all-out
                            float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.epi.mode == EPI_STORE && p.epi.accumulate) old = *reinterpret_cast<const float4 *>(crow + 4 * q);
                            float4 o;
                            o.x = epilogue_apply(p.epi, __uint_as_float(r[4 * q + 0]), m, n + 0, p.N, old.x);
                            o.y = epilogue_apply(p.epi, __uint_as_float(r[4 * q + 1]), m, n + 1, p.N, old.y);
                            o.z = epilogue_apply(p.epi, __uint_as_float(r[4 * q + 2]), m, n + 2, p.N, old.z);
                            o.w = epilogue_apply(p.epi, __uint_as_float(r[4 * q + 3]), m, n + 3, p.N, old.w);
                            *reinterpret_cast<float4 *>(crow + 4 * q) = o;
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));   // 4 arrivals free the accumulator
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, NACC * BN);
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

// 2-D fp32 tensor [outer][inner] with row pitch ld (elements); box = [box_outer][32] (128-B rows)
static int encode_map(CUtensorMap *map, const float *base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer,
                      bool mn_major)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(CTCASR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {32, box_outer};
    cuuint32_t estr[2] = {1, 1};
    // element type TFLOAT32: the TMA unit converts fp32 -> tf32 while copying, so the tensor cores
    // see rounded operands instead of truncating the low 13 mantissa bits themselves
    // (truncation measured as a systematic -7e-4 relative bias per GEMM).  CTCASR_TMA_F32=1 disables.
    static const bool plain_f32 = getenv("CTCASR_TMA_F32") != nullptr;
    CUresult r = fn(map, plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2,
                    const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CTCASR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                                      (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
    return CTCASR_OK;
}

}  // namespace tc

bool gemm_tc_eligible(const GemmArgs &g)
{
    if (g.M < 1 || g.N < 64 || g.K < 8) return false;
    if ((g.N % 4) || (g.lda % 4) || (g.ldb % 4) || (g.ldc % 4)) return false;      // TMA 16-B pitches, float4 stores
    if ((double)g.M * g.N * g.K < 4.0e6) return false;                             // not worth a persistent launch
    for (int z = 0; z < g.nz; ++z)
        if (((uintptr_t)g.A[z] | (uintptr_t)g.B[z] | (uintptr_t)g.C[z]) & 15) return false;
    if (g.epi.mode == EPI_MASK && ((g.epi.ldm % 4) || ((uintptr_t)g.epi.mask_y & 15))) return false;
    return true;
}

int gemm_tc(const GemmArgs &g, cudaStream_t stream)
{
    using namespace tc;
    if (!gemm_tc_eligible(g)) return fail(CTCASR_ERR_UNSUPPORTED, "gemm_tc: shape not eligible");
    CUtensorMap maps[4];
    for (int z = 0; z < 2; ++z) {
        const int zz = z < g.nz ? z : 0;
        int rc;
        if (!g.ta) rc = encode_map(&maps[2 * z], g.A[zz], g.K, g.M, g.lda, BM, false);      // A[m][k]: inner k
        else       rc = encode_map(&maps[2 * z], g.A[zz], g.M, g.K, g.lda, 32, true);       // A[k][m]: inner m
        if (rc != CTCASR_OK) return rc;
        if (g.tb)  rc = encode_map(&maps[2 * z + 1], g.B[zz], g.K, g.N, g.ldb, BN, false);  // B[n][k]: inner k
        else       rc = encode_map(&maps[2 * z + 1], g.B[zz], g.N, g.K, g.ldb, 32, true);   // B[k][n]: inner n
        if (rc != CTCASR_OK) return rc;
    }
    Params p;
    p.M = g.M; p.N = g.N; p.K = g.K; p.nz = g.nz; p.ta = g.ta; p.tb = g.tb; p.ldc = g.ldc;
    p.C[0] = g.C[0]; p.C[1] = g.nz > 1 ? g.C[1] : g.C[0];
    p.epi = g.epi;
    p.tiles_m = ceil_div(g.M, BM); p.tiles_n = ceil_div(g.N, BN); p.kblocks = ceil_div(g.K, BK);
    p.num_tiles = p.tiles_m * p.tiles_n * g.nz;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        CTCASR_CUDA_CHECK(cudaGetDevice(&dev));
        CTCASR_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    gemm_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}

}  // namespace ctcasr
