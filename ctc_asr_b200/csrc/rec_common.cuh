// rec_common.cuh — pieces shared by the persistent tcgen05 recurrence kernels (lstm_tc.cu, rec_tc.cu):
// the cross-CTA step barrier (per-direction release/acquire counters in global memory), the bf16 operand
// split, and the host-side tensor-map encoder for [pieces][rows][K] bf16 operands.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <mutex>

namespace ctcasr {
namespace rec {

constexpr int BK = 64;                  // bf16 k-elements per tile row (128 B, SWIZZLE_128B)

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }
// tanh from one exponential: (1 - e) / (1 + e), e = exp(-2|x|); absolute error ~1e-7 (2 ulp of __expf on e <= 1), against
// ~25 instructions of the library tanhf on the serial tail of every time step
__device__ __forceinline__ float tanhf_(float v)
{
    const float e = __expf(-2.f * fabsf(v));
    return copysignf(__fdividef(1.f - e, 1.f + e), v);
}
// release fence of the step barrier: fence.acq_rel (MEMBAR.ALL.GPU) — __threadfence() is the sequentially consistent one
__device__ __forceinline__ void fence_release_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Bounded spin on a step counter: a protocol bug ends in a trapped kernel, never in a hung GPU.
__device__ __forceinline__ void wait_counter(const unsigned int *ctr, unsigned int target, unsigned int *err)
{
    unsigned int v, spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) break;
        if (++spins > (1u << 22)) {
            *err = 1;
            printf("ctcasr recurrence: step barrier timed out (block %d, target %u, have %u)\n", blockIdx.x, target, v);
            __trap();
        }
    }
}
__device__ __forceinline__ void signal_counter(unsigned int *ctr)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}

__device__ __forceinline__ void split2(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo)
{
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}
// bf16 [pieces][rows][inner], box [box_pieces][box_rows][64], SWIZZLE_128B
inline int make_map(CUtensorMap *m, const void *base, uint64_t inner, uint64_t rows, uint32_t box_rows, uint32_t box_pieces = 2,
                    uint64_t pieces = 2)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CTCASR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {inner, rows, pieces};
    cuuint64_t strides[2] = {inner * 2, rows * inner * 2};
    cuuint32_t box[3] = {BK, box_rows, box_pieces};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CTCASR_ERR_CUDA, "recurrence: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CTCASR_OK;
}

}  // namespace rec
}  // namespace ctcasr
