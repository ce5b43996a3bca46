// Micro-benchmark 2: the issue loop of the recurrence kernels — groups of 4 MMAs (one 64-wide k-block), optionally
// a tcgen05.commit per group, optionally a (satisfied) mbarrier wait + tcgen05 fence per group, from 1 or 2 issuing warps.
#include <cstdio>
#include "ptx.cuh"
using namespace ctcasr;

template <int N, int NISSUE, bool COMMIT, int WAIT>
__global__ void __launch_bounds__(128, 1) k(int ngroups, long long *out)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ __align__(8) unsigned long long bars[8];
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1); ptx::mbar_fence_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&slot), 512);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    if (warp < NISSUE && (threadIdx.x & 31) == 0) {
        ptx::fence_proxy_async();
        const uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
        const uint64_t bd = ptx::make_smem_desc(base + 32768 + warp * 8192, 16, 1024, 2);
        const uint32_t acc = tm + warp * N, done = ptx::smem_u32(&bars[warp]), scratch = ptx::smem_u32(&bars[2 + warp]), ready = ptx::smem_u32(&bars[4 + warp]);
        ptx::mbar_arrive(ready);                        // phase 0 of `ready` is complete: waits on parity 0 pass at once
        long long t0 = clock64();
        for (int g = 0; g < ngroups; ++g) {
            if (WAIT & 1) ptx::mbar_wait(ready, 0);
            if (WAIT & 4) {   // tight PTX wait loop
                asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(ready), "r"(0) : "memory");
            }
            if (WAIT & 2) ptx::tc_fence_after();
            const uint32_t ta = tm + 256 + (g & 3) * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) ptx::mma_bf16_ts(acc, ta + 8 * j, bd + (uint64_t)(2 * j), idesc, 1);
            if (COMMIT) ptx::mma_commit(scratch);
        }
        ptx::mma_commit(done);
        long long t1 = clock64();
        ptx::mbar_wait(done, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

template <int N, int NISSUE, bool COMMIT, int WAIT>
void run(long long *out)
{
    const int ngroups = 128;
    cudaFuncSetAttribute(k<N, NISSUE, COMMIT, WAIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<N, NISSUE, COMMIT, WAIT><<<148, 128, 100 * 1024>>>(ngroups, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    printf("N=%3d, %d issuing warp(s), commit per group %d, wait(1)/fence(2)/tight wait(4) per group %d: issue %.1f, complete %.1f cycles per MMA (per issuer)\n",
           N, NISSUE, (int)COMMIT, (int)WAIT, (double)out[0] / (4 * ngroups), (double)out[1] / (4 * ngroups));
}

int main()
{
    long long *out; cudaMallocManaged(&out, 16);
    run<32, 1, false, 0>(out); run<32, 1, false, 1>(out); run<32, 2, true, 1>(out); run<32, 1, false, 2>(out); run<32, 1, false, 3>(out); run<32, 1, false, 4>(out); run<32, 1, false, 6>(out);
    run<32, 2, true, 0>(out); run<32, 2, true, 3>(out); run<32, 2, true, 6>(out); run<32, 2, true, 4>(out);
    return 0;
}
