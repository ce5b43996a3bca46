// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, of where the A operand
// lives (shared memory vs tensor memory) and of the number of independent accumulators the MMAs rotate over.
// One CTA per SM, one issuing thread, a fully unrolled issue loop (compile-time operands), commit + wait at the end.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ctc_asr_b200/csrc -o tools/ubench/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include "ptx.cuh"
using namespace ctcasr;

template <int N, int NACC, bool TMEMA>
__global__ void __launch_bounds__(128, 1) k(int nouter, long long *out)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t barA = ptx::smem_u32(&bar);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(barA, 1); ptx::mbar_fence_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&slot), 512);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        ptx::fence_proxy_async();
        const uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
        const uint64_t ad = ptx::make_smem_desc(base, 16, 1024, 2), bd = ptx::make_smem_desc(base + 32768, 16, 1024, 2);
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            for (int o = 0; o < nouter; ++o) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = i & 3;
                    const uint32_t acc = tm + (uint32_t)((i % NACC) * N);
                    if (TMEMA) ptx::mma_bf16_ts(acc, tm + 384 + 8 * j, bd + (uint64_t)(2 * j), idesc, 1);
                    else ptx::mma_bf16(acc, ad + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc, 1);
                }
            }
            ptx::mma_commit(barA);
            long long t1 = clock64();
            ptx::mbar_wait(barA, rep & 1);
            long long t2 = clock64();
            if (blockIdx.x == 0 && rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

template <int N, int NACC, bool TMEMA>
void run(long long *out)
{
    const int nouter = 32;
    cudaFuncSetAttribute(k<N, NACC, TMEMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<N, NACC, TMEMA><<<148, 128, 100 * 1024>>>(nouter, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    printf("A in %s, %d accumulator(s), M=128 N=%3d K=16: issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", TMEMA ? "TMEM" : "smem", NACC, N,
           (double)out[0] / (16 * nouter), (double)out[1] / (16 * nouter));
}

int main()
{
    long long *out; cudaMallocManaged(&out, 16);
    run<32, 1, false>(out); run<32, 2, false>(out); run<32, 4, false>(out);
    run<64, 1, false>(out); run<64, 2, false>(out); run<64, 4, false>(out);
    run<128, 1, false>(out); run<128, 2, false>(out); run<256, 1, false>(out);
    run<32, 1, true>(out); run<32, 2, true>(out); run<32, 4, true>(out);
    run<64, 1, true>(out); run<64, 2, true>(out); run<64, 4, true>(out);
    run<128, 1, true>(out); run<128, 2, true>(out); run<256, 1, true>(out);
    return 0;
}
