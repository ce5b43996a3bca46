// Micro-benchmark: L2 -> SM read bandwidth of the whole chip for an L2-resident working set (the bound the
// weight-streaming recurrence kernels are compared with), plus the same loop over a DRAM-sized buffer.
// Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/l2_peak tools/ubench/l2_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) rd(const uint4 *__restrict__ p, size_t n, int passes, unsigned *sink)
{
    unsigned acc = 0;
    for (int it = 0; it < passes; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));   // bypass L1
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}

static double run(const uint4 *buf, size_t bytes, int passes, unsigned *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        rd<<<148 * 8, 256>>>(buf, bytes / 16, 1, sink);          // warm the cache
        cudaEventRecord(e0);
        rd<<<148 * 8, 256>>>(buf, bytes / 16, passes, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gbs = (double)bytes * passes / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    return best;
}

int main()
{
    uint4 *buf; unsigned *sink;
    const size_t big = (size_t)4 << 30;
    cudaMalloc(&buf, big); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, big);
    printf("{\"how\": \"ld.global.cg.v4 grid-stride reads, 148x8 CTAs x 256 threads, best of 5\", \"l2_resident_gbs\": {");
    const int mb[] = {16, 32, 48, 64, 96};
    for (int i = 0; i < 5; ++i) printf("%s\"%d MiB\": %.1f", i ? ", " : "", mb[i], run(buf, (size_t)mb[i] << 20, 20, sink));
    printf("}, \"dram_4GiB_gbs\": %.1f}\n", run(buf, big, 1, sink));
    return 0;
}
