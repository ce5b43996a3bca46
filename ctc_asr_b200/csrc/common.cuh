// common.cuh — shared host/device helpers for libctcasr.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/ctcasr.h"

namespace ctcasr {

// ---- error plumbing (C-ABI: int return + last-error string, no exceptions) -------------------
extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launch_count;

inline int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

#define CTCASR_CUDA_CHECK(expr)                                                                   \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            return ::ctcasr::fail(CTCASR_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                                  cudaGetErrorString(err__));                                     \
    } while (0)

#define CTCASR_REQUIRE(cond, ...)                                                                 \
    do {                                                                                          \
        if (!(cond)) return ::ctcasr::fail(CTCASR_ERR_INVALID, __VA_ARGS__);                      \
    } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches
#define CTCASR_LAUNCH_CHECK()                                                                     \
    do {                                                                                          \
        ::ctcasr::g_launch_count.fetch_add(1, std::memory_order_relaxed);                         \
        CTCASR_CUDA_CHECK(cudaGetLastError());                                                    \
    } while (0)

// ---- optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline) ---
enum ProfTag : int { PROF_LSTM_FWD = 0, PROF_LSTM_BWD = 1, PROF_GEMM_TC = 2, PROF_CTC = 3, PROF_SPLIT = 4, PROF_NTAGS = 5 };
void prof_begin(int tag, cudaStream_t s);
void prof_end(int tag, cudaStream_t s);
struct ProfScope {
    int tag; cudaStream_t s;
    ProfScope(int t, cudaStream_t st) : tag(t), s(st) { prof_begin(tag, s); }
    ~ProfScope() { prof_end(tag, s); }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- dropout keep-mask: counter hash shared bit-for-bit with oracle/oracle_impl.h ------------
__host__ __device__ inline uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__host__ __device__ inline bool drop_keep(uint32_t seed, uint64_t idx, float rate)
{
    uint32_t h = hash32((uint32_t)idx ^ hash32(seed ^ (uint32_t)(idx >> 32) * 0x9e3779b9U));
    return ((float)(h >> 8) * (1.0f / 16777216.0f)) >= rate;
}

// ---- epilogue shared by the SIMT and tcgen05 GEMMs --------------------------------------------
enum EpiMode : int {
    EPI_STORE = 0,      // C = acc (+ C if accumulate)
    EPI_BIAS_ACT = 1,   // C = dropout(act(acc + bias[n]))
    EPI_MASK = 2        // C = acc * act'(mask_y[m,n])      (dgrad fused with the previous layer's act)
};

struct Epilogue {
    int mode = EPI_STORE;
    int accumulate = 0;
    const float *bias = nullptr;   // [N]
    int act = 0;                   // 1: min(relu, cutoff)
    float cutoff = 20.f;
    float drop_rate = 0.f;
    uint32_t seed = 0;
    const float *mask_y = nullptr; // [M, ldm]
    int ldm = 0;
};

// MODE is a compile-time copy of e.mode so that a caller that has already branched on it (the tcgen05
// epilogue) gets straight-line code; epilogue_apply() is the run-time dispatch used by the SIMT kernels.
template <int MODE>
__device__ __forceinline__ float epilogue_apply_m(const Epilogue &e, float acc, int m, int n, int N, float c_old)
{
    if (MODE == EPI_BIAS_ACT) {
        float v = acc + (e.bias ? e.bias[n] : 0.f);
        if (e.act == 1) v = fminf(fmaxf(v, 0.f), e.cutoff);
        if (e.drop_rate > 0.f) {
            const float inv_keep = 1.f / (1.f - e.drop_rate);
            v = drop_keep(e.seed, (uint64_t)m * N + n, e.drop_rate) ? v * inv_keep : 0.f;
        }
        return v;
    }
    if (MODE == EPI_MASK) {
        const float inv_keep = e.drop_rate > 0.f ? 1.f / (1.f - e.drop_rate) : 1.f;
        const float y = e.mask_y[(size_t)m * e.ldm + n];
        bool pass = e.drop_rate > 0.f ? drop_keep(e.seed, (uint64_t)m * N + n, e.drop_rate) : true;
        if (e.act == 1) pass = pass && (y > 0.f) && (y < e.cutoff * inv_keep);
        return pass ? acc * inv_keep : 0.f;
    }
    return e.accumulate ? acc + c_old : acc;
}

__device__ __forceinline__ float epilogue_apply(const Epilogue &e, float acc, int m, int n, int N,
                                                float c_old)
{
    if (e.mode == EPI_BIAS_ACT) return epilogue_apply_m<EPI_BIAS_ACT>(e, acc, m, n, N, c_old);
    if (e.mode == EPI_MASK) return epilogue_apply_m<EPI_MASK>(e, acc, m, n, N, c_old);
    return epilogue_apply_m<EPI_STORE>(e, acc, m, n, N, c_old);
}

}  // namespace ctcasr
