// api.cu — C-ABI entry points that are thin compositions of the kernels: dense layer forward /
// backward (asr/util/tf_contrib.py:52-58, asr/model.py:220-226,232), the generic GEMM, and the
// library-level bookkeeping (version, last error, launch counter).
#include "gemm.cuh"

#include <stdlib.h>

#include <utility>
#include <vector>

namespace ctcasr {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

// ---- profiling: event pairs recorded around the hot kernels when enabled ---------------------------
namespace {
struct ProfState {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pairs[PROF_NTAGS];
    cudaEvent_t open_start[PROF_NTAGS] = {};
    size_t used = 0;
    cudaEvent_t get()
    {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
} g_prof;
}  // namespace
void prof_begin(int tag, cudaStream_t s)
{
    if (!g_prof.on) return;
    cudaEvent_t e = g_prof.get();
    cudaEventRecord(e, s);
    g_prof.open_start[tag] = e;
}
void prof_end(int tag, cudaStream_t s)
{
    if (!g_prof.on || !g_prof.open_start[tag]) return;
    cudaEvent_t e = g_prof.get();
    cudaEventRecord(e, s);
    g_prof.pairs[tag].push_back({g_prof.open_start[tag], e});
    g_prof.open_start[tag] = nullptr;
}

int gemm(const GemmArgs &g, int compute, cudaStream_t stream)
{
    if (compute != CTCASR_COMPUTE_FP32 && gemm_tc_eligible(g)) return gemm_tc(g, compute, stream);
    return gemm_simt(g, stream);
}
}  // namespace ctcasr

using namespace ctcasr;

extern "C" int ctcasr_abi_version(void) { return CTCASR_ABI_VERSION; }
extern "C" const char *ctcasr_last_error(void) { return g_last_error; }
extern "C" uint64_t ctcasr_launch_count(void) { return g_launch_count.load(); }

namespace ctcasr { void lstm_tc_set_trace(unsigned long long *buf); void rec_tc_set_trace(unsigned long long *buf); }
// debugging aid (not in ctcasr.h): device buffer [grid][64][8] of globaltimer stamps written by the LSTM forward kernel
extern "C" int ctcasr_debug_lstm_trace(void *buf) { ctcasr::lstm_tc_set_trace(reinterpret_cast<unsigned long long *>(buf)); ctcasr::rec_tc_set_trace(reinterpret_cast<unsigned long long *>(buf)); return 0; }

extern "C" int ctcasr_profile_enable(int on)
{
    g_prof.on = on != 0;
    if (on) { g_prof.used = 0; for (auto &v : g_prof.pairs) v.clear(); }
    return CTCASR_OK;
}
// Synchronises the device and returns, per kernel class, the summed event time (ms) and the count.
extern "C" int ctcasr_profile_collect(double *ms, int *count, int ntags)
{
    CTCASR_REQUIRE(ms && count && ntags >= 1, "profile_collect: bad args");
    CTCASR_CUDA_CHECK(cudaDeviceSynchronize());
    for (int t = 0; t < ntags; ++t) {
        ms[t] = 0.0; count[t] = 0;
        if (t >= PROF_NTAGS) continue;
        for (auto &pr : g_prof.pairs[t]) {
            float v = 0.f;
            if (cudaEventElapsedTime(&v, pr.first, pr.second) == cudaSuccess) { ms[t] += v; count[t]++; }
        }
    }
    return CTCASR_OK;
}

extern "C" int ctcasr_gemm(const float *A, const float *B, float *C, int M, int N, int K,
                           int ta, int tb, int lda, int ldb, int ldc, int accumulate, int compute, void *stream)
{
    CTCASR_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0, "gemm: bad args");
    GemmArgs g;
    g.A[0] = A; g.B[0] = B; g.C[0] = C; g.M = M; g.N = N; g.K = K; g.ta = ta; g.tb = tb;
    g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.epi.accumulate = accumulate;
    return gemm(g, compute, (cudaStream_t)stream);
}

// A linear layer narrower than a tcgen05 tile (the 29-class logits layer, asr/model.py:229-232) over many rows: its
// weight / output-gradient operands are widened to 64 columns in the scratch arena and the products run on the tensor
// cores in the fp32-level bf16x3 arithmetic (in both bf16 compute modes), instead of three SIMT GEMMs (0.9 ms per step).
static bool narrow_on_tc(int compute, int M, int K, int N, int act, float drop_rate)
{
    static const bool enabled = !(getenv("CTCASR_NARROW_TC") && atoi(getenv("CTCASR_NARROW_TC")) == 0);
    return enabled && (compute == CTCASR_COMPUTE_BF16X3 || compute == CTCASR_COMPUTE_BF16) && N < 64 && M >= 2048 && M % 8 == 0 &&
           K >= 64 && K % 8 == 0 && act == 0 && drop_rate == 0.f;
}

extern "C" int ctcasr_dense_fwd(const float *x, const float *w, const float *bias, float *y,
                                int M, int K, int N, int act, float cutoff, float drop_rate, uint32_t seed,
                                int compute, void *stream)
{
    CTCASR_REQUIRE(x && w && y && M >= 0 && K >= 1 && N >= 1, "dense_fwd: bad args");
    CTCASR_REQUIRE(drop_rate >= 0.f && drop_rate < 1.f, "dense_fwd: drop_rate %f", drop_rate);
    GemmArgs g;
    g.A[0] = x; g.B[0] = w; g.C[0] = y; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = N; g.ldc = N;
    g.epi.mode = EPI_BIAS_ACT; g.epi.bias = bias; g.epi.act = act; g.epi.cutoff = cutoff;
    g.epi.drop_rate = drop_rate; g.epi.seed = seed;
    g.precise = act != 0;       // pre-activations near the ReLU / clip kinks decide the backward mask
    if (narrow_on_tc(compute, M, K, N, act, drop_rate)) {
        cudaStream_t st = (cudaStream_t)stream;
        const size_t elems[3] = {(size_t)M * K, (size_t)K * 64, (size_t)(M + K) * 64};
        SplitScope scope;
        if (int rc = split_scope_begin(CTCASR_COMPUTE_BF16X3, elems, 3)) return rc;
        scope.open = true;
        float *wp = reinterpret_cast<float *>(scratch_alloc((size_t)K * 64 * 4)), *yp = reinterpret_cast<float *>(scratch_alloc((size_t)M * 64 * 4));
        if (!wp || !yp) return CTCASR_ERR_WORKSPACE;
        if (int rc = pad_cols64(w, K, N, wp, st)) return rc;
        GemmArgs p;
        p.A[0] = x; p.B[0] = wp; p.C[0] = yp; p.M = M; p.N = 64; p.K = K; p.lda = K; p.ldb = 64; p.ldc = 64;
        if (gemm_tc_eligible(p)) {
            if (int rc = gemm_tc(p, CTCASR_COMPUTE_BF16X3, st)) return rc;
            return compact_cols64(yp, M, N, bias, y, st);
        }
    }
    return gemm(g, compute, (cudaStream_t)stream);
}

extern "C" int ctcasr_dense_bwd(const float *x, const float *w, const float *y, float *dy,
                                float *dx, float *dw, float *db, int M, int K, int N,
                                int act, float cutoff, float drop_rate, uint32_t seed,
                                int compute, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(x && w && dy && dw && db && M >= 0 && K >= 1 && N >= 1, "dense_bwd: bad args");
    CTCASR_REQUIRE(act == 0 || y, "dense_bwd: activation mask needs the forward output y");
    if (int rcs = gemm_scratch_check(compute, 1, K, N, M)) return rcs;
    if (int rcs = gemm_scratch_check(compute, 1, M, K, N)) return rcs;
    if (narrow_on_tc(compute, M, K, N, act, drop_rate)) {      // see ctcasr_dense_fwd
        const size_t elems[4] = {(size_t)M * K, (size_t)M * 64, (size_t)K * 64, (size_t)(M + 2 * K) * 64};
        SplitScope scope;
        if (int rcs = split_scope_begin(CTCASR_COMPUTE_BF16X3, elems, 4)) return rcs;
        scope.open = true;
        float *zp = reinterpret_cast<float *>(scratch_alloc((size_t)M * 64 * 4)), *wp = reinterpret_cast<float *>(scratch_alloc((size_t)K * 64 * 4));
        float *dwp = reinterpret_cast<float *>(scratch_alloc((size_t)K * 64 * 4));
        if (!zp || !wp || !dwp) return CTCASR_ERR_WORKSPACE;
        GemmArgs pw, px;
        pw.A[0] = x; pw.B[0] = zp; pw.C[0] = dwp; pw.ta = 1; pw.M = K; pw.N = 64; pw.K = M; pw.lda = K; pw.ldb = 64; pw.ldc = 64;
        px.A[0] = zp; px.B[0] = wp; px.C[0] = dx; px.tb = 1; px.M = M; px.N = K; px.K = 64; px.lda = 64; px.ldb = 64; px.ldc = K;
        if (gemm_tc_eligible(pw) && (!dx || gemm_tc_eligible(px))) {
            int rc = colsum(dy, M, N, N, db, stream);
            if (rc != CTCASR_OK) return rc;
            if ((rc = pad_cols64(dy, M, N, zp, stream)) != CTCASR_OK) return rc;
            if ((rc = gemm_tc(pw, CTCASR_COMPUTE_BF16X3, stream)) != CTCASR_OK) return rc;
            if ((rc = compact_cols64(dwp, K, N, nullptr, dw, stream)) != CTCASR_OK) return rc;
            if (dx) {
                if ((rc = pad_cols64(w, K, N, wp, stream)) != CTCASR_OK) return rc;
                if ((rc = gemm_tc(px, CTCASR_COMPUTE_BF16X3, stream)) != CTCASR_OK) return rc;
            }
            return CTCASR_OK;
        }
    }
    GemmArgs gw, gx;    // dW[K,N] = X^T dz;  dX[M,K] = dz W^T
    gw.A[0] = x; gw.B[0] = dy; gw.C[0] = dw; gw.ta = 1; gw.M = K; gw.N = N; gw.K = M; gw.lda = K; gw.ldb = N; gw.ldc = N;
    gx.A[0] = dy; gx.B[0] = w; gx.C[0] = dx; gx.tb = 1; gx.M = M; gx.N = K; gx.K = N; gx.lda = N; gx.ldb = N; gx.ldc = K;
    // Both products on the tcgen05 GEMM in a bf16 mode: one pass makes dz's mask, column sums and bf16 pieces (the fp32 dz
    // is then not needed); the two GEMMs share those pieces through the split scope.
    const int np = compute == CTCASR_COMPUTE_BF16 ? 1 : (compute == CTCASR_COMPUTE_BF16X3 ? 2 : 0);
    const bool fused = np && N <= 16384 && gemm_tc_eligible(gw);       // (independent of dx: so is the bias gradient's summation order)
    const bool fp32_dz = fused && dx && !gemm_tc_eligible(gx);         // the input-gradient product reads dz itself
    SplitScope scope;
    int rc;
    if (fused) {
        const size_t elems[3] = {(size_t)M * K, (size_t)M * N, (size_t)K * N};
        if ((rc = split_scope_begin(compute, elems, 3)) != CTCASR_OK) return rc;
        scope.open = true;
        rc = mask_colsum_split(dy, y, M, N, act, cutoff, drop_rate, seed, np, db, stream);
        if (rc != CTCASR_OK) return rc;
        if (fp32_dz && (rc = mask_inplace(dy, y, (size_t)M, N, act, cutoff, drop_rate, seed, stream)) != CTCASR_OK) return rc;
    } else {
        rc = mask_inplace(dy, y, (size_t)M, N, act, cutoff, drop_rate, seed, stream);   // dy -> dz
        if (rc != CTCASR_OK) return rc;
        rc = colsum(dy, M, N, N, db, stream);
        if (rc != CTCASR_OK) return rc;
    }
    rc = gemm(gw, compute, stream);
    if (rc != CTCASR_OK) return rc;
    if (dx) {
        rc = gemm(gx, compute, stream);
        if (rc != CTCASR_OK) return rc;
    }
    return CTCASR_OK;
}
