// gemm.cuh — GEMM argument block shared by the SIMT (gemm_simt.cu) and tcgen05 (gemm_tc.cu) paths.
#pragma once
#include "common.cuh"

#include <cuda_bf16.h>

namespace ctcasr {

struct GemmArgs {
    const float *A[2] = {nullptr, nullptr};
    const float *B[2] = {nullptr, nullptr};
    float *C[2] = {nullptr, nullptr};
    int nz = 1;                 // number of independent problems (1 or 2) sharing the shape
    int M = 0, N = 0, K = 0;
    int ta = 0, tb = 0;         // ta: A stored [K,M]; tb: B stored [N,K]
    int lda = 0, ldb = 0, ldc = 0;
    int precise = 0;            // BF16X3: use the 3-piece / 6-product split (output feeds a ReLU kink)
    Epilogue epi;
};

int gemm_simt(const GemmArgs &g, cudaStream_t stream);
// small-M (batch rows) product of one recurrence step, both directions; falls back to gemm_simt (step_gemm.cu)
int step_gemm(const GemmArgs &g, cudaStream_t stream);
// the same product with the cell math of a one-gate cell (tanh / ReLU RNN) in its epilogue
struct StepCell {
    int mode = 0;                 // 1 forward, 2 backward
    int cell = 0, use_len = 0;
    const int *seq_len = nullptr;
    int t[2] = {0, 0};            // frame of the forward / backward direction this step computes
    float *y[2] = {nullptr, nullptr};         // forward: layer output rows of those frames (column offset of the direction applied)
    const float *dy[2] = {nullptr, nullptr};  // backward: upstream gradient rows
    int ldy = 0;
};
int step_gemm_cell(const GemmArgs &g, const StepCell &sc, cudaStream_t stream);
// C = epilogue(sum over the S raw partial results part[s][z][M][N], added in slice order: deterministic)
int splitk_reduce(const GemmArgs &g, int S, float *part, cudaStream_t stream);
// returns CTCASR_ERR_UNSUPPORTED (without touching C) when the shape/alignment is not eligible
int gemm_tc(const GemmArgs &g, int compute, cudaStream_t stream);
bool gemm_tc_eligible(const GemmArgs &g);
int gemm_scratch_check(int compute, int nz, int M, int N, int K);
// Split scope: between begin and end the bf16 pieces of an operand are computed ONCE and shared by every
// GEMM that reads the same fp32 matrix or a sub-view of it (birnn_bwd reads dz in three GEMMs).  The caller
// guarantees that the fp32 operands are not written while the scope is open.  `elems[i]` = rows * padded
// columns of the distinct operands that will be split; begin() fails (before anything is launched) when
// their pieces do not fit in the scratch arena together.
int split_scope_begin(int compute, const size_t *elems, int n);
void split_scope_end();
// Inside an open scope: arena space for the np bf16 pieces [np][rows][cols] (cols a multiple of 8) of an operand the
// CALLER produces in that form, registered under `key` (the address a GEMM will name as its fp32 operand with pitch
// ld; the fp32 matrix itself need not exist).  The caller fills *pieces before the GEMM runs (stream order).
int split_reserve(const float *key, int rows, int cols, int ld, int np, __nv_bfloat16 **pieces);
// bf16 pieces [np][rows][*ldo] of the fp32 matrix x [rows][cols] (pitch ld) in the scratch arena: from the split cache of
// the open scope, or split now (and cached).  For kernels outside gemm_tc.cu that read split operands (conv_tc.cu).
int gemm_tc_pieces(const float *x, int rows, int cols, int ld, int np, cudaStream_t stream,
                   const __nv_bfloat16 **out, int *ldo, size_t *piece);
// cuTensorMapEncodeTiled for a bf16 tensor of `rank` <= 5 dimensions (dims / box / element strides innermost first,
// strides in bytes for dimensions 1 .. rank-1); swizzle_bytes 64 or 128; out-of-range elements read as zero.
int tma_encode_bf16(void *map, const void *base, int rank, const unsigned long long *dims, const unsigned long long *strides,
                    const unsigned *box, const unsigned *estr, int swizzle_bytes);
struct SplitScope {
    bool open = false;
    ~SplitScope() { if (open) split_scope_end(); }
};
// Inside an open scope: `bytes` of the arena for the caller's own temporaries (released with the scope); nullptr and
// CTCASR_ERR_WORKSPACE in ctcasr_last_error() when they do not fit.
void *scratch_alloc(size_t bytes);
// Free space of the caller's scratch arena behind the cached splits (not reserved: valid until the next
// GEMM call on the stream), or nullptr when `bytes` do not fit.
void *scratch_free(size_t bytes);
// Chained accumulation (gemm_tc.cu): k-blocks (of bk contraction elements) per chunk for a chain of `kblocks` — the whole
// chain up to 12288 elements, equal chunks of <= 8192 beyond (CTCASR_GEMM_CHAIN = elements per chunk, 0 = never chunk)
int chain_chunk_kblocks(int kblocks, int bk);
// dispatch on `compute`: TF32 -> tcgen05 when eligible, otherwise the SIMT kernel
int gemm(const GemmArgs &g, int compute, cudaStream_t stream);

// conv_tc.cu: implicit-GEMM forward pass of a conv layer (no patch matrix); inside an open split scope
bool conv_tc_eligible(int compute, int T, int B, int F, int C, int kt, int kf, int st, int sf, int N);
int conv_tc_fwd(const float *x, int x_pitch, const float *w, int ldw, const float *bias, float *y, int ldc,
                int T, int B, int F, int C, int kt, int kf, int st, int sf, int To, int Fo, int pt, int pf, int N_real,
                int np, int act, float cutoff, float drop_rate, uint32_t seed, cudaStream_t stream);
int conv_tc_wgrad(const float *x, int x_pitch, const float *dz, int ldz, float *dw, int ldw,
                  int T, int B, int F, int C, int kt, int kf, int st, int sf, int To, int Fo, int pt, int pf,
                  int np, cudaStream_t stream);
bool conv_tc_dgrad_eligible(int C, int kf, int st, int sf, int x_pitch);
int conv_tc_dgrad(const float *dz, int ldz, const float *w, int ldw, float *dx, int x_pitch,
                  int T, int B, int F, int C, int kt, int kf, int st, int sf, int To, int Fo, int pt, int pf,
                  int np, cudaStream_t stream);

// pointwise.cu
int colsum(const float *x, int M, int N, int ld, float *out, cudaStream_t stream);
int mask_inplace(float *dy, const float *y, size_t M, int N, int act, float cutoff, float drop_rate,
                 uint32_t seed, cudaStream_t stream);
// dst [rows, 64] = src [rows, n] with zero columns n .. 63;  dst [rows, n] = src [rows, 64] (+ bias[n])
int pad_cols64(const float *src, int rows, int n, float *dst, cudaStream_t stream);
int compact_cols64(const float *src, int rows, int n, const float *bias, float *dst, cudaStream_t stream);
// mask + column sums + bf16 pieces of dz in one pass (inside an open split scope; dy itself is left untouched)
int mask_colsum_split(const float *dy, const float *y, int M, int N, int act, float cutoff, float drop_rate, uint32_t seed,
                      int np, float *db, cudaStream_t stream);

}  // namespace ctcasr
