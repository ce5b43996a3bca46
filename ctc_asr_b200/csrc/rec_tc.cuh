// rec_tc.cuh — persistent tcgen05 recurrence of the one-gate cells (rec_tc.cu).
#pragma once
#include "common.cuh"

namespace ctcasr {

bool rec_tc_eligible(int T, int B, int H, int cell);
size_t rec_tc_workspace_bytes(int H);
double rec_tc_stream_bytes(int T, int H);
// gates [T*B, 2H] holds P = x Wx + b on entry and h on exit (= y)
int rec_tc_fwd(const int *seq_len, const float *wh, float *gates, float *y, int T, int B, int H, int cell, int use_len,
               void *ws, cudaStream_t stream);
// gates holds h on entry and dz on exit; dbias [2H] (optional) receives the column sums of dz
int rec_tc_bwd(const int *seq_len, const float *wh, float *gates, const float *dy, float *dbias,
               int T, int B, int H, int cell, int use_len, void *ws, cudaStream_t stream);

}  // namespace ctcasr
