// lstm_tc.cuh — persistent tcgen05 LSTM recurrence (lstm_tc.cu).
#pragma once
#include "common.cuh"

namespace ctcasr {

bool lstm_tc_eligible(int T, int B, int H, int cell);
size_t lstm_tc_workspace_bytes(int B, int H);
// gates [T*B, 2*4H] holds P = x Wx + b on entry and the gate activations on exit;
// cstate [T*B, 2H]; y [T*B, 2H].
int lstm_tc_fwd(const int *seq_len, const float *wh, float *gates, float *cstate, float *y,
                int T, int B, int H, int use_len, float forget_bias, void *ws, cudaStream_t stream);
// gates holds activations on entry and dz on exit.  dbias [8H] (optional) receives the column sums of dz when
// the kernel computes them itself; *dbias_done says whether it did (otherwise the caller sums the columns).
int lstm_tc_bwd(const int *seq_len, const float *wh, float *gates, const float *cstate, const float *dy,
                float *dbias, int *dbias_done, int T, int B, int H, int use_len, void *ws, cudaStream_t stream);

}  // namespace ctcasr
