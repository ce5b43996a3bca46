// lstm_tc.cuh — persistent tcgen05 LSTM recurrence (lstm_tc.cu).
#pragma once
#include "common.cuh"

namespace ctcasr {

bool lstm_tc_eligible(int T, int B, int H, int cell);           // cell: LSTM or GRU
size_t lstm_tc_workspace_bytes(int B, int H);
double lstm_tc_stream_bytes(int T, int H, int cell, int pieces, int backward);
// gates [T*B, 2*G*H] holds P = x Wx + b on entry and the gate activations on exit;
// cstate [T*B, 2H] (LSTM: c; GRU: q = h Rn + b_rn); y [T*B, 2H]; bias_rn [2, H] (GRU).
// pieces: 2 = bf16x3 arithmetic (operands split hi + lo), 1 = plain bf16 operands (compute 'bf16').
int lstm_tc_fwd(const int *seq_len, const float *wh, float *gates, float *cstate, float *y, const float *bias_rn,
                int T, int B, int H, int cell, int pieces, int use_len, float forget_bias, void *ws, cudaStream_t stream);
// gates holds activations on entry and dz on exit; dzr (GRU) [T*B, 2*3H] receives the gradient wrt h R.
// dbias (optional) receives the column sums of dz (GRU: followed by the b_rn sums) when the kernel computes them
// itself; *dbias_done says whether it did (otherwise the caller sums the columns).
int lstm_tc_bwd(const int *seq_len, const float *wh, float *gates, const float *cstate, const float *y, const float *dy,
                float *dzr, float *dbias, int *dbias_done, int T, int B, int H, int cell, int pieces, int use_len,
                void *ws, cudaStream_t stream);

}  // namespace ctcasr
