// rnn.cuh — argument block of one bidirectional recurrent layer (stepwise path).
#pragma once
#include "common.cuh"

namespace ctcasr {

struct RnnStep {
    int T, B, H, G, cell, use_len;
    float forget_bias;
    const int *seq_len;
    float *gates;        // [T*B, 2*G*H]  z -> activations (fwd) -> dz (bwd)
    float *cstate;       // [T*B, 2*H]    LSTM cell state per frame
    float *y;            // [T*B, 2*H]
    const float *dy;     // [T*B, 2*H]
    float *dh_rec;       // [2, B, H]
    float *dc_carry;     // [2, B, H]   LSTM: dc carry; GRU: direct path z * dh into h_{t-1}
    // GRU only (cuDNN formulation, gate order r, z, n)
    float *rh;           // [2, B, 3H]  h_{t-1} Wh of the current step (kept apart from the input projection)
    float *dzr;          // [T*B, 2*3H] gradient wrt h_{t-1} Wh (n columns scaled by r)
    const float *bias_rn; // [2, H]     recurrent bias of the candidate gate; cstate stores q = h Rn + b_rn
};

inline int num_gates(int cell) { return cell == CTCASR_CELL_LSTM ? 4 : (cell == CTCASR_CELL_GRU ? 3 : 1); }

int rnn_cell_fwd(const RnnStep &s, int i, cudaStream_t stream);
int rnn_cell_bwd(const RnnStep &s, int i, cudaStream_t stream);

}  // namespace ctcasr
