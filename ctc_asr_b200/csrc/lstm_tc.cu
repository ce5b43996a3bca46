// placeholder until the persistent tcgen05 LSTM kernel lands
#include "lstm_tc.cuh"
namespace ctcasr {
bool lstm_tc_eligible(int, int, int, int) { return false; }
size_t lstm_tc_workspace_bytes(int, int) { return 0; }
int lstm_tc_fwd(const int *, const float *, float *, float *, float *, int, int, int, int, float, void *, cudaStream_t) { return fail(CTCASR_ERR_UNSUPPORTED, "lstm_tc: not built"); }
int lstm_tc_bwd(const int *, const float *, float *, const float *, const float *, int, int, int, int, void *, cudaStream_t) { return fail(CTCASR_ERR_UNSUPPORTED, "lstm_tc: not built"); }
}
