// rnn.cu — one bidirectional recurrent layer, forward and backward (C-ABI entry points).
//
// Replaces a layer of tfc.rnn.stack_bidirectional_dynamic_rnn (asr/model.py:176-183) /
// tfc.cudnn_rnn.Cudnn* (asr/model.py:194-215).  Structure on the GPU:
//   1. hoisted input GEMM   P[T*B, 2GH] = X[T*B, in] Wx[in, 2GH] + bias   (both directions at once)
//   2. the recurrence       z_t = P_t + h_{t-1} Wh  ->  cell math, T strictly sequential steps,
//                           fw and bw directions advancing together
//   3. backward: reverse recurrence producing dz in place of the saved activations, then three
//      large GEMMs (dWx = X^T dz, dWh = H_prev^T dz, dX = dz Wx^T) and a column sum (dbias).
// Step 2 has two implementations: the stepwise one in this file (any shape / cell, two launches
// per frame) and the persistent tcgen05 LSTM kernel in lstm_tc.cu (selected when eligible).
#include "gemm.cuh"
#include "rnn.cuh"
#include "lstm_tc.cuh"
#include "rec_tc.cuh"

using namespace ctcasr;

namespace {
struct Reserve { float *gates; float *cstate; float *dzr; size_t bytes; };
Reserve carve_reserve(void *base, int T, int B, int H, int G)
{
    Reserve r;
    const size_t ng = align_up((size_t)T * B * 2 * G * H * sizeof(float), 256);
    const size_t nc = align_up((size_t)T * B * 2 * H * sizeof(float), 256);
    r.gates = reinterpret_cast<float *>(base);
    r.cstate = reinterpret_cast<float *>(reinterpret_cast<char *>(base) + ng);
    r.dzr = reinterpret_cast<float *>(reinterpret_cast<char *>(base) + ng + nc);     // GRU only (G == 3)
    r.bytes = ng + nc + (G == 3 ? ng : 0);
    return r;
}
// stepwise-path scratch at the start of ws: dh_rec [2,B,H], dc_carry [2,B,H], rh [2,B,3H] (GRU)
size_t step_ws_bytes(int B, int H) { return align_up((size_t)(4 + 6) * B * H * sizeof(float), 256); }
}  // namespace

extern "C" size_t ctcasr_birnn_reserve_bytes(int T, int B, int in, int H, int cell)
{
    (void)in;
    return carve_reserve(nullptr, T, B, H, num_gates(cell)).bytes;
}

extern "C" size_t ctcasr_birnn_workspace_bytes(int T, int B, int in, int H, int cell)
{
    (void)T; (void)in;
    // dh_rec + dc_carry (stepwise) and the persistent kernel's h exchange / barrier area
    (void)cell;
    return step_ws_bytes(B, H) + lstm_tc_workspace_bytes(B, H);
}

extern "C" double ctcasr_birnn_stream_bytes(int T, int B, int H, int cell, int compute, int backward)
{
    const int slices = (B + 31) / 32;
    if (compute == CTCASR_COMPUTE_FP32) return 0.0;
    if (lstm_tc_eligible(T, B, H, cell)) return slices * lstm_tc_stream_bytes(T, H, cell, compute == CTCASR_COMPUTE_BF16 ? 1 : 2, backward);
    if (rec_tc_eligible(T, B, H, cell)) return slices * rec_tc_stream_bytes(T, H);
    return 0.0;
}

extern "C" int ctcasr_birnn_fwd(const float *x, const int32_t *seq_len, const float *wx, const float *wh,
                                const float *bias, float *y, void *reserve,
                                int T, int B, int in, int H, int cell, int use_len, float forget_bias,
                                int compute, void *ws, size_t ws_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(x && wx && wh && bias && y && reserve, "birnn_fwd: null pointer");
    CTCASR_REQUIRE(T >= 1 && B >= 1 && in >= 1 && H >= 1, "birnn_fwd: bad dims");
    CTCASR_REQUIRE(!use_len || seq_len, "birnn_fwd: use_len needs seq_len");
    CTCASR_REQUIRE(cell >= 0 && cell <= 3, "birnn_fwd: bad cell %d", cell);
    if (ws_bytes < ctcasr_birnn_workspace_bytes(T, B, in, H, cell)) return fail(CTCASR_ERR_WORKSPACE, "birnn_fwd: workspace too small");
    const int G = num_gates(cell), GH = G * H;
    if (int rcs = gemm_scratch_check(compute, 1, T * B, 2 * GH, in)) return rcs;
    Reserve r = carve_reserve(reserve, T, B, H, G);

    // 1. hoisted input GEMM with the bias folded into the epilogue
    GemmArgs g;
    g.A[0] = x; g.B[0] = wx; g.C[0] = r.gates;
    g.M = T * B; g.N = 2 * GH; g.K = in; g.lda = in; g.ldb = 2 * GH; g.ldc = 2 * GH;
    g.epi.mode = EPI_BIAS_ACT; g.epi.bias = bias; g.epi.act = 0;
    g.precise = cell == CTCASR_CELL_RNN_RELU;      // the ReLU cell has a kink at 0; tanh / LSTM are smooth
    int rc = gemm(g, compute, stream);
    if (rc != CTCASR_OK) return rc;

    // 2. recurrence
    if (compute != CTCASR_COMPUTE_FP32 && lstm_tc_eligible(T, B, H, cell)) {
        char *wsb = reinterpret_cast<char *>(ws) + step_ws_bytes(B, H);
        const int pieces = compute == CTCASR_COMPUTE_BF16 ? 1 : 2;
        return lstm_tc_fwd(seq_len, wh, r.gates, r.cstate, y, bias + 2 * GH, T, B, H, cell, pieces, use_len, forget_bias, wsb, stream);
    }
    if (compute != CTCASR_COMPUTE_FP32 && rec_tc_eligible(T, B, H, cell)) {     // one-gate cells: resident-weight kernel
        char *wsb = reinterpret_cast<char *>(ws) + step_ws_bytes(B, H);
        return rec_tc_fwd(seq_len, wh, r.gates, y, T, B, H, cell, use_len, wsb, stream);
    }
    RnnStep s;
    s.T = T; s.B = B; s.H = H; s.G = G; s.cell = cell; s.use_len = use_len; s.forget_bias = forget_bias;
    s.seq_len = seq_len; s.gates = r.gates; s.cstate = r.cstate; s.y = y; s.dy = nullptr;
    s.dh_rec = nullptr; s.dc_carry = nullptr;
    const bool gru = cell == CTCASR_CELL_GRU;
    s.rh = reinterpret_cast<float *>(ws) + (size_t)4 * B * H; s.dzr = nullptr; s.bias_rn = bias + 2 * GH;
    const bool one_gate = cell == CTCASR_CELL_RNN_TANH || cell == CTCASR_CELL_RNN_RELU;
    for (int i = 0; i < T; ++i) {
        if (i > 0) {
            GemmArgs h;
            h.nz = 2; h.M = B; h.N = GH; h.K = H; h.lda = 2 * H; h.ldb = GH; h.ldc = 2 * GH;
            h.epi.accumulate = 1;
            const int tf = i, tb = T - 1 - i;
            h.A[0] = y + (size_t)(tf - 1) * B * 2 * H;            h.A[1] = y + (size_t)(tb + 1) * B * 2 * H + H;
            h.B[0] = wh;                                          h.B[1] = wh + (size_t)H * GH;
            h.C[0] = r.gates + (size_t)tf * B * 2 * GH;           h.C[1] = r.gates + (size_t)tb * B * 2 * GH + GH;
            if (one_gate) {     // product + cell math in one launch
                StepCell sc;
                sc.mode = 1; sc.cell = cell; sc.use_len = use_len; sc.seq_len = seq_len; sc.t[0] = tf; sc.t[1] = tb;
                sc.y[0] = y + (size_t)tf * B * 2 * H; sc.y[1] = y + (size_t)tb * B * 2 * H + H; sc.ldy = 2 * H;
                rc = step_gemm_cell(h, sc, stream);
                if (rc == CTCASR_OK) continue;
                if (rc != CTCASR_ERR_UNSUPPORTED) return rc;
            }
            if (gru) {      // the candidate gate multiplies h Rn by r: keep h Wh apart from the input projection
                h.epi.accumulate = 0; h.ldc = GH;
                h.C[0] = s.rh; h.C[1] = s.rh + (size_t)B * GH;
            }
            rc = step_gemm(h, stream);
            if (rc != CTCASR_OK) return rc;
        }
        rc = rnn_cell_fwd(s, i, stream);
        if (rc != CTCASR_OK) return rc;
    }
    return CTCASR_OK;
}

extern "C" int ctcasr_birnn_bwd(const float *x, const int32_t *seq_len, const float *wx, const float *wh,
                                const float *y, void *reserve, const float *dy,
                                float *dx, float *dwx, float *dwh, float *dbias,
                                int T, int B, int in, int H, int cell, int use_len,
                                int compute, void *ws, size_t ws_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(x && wx && wh && y && reserve && dy && dwx && dwh && dbias, "birnn_bwd: null pointer");
    CTCASR_REQUIRE(T >= 1 && B >= 1 && in >= 1 && H >= 1, "birnn_bwd: bad dims");
    CTCASR_REQUIRE(cell >= 0 && cell <= 3, "birnn_bwd: bad cell %d", cell);
    if (ws_bytes < ctcasr_birnn_workspace_bytes(T, B, in, H, cell)) return fail(CTCASR_ERR_WORKSPACE, "birnn_bwd: workspace too small");
    const int G = num_gates(cell), GH = G * H;
    Reserve r = carve_reserve(reserve, T, B, H, G);
    int rc, dbias_done = 0;
    // dz is read by three GEMMs (dWx, dWh, dX): its bf16 pieces are made once and shared
    auto pad8 = [](int v) { return (size_t)((v + 7) / 8 * 8); };
    const size_t TB = (size_t)T * B;
    const size_t operands[6] = {TB * pad8(in), TB * pad8(2 * GH), TB * pad8(H), TB * pad8(H), (size_t)in * pad8(2 * GH),
                                cell == CTCASR_CELL_GRU ? TB * pad8(2 * GH) : 0};       // GRU: dWh reads a second dz (dzr)
    SplitScope scope;
    if (int rcs = split_scope_begin(compute, operands, 6)) return rcs;
    scope.open = true;

    if (compute != CTCASR_COMPUTE_FP32 && lstm_tc_eligible(T, B, H, cell)) {
        char *wsb = reinterpret_cast<char *>(ws) + step_ws_bytes(B, H);
        const int pieces = compute == CTCASR_COMPUTE_BF16 ? 1 : 2;
        rc = lstm_tc_bwd(seq_len, wh, r.gates, r.cstate, y, dy, cell == CTCASR_CELL_GRU ? r.dzr : nullptr, dbias, &dbias_done,
                         T, B, H, cell, pieces, use_len, wsb, stream);
        if (rc != CTCASR_OK) return rc;
    } else if (compute != CTCASR_COMPUTE_FP32 && rec_tc_eligible(T, B, H, cell)) {
        char *wsb = reinterpret_cast<char *>(ws) + step_ws_bytes(B, H);
        rc = rec_tc_bwd(seq_len, wh, r.gates, dy, dbias, T, B, H, cell, use_len, wsb, stream);
        if (rc != CTCASR_OK) return rc;
        dbias_done = 1;
    } else {
        RnnStep s;
        s.T = T; s.B = B; s.H = H; s.G = G; s.cell = cell; s.use_len = use_len; s.forget_bias = 0.f;
        s.seq_len = seq_len; s.gates = r.gates; s.cstate = r.cstate; s.y = const_cast<float *>(y); s.dy = dy;
        s.dh_rec = reinterpret_cast<float *>(ws);
        s.dc_carry = s.dh_rec + (size_t)2 * B * H;
        s.rh = nullptr; s.dzr = r.dzr; s.bias_rn = nullptr;
        CTCASR_CUDA_CHECK(cudaMemsetAsync(ws, 0, (size_t)4 * B * H * sizeof(float), stream));
        const bool one_gate = cell == CTCASR_CELL_RNN_TANH || cell == CTCASR_CELL_RNN_RELU;
        bool fused_ok = one_gate;           // frame i's cell math rides in the epilogue of the product that feeds it
        for (int i = T - 1; i >= 0; --i) {
            if (!(fused_ok && i < T - 1)) {
                rc = rnn_cell_bwd(s, i, stream);
                if (rc != CTCASR_OK) return rc;
            }
            if (i > 0) {     // dh_rec[d] = dz_t[d] Wh[d]^T
                GemmArgs h;
                h.nz = 2; h.M = B; h.N = H; h.K = GH; h.tb = 1; h.lda = 2 * GH; h.ldb = GH; h.ldc = H;
                const int tf = i, tb = T - 1 - i;
                const float *dzh = cell == CTCASR_CELL_GRU ? r.dzr : r.gates;
                h.A[0] = dzh + (size_t)tf * B * 2 * GH;           h.A[1] = dzh + (size_t)tb * B * 2 * GH + GH;
                h.B[0] = wh;                                      h.B[1] = wh + (size_t)H * GH;
                h.C[0] = s.dh_rec;                                h.C[1] = s.dh_rec + (size_t)B * H;
                if (fused_ok) {     // writes dz of frame i-1 (processing order) straight into its gates slot
                    GemmArgs f = h;
                    const int pf = i - 1, pb = T - i;
                    f.ldc = 2 * GH;
                    f.C[0] = r.gates + (size_t)pf * B * 2 * GH;   f.C[1] = r.gates + (size_t)pb * B * 2 * GH + GH;
                    StepCell sc;
                    sc.mode = 2; sc.cell = cell; sc.use_len = use_len; sc.seq_len = seq_len; sc.t[0] = pf; sc.t[1] = pb;
                    sc.dy[0] = dy + (size_t)pf * B * 2 * H; sc.dy[1] = dy + (size_t)pb * B * 2 * H + H; sc.ldy = 2 * H;
                    rc = step_gemm_cell(f, sc, stream);
                    if (rc == CTCASR_OK) continue;
                    if (rc != CTCASR_ERR_UNSUPPORTED) return rc;
                    fused_ok = false;       // not eligible (decided at the first product): two launches per frame
                }
                rc = step_gemm(h, stream);
                if (rc != CTCASR_OK) return rc;
            }
        }
    }
    // r.gates now holds dz [T*B, 2GH] (wrt the input-side pre-activations); GRU: r.dzr wrt h Wh
    const float *dzh = cell == CTCASR_CELL_GRU ? r.dzr : r.gates;
    if (!dbias_done) {
        rc = colsum(r.gates, T * B, 2 * GH, 2 * GH, dbias, stream);
        if (rc != CTCASR_OK) return rc;
    }
    if (cell == CTCASR_CELL_GRU && !dbias_done)
        for (int d = 0; d < 2; ++d) {       // b_rn gradient: column sums of the n block of dzr
            rc = colsum(r.dzr + (size_t)d * GH + 2 * H, T * B, H, 2 * GH, dbias + 2 * GH + d * H, stream);
            if (rc != CTCASR_OK) return rc;
        }
    {   // dWx[in, 2GH] = X^T dz
        GemmArgs g;
        g.A[0] = x; g.B[0] = r.gates; g.C[0] = dwx; g.ta = 1;
        g.M = in; g.N = 2 * GH; g.K = T * B; g.lda = in; g.ldb = 2 * GH; g.ldc = 2 * GH;
        rc = gemm(g, compute, stream);
        if (rc != CTCASR_OK) return rc;
    }
    if (T > 1) {   // dWh[d][H, GH] = H_prev^T dz_d ; fw: (y[t-1], dz[t]), bw: (y[t+1], dz[t])
        GemmArgs g;
        g.nz = 2; g.ta = 1; g.M = H; g.N = GH; g.K = (T - 1) * B; g.lda = 2 * H; g.ldb = 2 * GH; g.ldc = GH;
        g.A[0] = y;                              g.B[0] = dzh + (size_t)B * 2 * GH;       g.C[0] = dwh;
        g.A[1] = y + (size_t)B * 2 * H + H;      g.B[1] = dzh + GH;                       g.C[1] = dwh + (size_t)H * GH;
        rc = gemm(g, compute, stream);
        if (rc != CTCASR_OK) return rc;
    } else {
        CTCASR_CUDA_CHECK(cudaMemsetAsync(dwh, 0, (size_t)2 * H * GH * sizeof(float), stream));
    }
    if (dx) {      // dX[T*B, in] = dz Wx^T
        GemmArgs g;
        g.A[0] = r.gates; g.B[0] = wx; g.C[0] = dx; g.tb = 1;
        g.M = T * B; g.N = in; g.K = 2 * GH; g.lda = 2 * GH; g.ldb = 2 * GH; g.ldc = in;
        rc = gemm(g, compute, stream);
        if (rc != CTCASR_OK) return rc;
    }
    return CTCASR_OK;
}
