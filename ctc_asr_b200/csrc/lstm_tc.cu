// lstm_tc.cu — persistent tcgen05 LSTM recurrence (forward and backward) for sm_100a.
//
// Replaces the T-step while-loop of tfc.rnn.stack_bidirectional_dynamic_rnn / CudnnLSTM
// (asr/model.py:176-183, :194-215; cell = TF LSTMCell, gate order i, j, f, o) for one layer, both
// directions at once.  The input projection is hoisted (rnn.cu); this kernel runs the strictly
// sequential part  z_t = P_t + h_{t-1} Wh  ->  gates -> (c_t, h_t)  and its reverse.
//
// One cooperative launch per layer and pass; grid = 2 directions x H/32 CTAs, one CTA per SM,
// all T steps inside the kernel (no per-step launches).  CTA (d, c) owns hidden units
// [32c, 32c+32) of direction d for the whole sequence:
//   forward   D[128 gate rows (4 gates x 32 units), 32 batch] = Wp_slice[128, H] . h_{t-1}^T
//             (swap-AB: the gate rows fill the MMA M dimension, the batch is N)
//   backward  D[32 batch (+96 unused lanes), 32 units] = dz_{t+1}[32, 4H] . Wh[units, 4H]^T
// Arithmetic: bf16x3 — weights are pre-split into two bf16 pieces (same bytes as fp32), h / dz are
// split by the epilogue that produces them, three tcgen05.mma kind::f16 products per k-step
// accumulate in fp32 TMEM (error ~2^-16 per product; see gemm_tc.cu).
// Warp roles: 0-3 cell math (TMEM -> registers -> smem gate exchange -> c/h update, c kept in
// registers across steps), 4 weight-tile TMA producer (runs ahead across step boundaries: weights
// do not depend on the recurrence), 5 state-tile TMA producer (gated by the per-direction step
// barrier), 6 MMA issuer.
// Cross-CTA step barrier: every CTA publishes its slice of h_t (bf16 pieces, global memory) and
// bumps a per-direction counter (release); the state producer acquires counter >= CPD * step.
// The recurrent weights (2 x 64 MiB in two bf16 pieces at H = 2048) do not fit on chip and are
// streamed every step: the kernel is bound by L2/HBM weight streaming, not by the tensor pipe.
#include "lstm_tc.cuh"
#include "rec_common.cuh"

#include <stdlib.h>

namespace ctcasr {
namespace lstm {

using rec::BK;
using rec::sigmoidf_;
using rec::split2;
using rec::wait_counter;
using rec::signal_counter;
using rec::make_map;

constexpr int UPC = 32;                 // hidden units per CTA
constexpr int NB = 32;                  // batch rows per MMA (N forward, real M rows backward)
constexpr int NTHREADS = 384;           // 12 warps: 0-3 TMEM readers + cell math, 4 weight tiles, 5 state tiles,
                                        //           6-7 MMA issuers, 8-11 cell math
constexpr int CELL_L = CTCASR_CELL_LSTM, CELL_G = CTCASR_CELL_GRU;

// ---- smem ring: per stage A = PIECES x [128 x 64] bf16 (16 KB each), B = PIECES x [32 x 64] (4 KB each).
// PIECES = 2: bf16x3 arithmetic (operands split hi + lo, three products); PIECES = 1: plain bf16 operands, one
// product (compute = 'bf16', BASELINE cfg3) — half the weight bytes, so most of a CTA's slice stays in tensor memory.
constexpr int F_A_PIECE = 128 * BK * 2, F_B_PIECE = NB * BK * 2;
// NBT = batch rows per launch = MMA N (x2 with the hi / lo pieces of the state stacked): 32, or 64 so that a batch of 64
// (BASELINE cfg4) shares ONE weight stream per time step instead of running as two 32-row launches.
template <int PIECES, int NBT = NB> struct Ring {
    static constexpr int BP = NBT * BK * 2;                                  // one piece of a state tile: 4 KB / 8 KB
    static constexpr int STAGE = PIECES * (F_A_PIECE + BP);                  // 40 KB / 20 KB (48 KB at NBT = 64)
    static constexpr int NSTAGE = PIECES == 2 ? (NBT == 64 ? 4 : 5) : 9;
    static constexpr int ACC_COLS = 2 * PIECES * NBT;                        // two issuers x (PIECES x NBT) accumulator columns
    static constexpr int KRES_MAX = (512 - ACC_COLS) / (PIECES * 32);        // weight k-blocks resident in tensor memory: 6 / 14 (4 at NBT = 64)
    static constexpr int XCH = 4 * NBT * UPC * 4;                            // gate exchange [4][NBT b][32 u] fp32
    static constexpr int BARS = 1024;
    static constexpr int SMEM = NSTAGE * STAGE + XCH + 1024 + BARS;
};
// PIECES = 1 runs two rings whose stages hold GK = 2 consecutive k-blocks: NSA stages of streamed weight tiles (2 x 16
// KB; the k-blocks resident in tensor memory never enter it, so the producer prefetches a whole ring of the next
// step's tiles during the serial tail of this one) and NSB stages of state tiles (2 x 4 KB).  A step visits the
// groups in the order of Params::order: pairs of streamed groups alternate with pairs of resident ones, so that the
// weight ring refills while the issuers work on resident groups.  Two k-blocks per barrier because a (satisfied) mbarrier wait and
// a commit cost the issuing thread as much as three of its 45-cycle MMAs (tools/ubench/mma_loop.cu).
// (One shared ring, or all state tiles in one burst with a 5-stage weight ring, kept too few bytes in flight: the step
// was bound by TMA latency per ring turn, 7.2 of 12.6 us at H = 2048.)
constexpr int GK = 2, NSA = 5, NSB = 6;
constexpr int SPLIT_SMEM = NSA * GK * F_A_PIECE + NSB * GK * F_B_PIECE + Ring<1>::XCH + 1024 + Ring<1>::BARS;
// The two MMA issuers take alternate k-blocks.  Each owns a private sub-ring of stages (issuer 0 the first
// ceil(n/2) stages, issuer 1 the rest) so that every full/empty barrier has its phases consumed by ONE thread in
// order.  With a shared ring of odd depth the successive fills of a stage alternate between the issuers,
// and an issuer that runs ahead (TMA completions are out of order: L2 hits overtake HBM misses) can reach
// a stage whose previous fill has not landed yet — its parity wait then passes on the older phase and the
// MMA reads a stale tile (seen as rare wrong results followed by a hung step barrier at H = 2048).
struct SubRing {
    int first, n, stage;
    uint32_t phase;
    __device__ SubRing(int issuer, int nstage) : first(issuer ? (nstage + 1) / 2 : 0), n(issuer ? nstage / 2 : (nstage + 1) / 2), stage(0), phase(0) {}
    __device__ int slot() const { return first + stage; }
    __device__ void advance() { if (++stage == n) { stage = 0; phase ^= 1; } }
};
// ---- single-CTA backward (LSTM, two pieces; shapes the cluster kernel does not take): per stage Z = 2 pieces x
// [32 x 64] (dz), W = 2 x [32 x 64] (weights); the MMA reads the Z tile as 128 rows (16 KB), so Z tiles are
// spaced 16 KB apart and the ring is padded at the end
constexpr int B_Z_SLOT = 128 * BK * 2, B_W_PIECE = UPC * BK * 2;
constexpr int B_STAGE = 2 * B_W_PIECE;                          // weights: 8 KB per stage
constexpr int B_NSTAGE = 6;
constexpr int B_ZSTAGE = 2 * 4096;                              // dz: 2 pieces x 4 KB actually loaded
constexpr int B_XCH = NB * (UPC + 1) * 4;
constexpr int B_SMEM = B_NSTAGE * (B_STAGE + B_ZSTAGE) + B_Z_SLOT /*over-read pad*/ + B_XCH + 1024 + 256;

struct Params {
    int T, B, BS, H, CPD, use_len;       // B: batch rows of this launch (<= 32); BS: batch rows per frame in the buffers
    float forget_bias;
    const int *seq_len;
    float *gates;            // [T*B, 2GH]  fwd: P -> activations;  bwd: activations -> dz
    float *cstate;           // [T*B, 2H]   LSTM: c;  GRU: q = h Rn + b_rn
    float *y;                // [T*B, 2H]   fwd out; GRU bwd reads h_{t-1} from it
    const float *dy;         // [T*B, 2H]   (bwd in)
    float *dzr;              // GRU bwd: [T*B, 2*3H] gradient wrt h R (n columns scaled by r), the dWh operand
    const float *bias_rn;    // GRU: [2, H] recurrent bias of the candidate gate
    __nv_bfloat16 *xbuf;     // fwd: hbuf [PIECES][2 dirs][2 parity][32][H]; bwd: dzbuf [PIECES][2][2][32][GH]
    unsigned int *counters;  // [2] step counters, [2] = error flag
    unsigned long long *trace;  // optional [grid][64 steps][8 slots] globaltimer stamps (tools/lstm_trace.py)
    int kres;                // weight k-blocks [0, kres) of every CTA stay resident in tensor memory
    const __nv_bfloat16 *wpack; // packed weights [PIECES][rows][K] the resident share is read from
    int wrows, wk;           // rows and K (row pitch) of wpack
    int stagger_ns;          // start delay of direction 1
    int kb_keep;             // weight k-blocks [0, kb_keep) are loaded with L2 evict_last, the rest evict_first
    int nunits, nsg;         // PIECES = 1: groups per step; the first nsg group ids are streamed groups, the rest resident
    unsigned char order[24]; //   group id visited at position u (issuer u & 1)
    float *dbias;            // bwd (cluster kernel): [2GH (+2H GRU)] column sums of dz, or null
    int db_accum;            // add to dbias instead of storing (batch slices after the first)
};

__device__ __forceinline__ void stamp(const Params &p, int step, int slot)
{
    if (p.trace && step < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[((size_t)blockIdx.x * 64 + step) * 8 + slot] = t;
    }
}
// The two directions are independent recurrences that share the memory system: starting one of them
// half a step late makes one direction stream weights while the other is in its serial phase
// (MMA tail, cell math, fence, step barrier) instead of both pulling at the same time.
__device__ __forceinline__ void stagger_wait(int ns)
{
    if (ns <= 0) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)ns);
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }      // single-CTA backward kernel
__device__ __forceinline__ void cell_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }     // 8 cell-math warps

// ================================================ forward =========================================
// CELL = LSTM: rows of a CTA tile = 4 gates (i, j, f, o) x 32 units.  CELL = GRU (cuDNN formulation, gate order
// r, z, n): 3 gates x 32 units, the fourth 32-row group of the tile is zero weights; the recurrent product of the
// candidate gate is kept apart from its input projection (n = tanh(P_n + r (h Rn + b_rn))).
template <int CELL, int PIECES, int NBT>
__global__ void __launch_bounds__(NTHREADS, 1)
gated_fwd_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapH, const Params p)
{
    using R = Ring<PIECES, NBT>;
    static_assert(NBT == 32 || (NBT == 64 && PIECES == 2), "64-row batch tiles: two-piece kernels only");
    constexpr int G = CELL == CELL_G ? 3 : 4;
    constexpr bool FULLB = PIECES == 1;
    constexpr int STAGE = R::STAGE;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    constexpr int NSTAGE = FULLB ? NSA : R::NSTAGE;
    const uint32_t bbuf = smem_base + (uint32_t)NSA * GK * F_A_PIECE;               // FULLB: ring of state tiles
    const uint32_t xch_base = FULLB ? bbuf + (uint32_t)NSB * GK * F_B_PIECE : smem_base + NSTAGE * STAGE;
    float *zs = reinterpret_cast<float *>(smem_gen + (xch_base - smem_base));        // [4][NBT b][32 u]
    const uint32_t bar_base = xch_base + R::XCH;
    auto fullA = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (10 + s); };
    auto fullB = [&](int s) { return bar_base + 8u * (20 + s); };                  // FULLB only: state-tile ring
    auto emptyB = [&](int s) { return bar_base + 8u * (32 + s); };
    const uint32_t tfull = bar_base + 8u * 44, tempty = tfull + 8;
    const uint32_t tmem_slot = tempty + 8;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
    auto a_addr = [&](int s, int pc) { return FULLB ? smem_base + s * GK * F_A_PIECE + pc * F_A_PIECE : smem_base + s * STAGE + pc * F_A_PIECE; };   // FULLB: pc = k-block of the group
    auto b_addr = [&](int s, int pc) { return smem_base + s * STAGE + PIECES * F_A_PIECE + pc * R::BP; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x / p.CPD, c = blockIdx.x % p.CPD;
    const int T = p.T, B = p.B, H = p.H, KB = H / BK;
    const bool TWO_ACC = (PIECES == 1 ? p.nunits : KB) > 1;      // the second issuer had work: add its accumulator

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(fullA(s), FULLB ? 1 : 2); ptx::mbar_init(empty(s), 1); }   // full: weight + state producer
        if (FULLB) for (int s = 0; s < NSB; ++s) { ptx::mbar_init(fullB(s), 1); ptx::mbar_init(emptyB(s), 1); }
        ptx::mbar_init(tfull, 2); ptx::mbar_init(tempty, 4);
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { ptx::tma_prefetch_desc(&mapW); ptx::tma_prefetch_desc(&mapH); }
    const uint32_t tmem_cols = p.kres > 0 ? 512u : (uint32_t)R::ACC_COLS;   // accumulators (+ resident weights)
    if (warp == 6) ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;
    const uint32_t tmem_w = tmem_d + R::ACC_COLS;         // resident weights: k-block kb, piece pc at column (kb*PIECES+pc)*32
    if (warp < 4 && p.kres > 0) {
        // my gate row's weights for k in [0, 64*kres): 32 columns (= 64 bf16) per k-block and piece
        const int row = (d * p.CPD + c) * 128 + warp * 32 + lane;
        for (int kb = 0; kb < p.kres; ++kb)
            for (int pc = 0; pc < PIECES; ++pc) {
                const uint4 *src = reinterpret_cast<const uint4 *>(p.wpack + ((size_t)pc * p.wrows + row) * p.wk + (size_t)kb * BK);
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = __ldg(src + j);
                    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
                }
                ptx::tmem_st32(tmem_w + ((uint32_t)(warp * 32) << 16) + (uint32_t)((kb * PIECES + pc) * 32), r);
            }
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (warp == 4) {
        // ---- weight tiles: independent of the recurrence, runs ahead across step boundaries ----
        if (lane == 0) {
            SubRing ring[2] = {SubRing(0, NSTAGE), SubRing(1, NSTAGE)};
            const int row0 = (d * p.CPD + c) * 128;
            const uint64_t keep = ptx::policy_evict_last(), stream = ptx::policy_evict_first();
            for (int i = 0; i < T; ++i)
                if (FULLB) {
                    for (int u = 0; u < p.nunits; ++u) {
                        const int gid = p.order[u];
                        if (gid >= p.nsg) continue;                 // resident group: no ring traffic at all
                        SubRing &r = ring[u & 1];
                        const int stage = r.slot(), cnt = min(GK, KB - p.kres - GK * gid);
                        ptx::mbar_wait(empty(stage), r.phase ^ 1);
                        ptx::mbar_expect_tx(fullA(stage), (uint32_t)cnt * F_A_PIECE);
                        for (int q2 = 0; q2 < cnt; ++q2) {
                            const int kb = p.kres + GK * gid + q2;
                            const uint64_t pol = kb - p.kres < p.kb_keep ? keep : stream;
                            ptx::tma_load_3d_hint(a_addr(stage, q2), &mapW, kb * BK, row0, 0, fullA(stage), pol);
                        }
                        r.advance();
                    }
                    stamp(p, i, 6);
                } else
                for (int m = 0; m < KB; ++m) {
                    const int kb = m;
                    SubRing &r = ring[m & 1];
                    const int stage = r.slot();
                    ptx::mbar_wait(empty(stage), r.phase ^ 1);
                    if (kb < p.kres) {
                        ptx::mbar_arrive(fullA(stage));     // resident in tensor memory: nothing to load, keep the phases in step
                    } else {
                        ptx::mbar_expect_tx(fullA(stage), PIECES * F_A_PIECE);
                        const uint64_t pol = kb - p.kres < p.kb_keep ? keep : stream;
                        ptx::tma_load_3d_hint(a_addr(stage, 0), &mapW, kb * BK, row0, 0, fullA(stage), pol);   // all pieces in one box
                    }
                    r.advance();
                    if (m == KB - 1) stamp(p, i, 6);
                }
        }
    } else if (warp == 5) {
        // ---- h_{t-1} tiles: gated by the step barrier of this direction ------------------------
        if (lane == 0) {
            SubRing ring[2] = {SubRing(0, NSTAGE), SubRing(1, NSTAGE)};
            SubRing ringB[2] = {SubRing(0, NSB), SubRing(1, NSB)};
            if (d == 1) stagger_wait(p.stagger_ns);
            for (int i = 0; i < T; ++i) {
                wait_counter(p.counters + d, (unsigned)(p.CPD * i), p.counters + 2);
                ptx::fence_proxy_async();
                stamp(p, i, 0);
                const int row0 = (d * 2 + (i & 1)) * NBT;
                if (FULLB) {        // own ring, GK tiles of 4 KB per stage, in the step's visiting order
                    for (int u = 0; u < p.nunits; ++u) {
                        const int gid = p.order[u];
                        const int kb0 = gid < p.nsg ? p.kres + GK * gid : GK * (gid - p.nsg);
                        const int cnt = min(GK, (gid < p.nsg ? KB : p.kres) - kb0);
                        SubRing &r = ringB[u & 1];
                        const int stage = r.slot();
                        ptx::mbar_wait(emptyB(stage), r.phase ^ 1);
                        ptx::mbar_expect_tx(fullB(stage), (uint32_t)cnt * F_B_PIECE);
                        for (int q2 = 0; q2 < cnt; ++q2)
                            ptx::tma_load_3d(bbuf + (uint32_t)(stage * GK + q2) * F_B_PIECE, &mapH, (kb0 + q2) * BK, row0, 0, fullB(stage));
                        r.advance();
                    }
                } else {
                    for (int kb = 0; kb < KB; ++kb) {
                        SubRing &r = ring[kb & 1];
                        const int stage = r.slot();
                        ptx::mbar_wait(empty(stage), r.phase ^ 1);
                        ptx::mbar_expect_tx(fullA(stage), PIECES * R::BP);
                        ptx::tma_load_3d(b_addr(stage, 0), &mapH, kb * BK, row0, 0, fullA(stage));   // all pieces in one box
                        r.advance();
                    }
                }
                stamp(p, i, 1);
            }
        }
    } else if (warp == 6 || warp == 7) {
        // ---- two MMA issuers: even / odd k-blocks into separate accumulators.  The per-k-block cost
        // of one issuing thread (barrier wait + descriptor set-up + small MMAs + commit) is what paces
        // the streaming phase, so it is split over two threads; the epilogue adds the accumulators.
        if (lane == 0) {
            const int me = warp - 6;
            const uint32_t idesc32 = ptx::make_idesc_bf16(128, NBT, 0, 0), idesc64 = ptx::make_idesc_bf16(128, 2 * NBT, 0, 0);     // N = NBT, 2 NBT
            const uint32_t acc = tmem_d + (uint32_t)(me * PIECES * NBT);
            uint32_t tphase = 0;
            SubRing ring(me, NSTAGE), ringB(me, NSB);
            for (int i = 0; i < T; ++i) {
                ptx::mbar_wait(tempty, tphase ^ 1);
                ptx::tc_fence_after();
                // a unit of issue = one k-block, or (FULLB) a group of GK k-blocks behind one pair of barriers
                const int NU = FULLB ? p.nunits : KB;
                for (int u = me; u < NU; u += 2) {
                    const int stage = ring.slot(), stageB = ringB.slot();
                    const int gid = FULLB ? p.order[u] : 0;
                    const bool ringed = !FULLB || gid < p.nsg;              // this unit's weights go through the ring
                    const int kb0 = FULLB ? (ringed ? p.kres + GK * gid : GK * (gid - p.nsg)) : u;
                    const int cnt = FULLB ? min(GK, (ringed ? KB : p.kres) - kb0) : 1;
                    if (FULLB) ptx::mbar_wait(fullB(stageB), ringB.phase);
                    if (ringed) ptx::mbar_wait(fullA(stage), ring.phase);
                    ptx::tc_fence_after();
                  for (int q2 = 0; q2 < cnt; ++q2) {
                    const int kb = kb0 + q2;
                    const int first = u == me && q2 == 0;
                    const uint64_t bd = ptx::make_smem_desc(FULLB ? bbuf + (uint32_t)(stageB * GK + q2) * F_B_PIECE : b_addr(stage, 0), 16, 1024, 2);
                    if (PIECES == 2) {
                        // bf16x3 with 2 MMAs per k-step: the two pieces of h are consecutive rows of one K-major tile,
                        // so  A_hi x [B_hi; B_lo]  is ONE N = 64 MMA (columns 0-31: hi*hi, 32-63: hi*lo) and
                        // A_lo x B_hi  accumulates into columns 0-31; the epilogue adds the column groups.
                        if (kb < p.kres) {          // A from tensor memory: 8 columns per 16-wide k-step
                            const uint32_t ta_hi = tmem_w + (uint32_t)((kb * 2 + 0) * 32), ta_lo = ta_hi + 32;
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j) {
                                ptx::mma_bf16_ts(acc, ta_hi + 8 * j, bd + (uint64_t)(2 * j), idesc64, !(first && j == 0));
                                ptx::mma_bf16_ts(acc, ta_lo + 8 * j, bd + (uint64_t)(2 * j), idesc32, 1);
                            }
                        } else {
                            const uint64_t ad_hi = ptx::make_smem_desc(a_addr(stage, 0), 16, 1024, 2);
                            const uint64_t ad_lo = ptx::make_smem_desc(a_addr(stage, 1), 16, 1024, 2);
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j) {
                                ptx::mma_bf16(acc, ad_hi + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc64, !(first && j == 0));
                                ptx::mma_bf16(acc, ad_lo + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc32, 1);
                            }
                        }
                    } else {
                        if (kb < p.kres) {
                            const uint32_t ta = tmem_w + (uint32_t)(kb * 32);
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j)
                                ptx::mma_bf16_ts(acc, ta + 8 * j, bd + (uint64_t)(2 * j), idesc32, !(first && j == 0));
                        } else {
                            const uint64_t ad = ptx::make_smem_desc(a_addr(stage, FULLB ? q2 : 0), 16, 1024, 2);
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j)
                                ptx::mma_bf16(acc, ad + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc32, !(first && j == 0));
                        }
                    }
                  }
                    if (ringed) { ptx::mma_commit(empty(stage)); ring.advance(); }
                    if (FULLB) { ptx::mma_commit(emptyB(stageB)); ringB.advance(); }
                }
                ptx::mma_commit(tfull);
                if (me == 0) stamp(p, i, 2);
                tphase ^= 1;
            }
        }
    } else if (warp < 4 || warp >= 8) {
        // ---- cell math.  Warps 0-3 (warp = gate) also read the accumulators (LSTM: and add the input projection);
        // then all 8 cell warps own (unit, CPT batch rows) cells with the state kept in registers across steps.
        constexpr int NH = NBT / 32, CPT = NBT / 8;                          // 32-row halves of the batch tile; cells per thread
        const bool reader = warp < 4;
        const int g = warp, ul = lane;
        const int tid = threadIdx.x;
        const int e = reader ? tid : tid - 128;                              // 0..255
        const int cu = e & 31, bg = e >> 5;                                  // cell ownership: rows bg*CPT .. bg*CPT+CPT-1
        const int ucol = c * UPC;                                            // first unit of this CTA
        float sreg[CPT];                                                     // LSTM: c;  GRU: h
        int lenv[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) { const int b = bg * CPT + j; sreg[j] = 0.f; lenv[j] = b < B ? (p.use_len ? min(p.seq_len[b], T) : T) : 0; }
        const float brn = CELL == CELL_G ? p.bias_rn[d * H + ucol + cu] : 0.f;
        uint32_t tphase = 0;
        const size_t GW = (size_t)2 * G * H;                                 // gates row pitch
        for (int i = 0; i < T; ++i) {
            const int tt = d == 0 ? i : T - 1 - i;
            float pin[CELL == CELL_G ? 3 * CPT : 1];                         // GRU: input projections of my cells
            if (CELL == CELL_G) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int b = bg * CPT + j;
                    const float *prow = p.gates + ((size_t)tt * p.BS + b) * GW + (size_t)d * G * H + ucol + cu;
#pragma unroll
                    for (int q = 0; q < 3; ++q) pin[j * 3 + q] = b < B ? __ldg(prow + (size_t)q * H) : 0.f;
                }
            }
            if (reader) {
                // LSTM: pre-activations of my gate row for all batch rows (independent of the recurrence)
                float pz[NH][NB];
#pragma unroll
                for (int hf = 0; hf < NH; ++hf) {
                    const float *prow = p.gates + (size_t)tt * p.BS * GW + (size_t)d * G * H + (size_t)g * H + ucol + ul;
#pragma unroll
                    for (int b = 0; b < NB; ++b) pz[hf][b] = (CELL == CELL_L && hf * NB + b < B) ? __ldg(prow + (size_t)(hf * NB + b) * GW) : 0.f;
                }
                ptx::mbar_wait(tfull, tphase);
                if (tid == 0) stamp(p, i, 3);
                ptx::tc_fence_after();
                if (g < G) {
                    // accumulator columns of issuer m: [m PIECES NBT, +NBT) = hi*hi + lo*hi (or the single product), then NBT of hi*lo
                    const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll
                    for (int hf = 0; hf < NH; ++hf) {
                        uint32_t r[32], r3[32];
                        ptx::tmem_ld32(lane_base + hf * NB, r);                                 // issuer 0
                        ptx::tmem_ld32(lane_base + PIECES * NBT + hf * NB, r3);                 // issuer 1 (odd k-blocks)
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int b = 0; b < NB; ++b) pz[hf][b] += __uint_as_float(r[b]) + (TWO_ACC ? __uint_as_float(r3[b]) : 0.f);
                        if (PIECES == 2) {
                            ptx::tmem_ld32(lane_base + NBT + hf * NB, r);                       // hi*lo of both issuers
                            ptx::tmem_ld32(lane_base + PIECES * NBT + NBT + hf * NB, r3);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int b = 0; b < NB; ++b) pz[hf][b] += __uint_as_float(r[b]) + (TWO_ACC ? __uint_as_float(r3[b]) : 0.f);
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty);
                tphase ^= 1;
                if (g < G) {
#pragma unroll
                    for (int hf = 0; hf < NH; ++hf)
#pragma unroll
                        for (int b = 0; b < NB; ++b) zs[(g * NBT + hf * NB + b) * UPC + ul] = pz[hf][b];
                }
            }
            cell_bar();
            __nv_bfloat16 *hb = p.xbuf + ((size_t)(d * 2 + ((i + 1) & 1)) * NBT) * H + ucol + cu;
            const size_t piece = (size_t)2 * 2 * NBT * H;
            float o_g[CPT][4], o_s[CPT], o_h[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int b = bg * CPT + j;
                const bool live = tt < lenv[j];
                float h = 0.f;
                o_g[j][0] = o_g[j][1] = o_g[j][2] = o_g[j][3] = 0.f;
                if (CELL == CELL_L) {
                    const float zi = zs[(0 * NBT + b) * UPC + cu], zj = zs[(1 * NBT + b) * UPC + cu];
                    const float zf = zs[(2 * NBT + b) * UPC + cu], zo = zs[(3 * NBT + b) * UPC + cu];
                    if (live) {
                        const float gi = sigmoidf_(zi), gj = rec::tanhf_(zj), gf = sigmoidf_(zf + p.forget_bias), go = sigmoidf_(zo);
                        sreg[j] = gf * sreg[j] + gi * gj;
                        h = go * rec::tanhf_(sreg[j]);
                        o_g[j][0] = gi; o_g[j][1] = gj; o_g[j][2] = gf; o_g[j][3] = go;
                    }
                    o_s[j] = sreg[j];
                } else {
                    const float ur = zs[(0 * NBT + b) * UPC + cu], uz = zs[(1 * NBT + b) * UPC + cu], un = zs[(2 * NBT + b) * UPC + cu];
                    o_s[j] = 0.f;
                    if (live) {
                        const float q = un + brn;
                        const float gr = sigmoidf_(pin[j * 3 + 0] + ur), gz = sigmoidf_(pin[j * 3 + 1] + uz);
                        const float gn = rec::tanhf_(pin[j * 3 + 2] + gr * q);
                        h = (1.f - gz) * gn + gz * sreg[j];
                        sreg[j] = h;
                        o_g[j][0] = gr; o_g[j][1] = gz; o_g[j][2] = gn;
                        o_s[j] = q;
                    }
                }
                o_h[j] = h;
                // h_t is the only thing the other CTAs wait for: publish it first
                if (PIECES == 2) {
                    __nv_bfloat16 hi, lo;
                    split2(h, hi, lo);
                    hb[(size_t)b * H] = hi;
                    hb[piece + (size_t)b * H] = lo;
                } else {
                    hb[(size_t)b * H] = __float2bfloat16_rn(h);
                }
            }
            if (tid == 0) stamp(p, i, 4);
            // every writing thread orders its own h pieces for the other CTAs' TMA (async proxy) reads: a single
            // fence by the signalling thread after the barrier (the grid.sync idiom) is NOT enough here —
            // measured: nondeterministic results and rare hangs at H = 2048
            rec::fence_release_gpu();
            ptx::fence_proxy_async();
            cell_bar();
            if (tid == 0) { signal_counter(p.counters + d); stamp(p, i, 5); }
            // bulk stores (activations for the backward pass, layer output) after the signal
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int b = bg * CPT + j;
                if (b < B) {
                    float *grow = p.gates + ((size_t)tt * p.BS + b) * GW + (size_t)d * G * H + ucol + cu;
#pragma unroll
                    for (int q = 0; q < G; ++q) grow[(size_t)q * H] = o_g[j][q];
                    p.cstate[((size_t)tt * p.BS + b) * 2 * H + (size_t)d * H + ucol + cu] = o_s[j];
                    p.y[((size_t)tt * p.BS + b) * 2 * H + (size_t)d * H + ucol + cu] = o_h[j];
                }
            }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 6) ptx::tmem_dealloc(tmem_d, tmem_cols);
}

// ================================================ backward ========================================
__global__ void __launch_bounds__(NTHREADS, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapZ, const Params p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    // layout: [W ring: B_NSTAGE x 8 KB][Z ring: B_NSTAGE x 8 KB][16 KB over-read pad][exchange][barriers]
    const uint32_t w_base = smem_base, z_base = smem_base + B_NSTAGE * B_STAGE;
    const uint32_t xch_off = B_NSTAGE * (B_STAGE + B_ZSTAGE) + B_Z_SLOT;
    float *dhs = reinterpret_cast<float *>(smem_gen + xch_off);                   // [32 b][33]
    const uint32_t bar_base = smem_base + xch_off + B_XCH;
    auto fullW = [&](int s) { return bar_base + 8u * s; };
    auto fullZ = [&](int s) { return bar_base + 8u * (B_NSTAGE + s); };
    auto empty = [&](int s) { return bar_base + 8u * (2 * B_NSTAGE + s); };
    const uint32_t tfull = bar_base + 8u * (3 * B_NSTAGE), tempty = tfull + 8;
    const uint32_t tmem_slot = tempty + 8;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
    auto w_addr = [&](int s, int pc) { return w_base + s * B_STAGE + pc * B_W_PIECE; };
    auto z_addr = [&](int s, int pc) { return z_base + s * B_ZSTAGE + pc * 4096; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x / p.CPD, c = blockIdx.x % p.CPD;
    const int T = p.T, B = p.B, H = p.H, KB = 4 * H / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < B_NSTAGE; ++s) { ptx::mbar_init(fullW(s), 1); ptx::mbar_init(fullZ(s), 1); ptx::mbar_init(empty(s), 1); }
        ptx::mbar_init(tfull, 1); ptx::mbar_init(tempty, 1);
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { ptx::tma_prefetch_desc(&mapW); ptx::tma_prefetch_desc(&mapZ); }
    if (warp == 6) ptx::tmem_alloc(tmem_slot, 32);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;

    if (warp == 4) {
        if (lane == 0) {        // weight rows of my 32 units: Wh[d][u][0..4H), K-major as stored
            int stage = 0; uint32_t phase = 0;
            const int row0 = d * H + c * UPC;
            for (int n = 0; n < T; ++n)
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(empty(stage), phase ^ 1);
                    ptx::mbar_expect_tx(fullW(stage), 2 * B_W_PIECE);
                    ptx::tma_load_3d(w_addr(stage, 0), &mapW, kb * BK, row0, 0, fullW(stage));
                    ptx::tma_load_3d(w_addr(stage, 1), &mapW, kb * BK, row0, 1, fullW(stage));
                    if (++stage == B_NSTAGE) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 5) {
        if (lane == 0) {        // dz of the step processed before (all 4H gate columns of this direction)
            int stage = 0; uint32_t phase = 0;
            for (int n = 0; n < T; ++n) {
                wait_counter(p.counters + d, (unsigned)(p.CPD * n), p.counters + 2);
                ptx::fence_proxy_async();
                const int row0 = (d * 2 + (n & 1)) * NB;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(empty(stage), phase ^ 1);
                    ptx::mbar_expect_tx(fullZ(stage), 2 * 4096);
                    ptx::tma_load_3d(z_addr(stage, 0), &mapZ, kb * BK, row0, 0, fullZ(stage));
                    ptx::tma_load_3d(z_addr(stage, 1), &mapZ, kb * BK, row0, 1, fullZ(stage));
                    if (++stage == B_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 6) {
        if (lane == 0) {
            // A = dz [128 lanes of which 32 batch rows are real], B = weights [32 units], K = 4H
            const uint32_t idesc = ptx::make_idesc_bf16(128, UPC, 0, 0);
            constexpr int PA[3] = {0, 0, 1}, PB[3] = {0, 1, 0};
            int stage = 0; uint32_t phase = 0, tphase = 0;
            for (int n = 0; n < T; ++n) {
                ptx::mbar_wait(tempty, tphase ^ 1);
                ptx::tc_fence_after();
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(fullW(stage), phase);
                    ptx::mbar_wait(fullZ(stage), phase);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint64_t ad = ptx::make_smem_desc(z_addr(stage, PA[q]), 16, 1024, 2);
                        const uint64_t bd = ptx::make_smem_desc(w_addr(stage, PB[q]), 16, 1024, 2);
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j)
                            ptx::mma_bf16(tmem_d, ad + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc, (kb | q | j) != 0);
                    }
                    ptx::mma_commit(empty(stage));
                    if (++stage == B_NSTAGE) { stage = 0; phase ^= 1; }
                }
                ptx::mma_commit(tfull);
                tphase ^= 1;
            }
        }
    } else if (warp < 4) {
        const int tid = threadIdx.x, cu = tid & 31, bg = tid >> 5;
        const int ucol = c * UPC;
        float dcreg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dcreg[j] = 0.f;
        int len8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int b = bg * 8 + j; len8[j] = b < B ? (p.use_len ? min(p.seq_len[b], T) : T) : 0; }
        uint32_t tphase = 0;
        const size_t GW = (size_t)8 * H;
        for (int n = 0; n < T; ++n) {
            const int i = T - 1 - n;                       // processing step of the forward pass
            const int tt = d == 0 ? i : T - 1 - i;
            const int tp = d == 0 ? tt - 1 : tt + 1;       // frame processed before tt in the forward pass
            // everything the cell needs that does not depend on the recurrence
            float gi[8], gj[8], gf[8], go[8], cc[8], cp[8], dyv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = bg * 8 + j;
                gi[j] = gj[j] = gf[j] = go[j] = cc[j] = cp[j] = dyv[j] = 0.f;
                if (b < B) {
                    const float *grow = p.gates + ((size_t)tt * p.BS + b) * GW + (size_t)d * 4 * H + ucol + cu;
                    gi[j] = grow[0]; gj[j] = grow[H]; gf[j] = grow[2 * (size_t)H]; go[j] = grow[3 * (size_t)H];
                    const size_t so = (size_t)d * H + ucol + cu;
                    cc[j] = p.cstate[((size_t)tt * p.BS + b) * 2 * H + so];
                    if (i > 0) cp[j] = p.cstate[((size_t)tp * p.BS + b) * 2 * H + so];
                    dyv[j] = p.dy[((size_t)tt * p.BS + b) * 2 * H + so];
                }
            }
            ptx::mbar_wait(tfull, tphase);
            ptx::tc_fence_after();
            if (warp == 0) {                                // lanes 0..31 of TMEM hold the 32 batch rows
                uint32_t r[32];
                ptx::tmem_ld32(tmem_d, r);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
#pragma unroll
                for (int u = 0; u < UPC; ++u) dhs[lane * (UPC + 1) + u] = __uint_as_float(r[u]);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty);
            }
            tphase ^= 1;
            epi_bar();
            __nv_bfloat16 *zb = p.xbuf + ((size_t)(d * 2 + ((n + 1) & 1)) * NB) * 4 * H + ucol + cu;
            const size_t piece = (size_t)2 * 2 * NB * 4 * H;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = bg * 8 + j;
                const bool live = tt < len8[j];
                float dzi = 0.f, dzj = 0.f, dzf = 0.f, dzo = 0.f;
                if (live) {
                    const float dh = dyv[j] + dhs[b * (UPC + 1) + cu];
                    const float tc = rec::tanhf_(cc[j]);
                    const float dc = dh * go[j] * (1.f - tc * tc) + dcreg[j];
                    dzi = dc * gj[j] * gi[j] * (1.f - gi[j]);
                    dzj = dc * gi[j] * (1.f - gj[j] * gj[j]);
                    dzf = dc * cp[j] * gf[j] * (1.f - gf[j]);
                    dzo = dh * tc * go[j] * (1.f - go[j]);
                    dcreg[j] = dc * gf[j];
                } else {
                    dcreg[j] = 0.f;
                }
                if (b < B) {
                    float *grow = p.gates + ((size_t)tt * p.BS + b) * GW + (size_t)d * 4 * H + ucol + cu;
                    grow[0] = dzi; grow[H] = dzj; grow[2 * (size_t)H] = dzf; grow[3 * (size_t)H] = dzo;
                }
                const float dzv[4] = {dzi, dzj, dzf, dzo};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    __nv_bfloat16 hi, lo;
                    split2(dzv[q], hi, lo);
                    zb[(size_t)b * 4 * H + (size_t)q * H] = hi;
                    zb[piece + (size_t)b * 4 * H + (size_t)q * H] = lo;
                }
            }
            rec::fence_release_gpu();
            ptx::fence_proxy_async();
            epi_bar();
            if (tid == 0) signal_counter(p.counters + d);
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 6) ptx::tmem_dealloc(tmem_d, 32);
}

// ================================== backward, 4-CTA cluster split-K ================================
// dh_{t-1}[b, u] = sum over the G*H gate columns of dz_t[b, :] Wh[u, :].  A cluster of 4 CTAs owns 128
// hidden units; CTA q contracts the K-quarter [q GH/4, (q+1) GH/4) with a full M = 128 tile
//   D_q[128 units, 32 batch] = Wh[units, K-quarter] . dz_t[:, K-quarter]^T
// (the same pipeline shape as the forward kernel), then warp w of every CTA ships its 32 unit rows
// to CTA w of the cluster through distributed shared memory, and CTA w adds the four partials and
// runs the cell backward for those 32 units.  Per step and CTA (LSTM, H = 2048): 1 MB of weights (streamed), 256 KB
// of dz (L2), 384 MMAs.  GRU: dz_t here is the gradient wrt h R (the n columns scaled by r).
template <int CELL, int PIECES, int NBT>
__global__ void __launch_bounds__(NTHREADS, 1)
gated_bwd_cluster_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapZ, const Params p)
{
    using R = Ring<PIECES, NBT>;
    static_assert(NBT == 32 || (NBT == 64 && PIECES == 2), "64-row batch tiles: two-piece kernels only");
    constexpr int G = CELL == CELL_G ? 3 : 4;
    constexpr bool FULLB = PIECES == 1;
    constexpr int STAGE = R::STAGE;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    constexpr int NSTAGE = FULLB ? NSA : R::NSTAGE;
    const uint32_t bbuf = smem_base + (uint32_t)NSA * GK * F_A_PIECE;               // FULLB: ring of state tiles
    const uint32_t xch_base = FULLB ? bbuf + (uint32_t)NSB * GK * F_B_PIECE : smem_base + NSTAGE * STAGE;
    const float *slots = reinterpret_cast<const float *>(smem_gen + (xch_base - smem_base));   // [4][NBT b][32 u]
    const uint32_t bar_base = xch_base + R::XCH;
    auto fullA = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (10 + s); };
    auto fullB = [&](int s) { return bar_base + 8u * (20 + s); };                  // FULLB only: state-tile ring
    auto emptyB = [&](int s) { return bar_base + 8u * (32 + s); };
    const uint32_t tfull = bar_base + 8u * 44, tempty = tfull + 8, xfull = tempty + 8;
    const uint32_t tmem_slot = xfull + 8;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
    auto a_addr = [&](int s, int pc) { return FULLB ? smem_base + s * GK * F_A_PIECE + pc * F_A_PIECE : smem_base + s * STAGE + pc * F_A_PIECE; };   // FULLB: pc = k-block of the group
    auto b_addr = [&](int s, int pc) { return smem_base + s * STAGE + PIECES * F_A_PIECE + pc * R::BP; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = (int)ptx::cluster_ctarank();                  // K-quarter of this CTA
    const int cid = blockIdx.x >> 2;
    const int UBD = p.H / 128;                                  // unit blocks per direction
    const int d = cid / UBD, ub = cid % UBD;
    const int T = p.T, B = p.B, H = p.H, GH = G * H, KQ = GH / 4, KB = KQ / BK;
    const bool TWO_ACC = (PIECES == 1 ? p.nunits : KB) > 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(fullA(s), FULLB ? 1 : 2); ptx::mbar_init(empty(s), 1); }   // full: weight + state producer
        if (FULLB) for (int s = 0; s < NSB; ++s) { ptx::mbar_init(fullB(s), 1); ptx::mbar_init(emptyB(s), 1); }
        ptx::mbar_init(tfull, 2); ptx::mbar_init(tempty, 4); ptx::mbar_init(xfull, 4);
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { ptx::tma_prefetch_desc(&mapW); ptx::tma_prefetch_desc(&mapZ); }
    const uint32_t tmem_cols = p.kres > 0 ? 512u : (uint32_t)R::ACC_COLS;
    if (warp == 6) ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // every CTA's barriers exist before anyone arrives remotely
    ptx::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;
    const uint32_t tmem_w = tmem_d + R::ACC_COLS;         // resident weights, as in the forward kernel
    if (warp < 4 && p.kres > 0) {
        const int row = d * H + ub * 128 + warp * 32 + lane;
        for (int kb = 0; kb < p.kres; ++kb)
            for (int pc = 0; pc < PIECES; ++pc) {
                const uint4 *src = reinterpret_cast<const uint4 *>(p.wpack + ((size_t)pc * p.wrows + row) * p.wk + (size_t)q * KQ + (size_t)kb * BK);
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = __ldg(src + j);
                    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
                }
                ptx::tmem_st32(tmem_w + ((uint32_t)(warp * 32) << 16) + (uint32_t)((kb * PIECES + pc) * 32), r);
            }
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (warp == 4) {
        if (lane == 0) {        // Wh[d][128 units of the cluster][K-quarter], K-major as stored
            SubRing ring[2] = {SubRing(0, NSTAGE), SubRing(1, NSTAGE)};
            const int row0 = d * H + ub * 128;
            const uint64_t keep = ptx::policy_evict_last(), stream = ptx::policy_evict_first();
            for (int n = 0; n < T; ++n)
                if (FULLB) {
                    for (int u = 0; u < p.nunits; ++u) {
                        const int gid = p.order[u];
                        if (gid >= p.nsg) continue;                 // resident group: no ring traffic at all
                        SubRing &r = ring[u & 1];
                        const int stage = r.slot(), cnt = min(GK, KB - p.kres - GK * gid);
                        ptx::mbar_wait(empty(stage), r.phase ^ 1);
                        ptx::mbar_expect_tx(fullA(stage), (uint32_t)cnt * F_A_PIECE);
                        for (int q2 = 0; q2 < cnt; ++q2) {
                            const int kb = p.kres + GK * gid + q2;
                            const uint64_t pol = kb - p.kres < p.kb_keep ? keep : stream;
                            ptx::tma_load_3d_hint(a_addr(stage, q2), &mapW, q * KQ + kb * BK, row0, 0, fullA(stage), pol);
                        }
                        r.advance();
                    }
                    
                } else
                for (int m = 0; m < KB; ++m) {
                    const int kb = m;
                    SubRing &r = ring[m & 1];
                    const int stage = r.slot();
                    ptx::mbar_wait(empty(stage), r.phase ^ 1);
                    if (kb < p.kres) {
                        ptx::mbar_arrive(fullA(stage));     // resident in tensor memory
                    } else {
                        ptx::mbar_expect_tx(fullA(stage), PIECES * F_A_PIECE);
                        const uint64_t pol = kb - p.kres < p.kb_keep ? keep : stream;
                        ptx::tma_load_3d_hint(a_addr(stage, 0), &mapW, q * KQ + kb * BK, row0, 0, fullA(stage), pol);   // all pieces
                    }
                    r.advance();
                }
        }
    } else if (warp == 5) {
        if (lane == 0) {        // dz of the step processed before, my K-quarter of the gate columns, all batch rows
            SubRing ring[2] = {SubRing(0, NSTAGE), SubRing(1, NSTAGE)};
            SubRing ringB[2] = {SubRing(0, NSB), SubRing(1, NSB)};
            if (d == 1) stagger_wait(p.stagger_ns);
            for (int n = 0; n < T; ++n) {
                wait_counter(p.counters + d, (unsigned)(p.CPD * n), p.counters + 2);
                ptx::fence_proxy_async();
                const int row0 = (d * 2 + (n & 1)) * NBT;
                if (FULLB) {        // own ring, GK tiles of 4 KB per stage, in the step's visiting order
                    for (int u = 0; u < p.nunits; ++u) {
                        const int gid = p.order[u];
                        const int kb0 = gid < p.nsg ? p.kres + GK * gid : GK * (gid - p.nsg);
                        const int cnt = min(GK, (gid < p.nsg ? KB : p.kres) - kb0);
                        SubRing &r = ringB[u & 1];
                        const int stage = r.slot();
                        ptx::mbar_wait(emptyB(stage), r.phase ^ 1);
                        ptx::mbar_expect_tx(fullB(stage), (uint32_t)cnt * F_B_PIECE);
                        for (int q2 = 0; q2 < cnt; ++q2)
                            ptx::tma_load_3d(bbuf + (uint32_t)(stage * GK + q2) * F_B_PIECE, &mapZ, q * KQ + (kb0 + q2) * BK, row0, 0, fullB(stage));
                        r.advance();
                    }
                } else {
                    for (int kb = 0; kb < KB; ++kb) {
                        SubRing &r = ring[kb & 1];
                        const int stage = r.slot();
                        ptx::mbar_wait(empty(stage), r.phase ^ 1);
                        ptx::mbar_expect_tx(fullA(stage), PIECES * R::BP);
                        ptx::tma_load_3d(b_addr(stage, 0), &mapZ, q * KQ + kb * BK, row0, 0, fullA(stage));   // all pieces
                        r.advance();
                    }
                }
            }
        }
    } else if (warp == 6 || warp == 7) {
        if (lane == 0) {        // two MMA issuers (even / odd k-blocks, separate accumulators), see the forward kernel
            const int me = warp - 6;
            const uint32_t idesc32 = ptx::make_idesc_bf16(128, NBT, 0, 0), idesc64 = ptx::make_idesc_bf16(128, 2 * NBT, 0, 0);     // N = NBT, 2 NBT
            const uint32_t acc = tmem_d + (uint32_t)(me * PIECES * NBT);
            uint32_t tphase = 0;
            SubRing ring(me, NSTAGE), ringB(me, NSB);
            for (int n = 0; n < T; ++n) {
                ptx::mbar_wait(tempty, tphase ^ 1);
                ptx::tc_fence_after();
                // a unit of issue = one k-block, or (FULLB) a group of GK k-blocks behind one pair of barriers
                const int NU = FULLB ? p.nunits : KB;
                for (int u = me; u < NU; u += 2) {
                    const int stage = ring.slot(), stageB = ringB.slot();
                    const int gid = FULLB ? p.order[u] : 0;
                    const bool ringed = !FULLB || gid < p.nsg;              // this unit's weights go through the ring
                    const int kb0 = FULLB ? (ringed ? p.kres + GK * gid : GK * (gid - p.nsg)) : u;
                    const int cnt = FULLB ? min(GK, (ringed ? KB : p.kres) - kb0) : 1;
                    if (FULLB) ptx::mbar_wait(fullB(stageB), ringB.phase);
                    if (ringed) ptx::mbar_wait(fullA(stage), ring.phase);
                    ptx::tc_fence_after();
                  for (int q2 = 0; q2 < cnt; ++q2) {
                    const int kb = kb0 + q2;
                    const int first = u == me && q2 == 0;
                    const uint64_t bd = ptx::make_smem_desc(FULLB ? bbuf + (uint32_t)(stageB * GK + q2) * F_B_PIECE : b_addr(stage, 0), 16, 1024, 2);
                    if (PIECES == 2) {
                        if (kb < p.kres) {
                            const uint32_t ta_hi = tmem_w + (uint32_t)((kb * 2 + 0) * 32), ta_lo = ta_hi + 32;
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j) {
                                ptx::mma_bf16_ts(acc, ta_hi + 8 * j, bd + (uint64_t)(2 * j), idesc64, !(first && j == 0));
                                ptx::mma_bf16_ts(acc, ta_lo + 8 * j, bd + (uint64_t)(2 * j), idesc32, 1);
                            }
                        } else {
                            const uint64_t ad_hi = ptx::make_smem_desc(a_addr(stage, 0), 16, 1024, 2);
                            const uint64_t ad_lo = ptx::make_smem_desc(a_addr(stage, 1), 16, 1024, 2);
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j) {
                                ptx::mma_bf16(acc, ad_hi + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc64, !(first && j == 0));
                                ptx::mma_bf16(acc, ad_lo + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc32, 1);
                            }
                        }
                    } else {
                        if (kb < p.kres) {
                            const uint32_t ta = tmem_w + (uint32_t)(kb * 32);
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j)
                                ptx::mma_bf16_ts(acc, ta + 8 * j, bd + (uint64_t)(2 * j), idesc32, !(first && j == 0));
                        } else {
                            const uint64_t ad = ptx::make_smem_desc(a_addr(stage, FULLB ? q2 : 0), 16, 1024, 2);
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j)
                                ptx::mma_bf16(acc, ad + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc32, !(first && j == 0));
                        }
                    }
                  }
                    if (ringed) { ptx::mma_commit(empty(stage)); ring.advance(); }
                    if (FULLB) { ptx::mma_commit(emptyB(stageB)); ringB.advance(); }
                }
                ptx::mma_commit(tfull);
                tphase ^= 1;
            }
        }
    } else if (warp < 4 || warp >= 8) {
        constexpr int NH = NBT / 32, CPT = NBT / 8, NP = CPT / 4;          // 32-row halves; cells per thread, in passes of 4
        const bool reader = warp < 4;                                       // warps 0-3 also ship the accumulator rows
        const int tid = threadIdx.x;
        const int e = reader ? tid : tid - 128;                             // 0..255
        const int cu = e & 31, bg = e >> 5;                                 // cell ownership: rows bg*CPT .. bg*CPT+CPT-1
        const int ucol = ub * 128 + q * UPC;                                // the 32 units whose cells this CTA owns
        float carry[CPT];                                                   // LSTM: dc carried to the previous frame; GRU: z * dh
        float dbacc[4] = {0.f, 0.f, 0.f, 0.f};                              // bias gradient: sum of dz over time and my rows (GRU [3]: b_rn)
        int lenv[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) { const int b = bg * CPT + j; carry[j] = 0.f; lenv[j] = b < B ? (p.use_len ? min(p.seq_len[b], T) : T) : 0; }
        uint32_t tphase = 0;
        const size_t GW = (size_t)2 * GH;
        // destination of my TMEM rows: CTA `warp` of the cluster, slot q, [b][lane]
        const uint32_t remote_slot = ptx::mapa(xch_base + (uint32_t)(q * NBT * UPC) * 4u, (uint32_t)(warp & 3));
        const uint32_t remote_bar = ptx::mapa(xfull, (uint32_t)(warp & 3));
        for (int n = 0; n < T; ++n) {
            const int i = T - 1 - n;
            const int tt = d == 0 ? i : T - 1 - i;
            const int tp = d == 0 ? tt - 1 : tt + 1;
            // everything the cell needs that does not depend on the recurrence, for one pass of 4 cells
            float ga[4][4], sc[4], sp[4], dyv[4];            // gate activations; LSTM: c_t, c_{t-1};  GRU: q_t, h_{t-1}
            auto load_pass = [&](int ps) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int b = bg * CPT + ps * 4 + j;
                    ga[j][0] = ga[j][1] = ga[j][2] = ga[j][3] = sc[j] = sp[j] = dyv[j] = 0.f;
                    if (b < B) {
                        const float *grow = p.gates + ((size_t)tt * p.BS + b) * GW + (size_t)d * GH + ucol + cu;
#pragma unroll
                        for (int kk = 0; kk < G; ++kk) ga[j][kk] = grow[(size_t)kk * H];
                        const size_t so = (size_t)d * H + ucol + cu;
                        sc[j] = p.cstate[((size_t)tt * p.BS + b) * 2 * H + so];
                        if (i > 0) sp[j] = CELL == CELL_L ? p.cstate[((size_t)tp * p.BS + b) * 2 * H + so] : p.y[((size_t)tp * p.BS + b) * 2 * H + so];
                        dyv[j] = p.dy[((size_t)tt * p.BS + b) * 2 * H + so];
                    }
                }
            };
            load_pass(0);
            if (reader) {
                ptx::mbar_wait(tfull, tphase);
                ptx::tc_fence_after();
                // accumulator columns of issuer m: [m PIECES NBT, +NBT) = hi*hi + lo*hi (or the single product), then NBT of hi*lo
                const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);      // rows = units 32*warp + lane of the block
#pragma unroll
                for (int hf = 0; hf < NH; ++hf) {
                    float part[NB];
                    uint32_t r[32], r3[32];
                    ptx::tmem_ld32(lane_base + hf * NB, r);
                    ptx::tmem_ld32(lane_base + PIECES * NBT + hf * NB, r3);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int b = 0; b < NB; ++b) part[b] = __uint_as_float(r[b]) + (TWO_ACC ? __uint_as_float(r3[b]) : 0.f);
                    if (PIECES == 2) {
                        ptx::tmem_ld32(lane_base + NBT + hf * NB, r);
                        ptx::tmem_ld32(lane_base + PIECES * NBT + NBT + hf * NB, r3);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int b = 0; b < NB; ++b) part[b] += __uint_as_float(r[b]) + (TWO_ACC ? __uint_as_float(r3[b]) : 0.f);
                    }
#pragma unroll
                    for (int b = 0; b < NB; ++b)
                        ptx::st_cluster_f32(remote_slot + (uint32_t)((hf * NB + b) * UPC + lane) * 4u, part[b]);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) { ptx::mbar_arrive(tempty); ptx::mbar_arrive_remote(remote_bar); }
            }
            ptx::mbar_wait_cluster(xfull, tphase);                          // the four partials of my units have landed
            tphase ^= 1;
            __nv_bfloat16 *zb = p.xbuf + ((size_t)(d * 2 + ((n + 1) & 1)) * NBT) * GH + ucol + cu;
            const size_t piece = (size_t)2 * 2 * NBT * GH;
            float o_dz[4][4], o_zr[4];                                      // o_zr (GRU): dn_pre * r, the n column of dzr
            auto store_pass = [&](int ps) {             // fp32 dz for the weight-gradient GEMMs
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int b = bg * CPT + ps * 4 + j;
                    if (b < B) {
                        float *grow = p.gates + ((size_t)tt * p.BS + b) * GW + (size_t)d * GH + ucol + cu;
#pragma unroll
                        for (int kk = 0; kk < G; ++kk) grow[(size_t)kk * H] = o_dz[j][kk];
                        if (CELL == CELL_G) {
                            float *zrow = p.dzr + ((size_t)tt * p.BS + b) * GW + (size_t)d * GH + ucol + cu;
                            zrow[0] = o_dz[j][0]; zrow[H] = o_dz[j][1]; zrow[2 * (size_t)H] = o_zr[j];
                        }
                    }
                }
            };
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) {
                if (ps > 0) load_pass(ps);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int jj = ps * 4 + j, b = bg * CPT + jj;
                    const bool live = tt < lenv[jj];
                    o_dz[j][0] = o_dz[j][1] = o_dz[j][2] = o_dz[j][3] = 0.f;
                    o_zr[j] = 0.f;
                    if (live) {
                        const int o = b * UPC + cu;
                        const float dh = dyv[j] + ((slots[o] + slots[NBT * UPC + o]) + (slots[2 * NBT * UPC + o] + slots[3 * NBT * UPC + o]));
                        if (CELL == CELL_L) {
                            const float gi = ga[j][0], gj = ga[j][1], gf = ga[j][2], go = ga[j][3];
                            const float tc = rec::tanhf_(sc[j]);
                            const float dc = dh * go * (1.f - tc * tc) + carry[jj];
                            o_dz[j][0] = dc * gj * gi * (1.f - gi);
                            o_dz[j][1] = dc * gi * (1.f - gj * gj);
                            o_dz[j][2] = dc * sp[j] * gf * (1.f - gf);
                            o_dz[j][3] = dh * tc * go * (1.f - go);
                            carry[jj] = dc * gf;
                        } else {
                            const float gr = ga[j][0], gz = ga[j][1], gn = ga[j][2];
                            const float dht = dh + carry[jj];
                            const float dn_pre = dht * (1.f - gz) * (1.f - gn * gn);
                            o_dz[j][0] = dn_pre * sc[j] * gr * (1.f - gr);
                            o_dz[j][1] = dht * (sp[j] - gn) * gz * (1.f - gz);
                            o_dz[j][2] = dn_pre;
                            o_zr[j] = dn_pre * gr;
                            carry[jj] = dht * gz;
                        }
                    } else {
                        carry[jj] = 0.f;
                    }
#pragma unroll
                    for (int g4 = 0; g4 < G; ++g4) {        // the bf16 pieces are what the other CTAs wait for
                        const float v = (CELL == CELL_G && g4 == 2) ? o_zr[j] : o_dz[j][g4];
                        dbacc[g4] += o_dz[j][g4];
                        if (PIECES == 2) {
                            __nv_bfloat16 hi, lo;
                            split2(v, hi, lo);
                            zb[(size_t)b * GH + (size_t)g4 * H] = hi;
                            zb[piece + (size_t)b * GH + (size_t)g4 * H] = lo;
                        } else {
                            zb[(size_t)b * GH + (size_t)g4 * H] = __float2bfloat16_rn(v);
                        }
                    }
                    if (CELL == CELL_G) dbacc[3] += o_zr[j];
                }
                if (NP > 1) store_pass(ps);             // 8 cells per thread: no room to hold the stores back behind the signal
            }
            rec::fence_release_gpu();       // per-thread fences: see the forward kernel
            ptx::fence_proxy_async();
            cell_bar();
            if (tid == 0) signal_counter(p.counters + d);
            if (NP == 1) store_pass(0);                 // after the signal
        }
        if (p.dbias) {
            // column sums of dz for my 32 units x 4 sums: the 8 row groups meet in the (now idle) exchange slots
            float *red = const_cast<float *>(slots);                        // [8 bg][4][32 cu]
            cell_bar();
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) red[(bg * 4 + g4) * UPC + cu] = dbacc[g4];
            cell_bar();
            if (e < 4 * UPC) {
                const int g4 = e >> 5;
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) v += red[(k * 4 + g4) * UPC + cu];
                // LSTM: 4 gate blocks of [2][4H];  GRU: 3 gate blocks of [2][3H], then b_rn [2][H]
                float *dst = (CELL == CELL_G && g4 == 3) ? p.dbias + (size_t)2 * GH + (size_t)d * H + ucol + cu
                                                         : p.dbias + (size_t)d * GH + (size_t)g4 * H + ucol + cu;
                *dst = p.db_accum ? *dst + v : v;
            }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // no CTA exits while a peer may still write into its shared memory
    if (warp == 6) ptx::tmem_dealloc(tmem_d, tmem_cols);
}

// ---- weight pre-packs --------------------------------------------------------------------------------
// forward: Wh fp32 [2][H][G*H] -> Wp bf16 [pieces][2*4H rows][H], row (d, c, g, ul) = gate column
// g*H + 32c + ul of direction d, K (= h index) contiguous: the K-major A operand of the swap-AB MMA.
// A CTA tile always has four 32-row gate groups; with G = 3 (GRU) the fourth stays zero (the buffer is cleared).
__global__ void pack_wh_fwd_kernel(const float *__restrict__ wh, __nv_bfloat16 *__restrict__ wp, int H, int G, int pieces)
{
    __shared__ float tile[32][33];
    const int d = blockIdx.z, k0 = blockIdx.y * 32, col0 = blockIdx.x * 32;     // col in [0, GH)
    const int tx = threadIdx.x, ty = threadIdx.y;                                // 32 x 8
    const size_t GH = (size_t)G * H, R4 = (size_t)4 * H;
    for (int j = ty; j < 32; j += 8) tile[j][tx] = wh[((size_t)d * H + k0 + j) * GH + col0 + tx];
    __syncthreads();
    const size_t piece = (size_t)2 * R4 * H;
    for (int j = ty; j < 32; j += 8) {
        const int col = col0 + j;                  // gate column g*H + u
        const int g = col / H, u = col % H;
        const size_t row = (size_t)d * R4 + (size_t)(u / UPC) * 128 + g * UPC + (u % UPC);
        if (pieces == 2) {
            __nv_bfloat16 hi, lo;
            split2(tile[tx][j], hi, lo);
            wp[row * H + k0 + tx] = hi;
            wp[piece + row * H + k0 + tx] = lo;
        } else {
            wp[row * H + k0 + tx] = __float2bfloat16_rn(tile[tx][j]);
        }
    }
}
// backward: bf16 pieces of Wh viewed as [2H rows][GH]
__global__ void split_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ out, size_t n, int pieces)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (pieces == 2) {
            __nv_bfloat16 hi, lo;
            split2(x[i], hi, lo);
            out[i] = hi;
            out[n + i] = lo;
        } else {
            out[i] = __float2bfloat16_rn(x[i]);
        }
    }
}

static unsigned long long *g_trace = nullptr;
struct WsLayout { size_t wpack, xbuf, counters, total; };
static WsLayout ws_layout(int H)
{
    WsLayout w;
    w.wpack = 0;
    const size_t wbytes = align_up((size_t)2 * 2 * 4 * H * H * 2, 1024);                // 2 pieces x [2*4H][H] bf16
    w.xbuf = wbytes;
    const size_t xbytes = align_up((size_t)2 * 2 * 2 * 64 * 4 * H * 2, 1024);           // dzbuf (64-row tiles) is the larger user
    w.counters = w.xbuf + xbytes;
    w.total = w.counters + 1024;
    return w;
}

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// How many of the streamed weight k-blocks per CTA are loaded with the L2 evict_last policy.  The two
// directions' recurrent weights are 16 H^2 bytes as two bf16 pieces (128 MiB at H = 2048) against ~126 MB of
// L2, so only a share can stay resident across time steps; the rest streams from HBM with evict_first so that
// it does not push the resident share out.  With one piece (64 MiB) everything stays.
// CTCASR_LSTM_L2_KEEP_MB overrides the resident budget.
static int keep_kblocks(int H, int G, int pieces, int KB)
{
    static const double budget_mb = (double)env_int("CTCASR_LSTM_L2_KEEP_MB", 64);
    const double total_mb = 2.0 * H * G * H * 2.0 * pieces / 1048576.0;
    int k = (int)(KB * budget_mb / total_mb);
    return k < 0 ? 0 : (k > KB ? KB : k);
}

// Tensor-memory-resident weight share: the columns next to the accumulators hold the first k-blocks of every
// CTA's weight slice for the whole sequence (6 with two pieces, 14 with one), read by the MMA as a TMEM A operand.
static int resident_kblocks(int KB, int pieces, int nbt = NB)
{
    static const int env = env_int("CTCASR_LSTM_KRES", -1);
    const int kmax = pieces == 2 ? (nbt == 64 ? Ring<2, 64>::KRES_MAX : Ring<2>::KRES_MAX) : Ring<1>::KRES_MAX;
    int kres = env >= 0 ? env : kmax;
    if (kres > kmax) kres = kmax;
    return kres > KB ? KB : kres;
}

// visiting order of the single-piece kernels: group ids 0 .. nsg-1 are the streamed groups (GK k-blocks from kres on),
// nsg .. the resident ones; pairs of streamed groups alternate with pairs of resident groups (each issuer gets one of
// every pair), the longer kind fills the end
static void fill_order(Params &p, int KB)
{
    const int ns = KB - p.kres, nsg = (ns + GK - 1) / GK, nrg = (p.kres + GK - 1) / GK;
    p.nsg = nsg; p.nunits = 0;
    int is = 0, ir = 0;
    while (is < nsg || ir < nrg) {
        for (int k = 0; k < 2 && is < nsg; ++k) p.order[p.nunits++] = (unsigned char)(is++);
        for (int k = 0; k < 2 && ir < nrg; ++k) p.order[p.nunits++] = (unsigned char)(nsg + ir++);
    }
}

static int check_coop(const void *kernel, int smem, int grid)
{
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    CTCASR_CUDA_CHECK(cudaGetDevice(&dev));
    CTCASR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CTCASR_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CTCASR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CTCASR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NTHREADS, smem));
    if (!coop || per_sm * sms < grid)
        return fail(CTCASR_ERR_UNSUPPORTED, "lstm_tc: %d CTAs cannot be co-resident (%d SMs x %d)", grid, sms, per_sm);
    return CTCASR_OK;
}

typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const Params);
static KernelFn fwd_kernel(int cell, int pieces, int nbt)
{
    if (nbt == 64) return cell == CELL_G ? gated_fwd_kernel<CELL_G, 2, 64> : gated_fwd_kernel<CELL_L, 2, 64>;
    if (cell == CELL_G) return pieces == 2 ? gated_fwd_kernel<CELL_G, 2, 32> : gated_fwd_kernel<CELL_G, 1, 32>;
    return pieces == 2 ? gated_fwd_kernel<CELL_L, 2, 32> : gated_fwd_kernel<CELL_L, 1, 32>;
}
static KernelFn bwd_kernel(int cell, int pieces, int nbt)
{
    if (nbt == 64) return cell == CELL_G ? gated_bwd_cluster_kernel<CELL_G, 2, 64> : gated_bwd_cluster_kernel<CELL_L, 2, 64>;
    if (cell == CELL_G) return pieces == 2 ? gated_bwd_cluster_kernel<CELL_G, 2, 32> : gated_bwd_cluster_kernel<CELL_G, 1, 32>;
    return pieces == 2 ? gated_bwd_cluster_kernel<CELL_L, 2, 32> : gated_bwd_cluster_kernel<CELL_L, 1, 32>;
}

}  // namespace lstm

void lstm_tc_set_trace(unsigned long long *buf) { lstm::g_trace = buf; }

bool lstm_tc_eligible(int T, int B, int H, int cell)
{
    if (T < 1 || B < 1 || 2 * (H / lstm::UPC) > 148) return false;
    if (cell == CTCASR_CELL_LSTM) return H >= 64 && H % 64 == 0;
    if (cell == CTCASR_CELL_GRU) return H >= 256 && H % 256 == 0;       // K-quarters of 3H in whole k-blocks
    return false;
}

// Bytes one launch pulls through TMA from L2 / HBM: the weight tiles that are not resident in tensor memory and the
// state tiles of every step, over all CTAs (what the L2-streaming bound of bench.py is computed from).
double lstm_tc_stream_bytes(int T, int H, int cell, int pieces, int backward)
{
    using namespace lstm;
    const int G = cell == CELL_G ? 3 : 4;
    const int KB = backward ? G * H / 4 / BK : H / BK;
    const int kres = (backward && H % 128) ? 0 : resident_kblocks(KB, pieces);
    const double per_step = (double)(KB - kres) * pieces * F_A_PIECE + (double)KB * pieces * F_B_PIECE;
    return per_step * (2.0 * H / UPC) * T;
}

size_t lstm_tc_workspace_bytes(int B, int H)
{
    (void)B;
    if (H < 64 || H % 64) return 0;
    return lstm::ws_layout(H).total + 1024;
}

int lstm_tc_fwd(const int *seq_len, const float *wh, float *gates, float *cstate, float *y, const float *bias_rn,
                int T, int B, int H, int cell, int pieces, int use_len, float forget_bias, void *ws, cudaStream_t stream)
{
    using namespace lstm;
    char *base = reinterpret_cast<char *>(align_up((size_t)(uintptr_t)ws, 1024));
    const WsLayout L = ws_layout(H);
    __nv_bfloat16 *wp = reinterpret_cast<__nv_bfloat16 *>(base + L.wpack);
    __nv_bfloat16 *hbuf = reinterpret_cast<__nv_bfloat16 *>(base + L.xbuf);
    unsigned int *ctr = reinterpret_cast<unsigned int *>(base + L.counters);
    const int G = cell == CELL_G ? 3 : 4;
    const int CPD = H / UPC, grid = 2 * CPD, KB = H / BK;
    // a batch above 32 rows takes the 64-row tile (two-piece arithmetic): one weight stream per step for all of them
    static const int wide_ok = env_int("CTCASR_LSTM_NO_WIDE", 0) == 0;
    const int NBT = (B > NB && pieces == 2 && wide_ok) ? 64 : NB;
    const int smem = pieces == 2 ? (NBT == 64 ? Ring<2, 64>::SMEM : Ring<2>::SMEM) : SPLIT_SMEM;
    KernelFn kernel = fwd_kernel(cell, pieces, NBT);
    static int checked_grid[2][3] = {{0, 0, 0}, {0, 0, 0}};
    int &chk = checked_grid[cell == CELL_G][NBT == 64 ? 2 : pieces - 1];
    if (chk != grid) { int rc = check_coop((const void *)kernel, smem, grid); if (rc) return rc; chk = grid; }

    if (G == 3) CTCASR_CUDA_CHECK(cudaMemsetAsync(wp, 0, (size_t)pieces * 2 * 4 * H * H * 2, stream));     // the unused fourth gate group
    pack_wh_fwd_kernel<<<dim3(G * H / 32, H / 32, 2), dim3(32, 8), 0, stream>>>(wh, wp, H, G, pieces);
    CTCASR_LAUNCH_CHECK();
    CUtensorMap mapW, mapH;
    int rc = make_map(&mapW, wp, H, (uint64_t)2 * 4 * H, 128, pieces, pieces);
    if (rc) return rc;
    rc = make_map(&mapH, hbuf, H, (uint64_t)2 * 2 * NBT, NBT, pieces, pieces);
    if (rc) return rc;
    static const int stagger = env_int("CTCASR_LSTM_STAGGER_NS", 11000);
    // larger batches run as consecutive launches over NBT-row slices of the same buffers
    for (int b0 = 0; b0 < B; b0 += NBT) {
        CTCASR_CUDA_CHECK(cudaMemsetAsync(hbuf, 0, (size_t)pieces * 2 * 2 * NBT * H * 2, stream));  // h_{-1} = 0, padded batch rows = 0
        CTCASR_CUDA_CHECK(cudaMemsetAsync(ctr, 0, 64, stream));
        Params p = {};
        p.T = T; p.B = B - b0 < NBT ? B - b0 : NBT; p.BS = B; p.H = H; p.CPD = CPD; p.use_len = use_len;
        p.forget_bias = forget_bias; p.seq_len = seq_len ? seq_len + b0 : nullptr;
        p.gates = gates + (size_t)b0 * 2 * G * H; p.cstate = cstate + (size_t)b0 * 2 * H; p.y = y + (size_t)b0 * 2 * H;
        p.bias_rn = bias_rn; p.xbuf = hbuf; p.counters = ctr;
        p.kres = resident_kblocks(KB, pieces, NBT);
        p.kb_keep = keep_kblocks(H, 4, pieces, KB);
        p.wpack = wp; p.wrows = 2 * 4 * H; p.wk = H;
        fill_order(p, KB);
        p.trace = g_trace;
        p.stagger_ns = stagger;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        ProfScope prof(PROF_LSTM_FWD, stream);
        CTCASR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, mapW, mapH, p));
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
    }
    return CTCASR_OK;
}

int lstm_tc_bwd(const int *seq_len, const float *wh, float *gates, const float *cstate, const float *y, const float *dy,
                float *dzr, float *dbias, int *dbias_done, int T, int B, int H, int cell, int pieces, int use_len,
                void *ws, cudaStream_t stream)
{
    using namespace lstm;
    char *base = reinterpret_cast<char *>(align_up((size_t)(uintptr_t)ws, 1024));
    const WsLayout L = ws_layout(H);
    __nv_bfloat16 *wq = reinterpret_cast<__nv_bfloat16 *>(base + L.wpack);
    __nv_bfloat16 *zbuf = reinterpret_cast<__nv_bfloat16 *>(base + L.xbuf);
    unsigned int *ctr = reinterpret_cast<unsigned int *>(base + L.counters);
    const int G = cell == CELL_G ? 3 : 4, GH = G * H;
    const int CPD = H / UPC, grid = 2 * CPD;
    const size_t nw = (size_t)2 * H * GH;
    CUtensorMap mapW, mapZ;
    int rc = CTCASR_OK;

    // preferred: 4-CTA cluster split-K kernel (needs H % 128 == 0 and all clusters co-resident)
    static int cluster_ok_grid[2][3] = {{0, 0, 0}, {0, 0, 0}}, cluster_bad_grid = 0, checked_grid = 0;
    static const int wide_ok = env_int("CTCASR_LSTM_NO_WIDE", 0) == 0;
    int NBT = (B > NB && pieces == 2 && H % 128 == 0 && wide_ok) ? 64 : NB;     // 64-row batch tile: see lstm_tc_fwd
    static const bool no_cluster = getenv("CTCASR_LSTM_NO_CLUSTER") != nullptr;
    bool use_cluster = false;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    KernelFn ckernel = bwd_kernel(cell, pieces, NBT);
    const int csmem = pieces == 2 ? (NBT == 64 ? Ring<2, 64>::SMEM : Ring<2>::SMEM) : SPLIT_SMEM;
    if (H % 128 == 0 && cluster_bad_grid != grid && (!no_cluster || cell == CELL_G)) {
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = csmem; cfg.stream = stream;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int &okg = cluster_ok_grid[cell == CELL_G][NBT == 64 ? 2 : pieces - 1];
        if (okg != grid) {
            CTCASR_CUDA_CHECK(cudaFuncSetAttribute(ckernel, cudaFuncAttributeMaxDynamicSharedMemorySize, csmem));
            int nclusters = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, ckernel, &cfg);
            if (e == cudaSuccess && nclusters * 4 >= grid) okg = grid;
            else { cluster_bad_grid = grid; (void)cudaGetLastError(); }
        }
        use_cluster = okg == grid;
    }
    if (!use_cluster) {
        if (cell != CELL_L) return fail(CTCASR_ERR_UNSUPPORTED, "lstm_tc: the GRU backward pass needs the cluster kernel");
        pieces = 2;                                       // the single-CTA kernel is two-piece, 32 rows only
        NBT = NB;
    }
    split_kernel<<<148 * 8, 256, 0, stream>>>(wh, wq, nw, pieces);
    CTCASR_LAUNCH_CHECK();
    rc = make_map(&mapZ, zbuf, (uint64_t)GH, (uint64_t)2 * 2 * NBT, NBT, use_cluster ? pieces : 1, pieces);
    if (rc) return rc;
    if (use_cluster) {
        rc = make_map(&mapW, wq, (uint64_t)GH, (uint64_t)2 * H, 128, pieces, pieces);
    } else {
        if (checked_grid != grid) { rc = check_coop((const void *)lstm_bwd_kernel, B_SMEM, grid); if (rc) return rc; checked_grid = grid; }
        rc = make_map(&mapW, wq, (uint64_t)GH, (uint64_t)2 * H, UPC, 1);
    }
    if (rc) return rc;
    static const int stagger = env_int("CTCASR_LSTM_STAGGER_NS", 11000);
    const int KB = GH / 4 / BK;
    for (int b0 = 0; b0 < B; b0 += NBT) {
        CTCASR_CUDA_CHECK(cudaMemsetAsync(zbuf, 0, (size_t)pieces * 2 * 2 * NBT * GH * 2, stream));  // no recurrent gradient into the last step
        CTCASR_CUDA_CHECK(cudaMemsetAsync(ctr, 0, 64, stream));
        Params p = {};
        p.T = T; p.B = B - b0 < NBT ? B - b0 : NBT; p.BS = B; p.H = H; p.CPD = CPD; p.use_len = use_len; p.forget_bias = 0.f;
        p.seq_len = seq_len ? seq_len + b0 : nullptr;
        p.gates = gates + (size_t)b0 * 2 * GH; p.cstate = const_cast<float *>(cstate) + (size_t)b0 * 2 * H;
        p.y = const_cast<float *>(y) + (size_t)b0 * 2 * H;
        p.dy = dy + (size_t)b0 * 2 * H; p.dzr = dzr ? dzr + (size_t)b0 * 2 * GH : nullptr; p.xbuf = zbuf; p.counters = ctr;
        p.kres = use_cluster ? resident_kblocks(KB, pieces, NBT) : 0; p.wpack = wq; p.wrows = 2 * H; p.wk = GH;
        p.kb_keep = keep_kblocks(H, G, pieces, KB);
        fill_order(p, KB);
        p.dbias = use_cluster ? dbias : nullptr; p.db_accum = b0 > 0;
        p.trace = nullptr;
        p.stagger_ns = stagger;
        ProfScope prof(PROF_LSTM_BWD, stream);
        if (use_cluster) {
            CTCASR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ckernel, mapW, mapZ, p));
        } else {
            void *args[] = {&mapW, &mapZ, &p};
            CTCASR_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)lstm_bwd_kernel, dim3(grid), dim3(NTHREADS), args, B_SMEM, stream));
        }
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
    }
    if (dbias_done) *dbias_done = use_cluster && dbias;      // the cluster kernel sums dz over time itself
    return CTCASR_OK;
}

}  // namespace ctcasr
