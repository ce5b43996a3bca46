// rec_tc.cu — persistent tcgen05 recurrence for the one-gate cells (rnn_tanh / rnn_relu), forward and backward.
//
// Replaces the T-step loop of tfc.rnn.stack_bidirectional_dynamic_rnn with BasicRNNCell(tanh)
// (asr/util/tf_contrib.py:189, asr/model.py:176-183) and of CudnnRNNRelu / CudnnRNNTanh (asr/model.py:194-199;
// rnn_relu is the reference's default cell, asr/params.py:48) for one layer, both directions at once.
// The input projection is hoisted (rnn.cu); this kernel runs the strictly sequential part
//     forward    h_t      = act(P_t + h_{t-1} Wh)          backward   dh_{t-1} = dz_t Wh^T,  dz = dh * act'(h)
// ONE launch per layer and pass (the stepwise path needs one launch per frame).
//
// Both passes are the same contraction  D[128 units, 32 batch] = W'[units, K = H] . x^T  with W' = Wh^T (forward)
// or Wh (backward), x = h_{t-1} or dz_t.  A cluster of 4 CTAs owns 128 units of one direction; CTA q contracts
// the K-quarter [q H/4, (q+1) H/4): its weight slice [128 x H/4] in two bf16 pieces is 256 KB at H = 2048 and
// stays ON CHIP for the whole sequence — the first 6 k-blocks in tensor memory (read by the MMA as a TMEM A
// operand), the rest in shared memory — so a time step moves only the K-quarter of h / dz (64 KB per CTA)
// through TMA.  The four partial [128 x 32] accumulators are exchanged through distributed shared memory (warp w
// ships its 32 unit rows to CTA w), CTA w adds them and runs the cell math for its 32 units x 32 batch rows,
// publishes h_t / dz_t as bf16 pieces for the next step and bumps the per-direction step counter.
// Arithmetic: bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM), as in lstm_tc.cu.
#include "rec_tc.cuh"
#include "rec_common.cuh"

namespace ctcasr {
namespace rnn1 {

using rec::BK;
constexpr int NB = 32;                  // batch rows per launch (MMA N; hi and lo pieces stacked: N = 64)
constexpr int UPC = 32;                 // units whose cells one CTA owns
constexpr int NTHREADS = 384;           // warps 0-3 TMEM readers + cells, 4 TMA producer, 6-7 MMA issuers, 8-11 cells
constexpr int A_PIECE = 128 * BK * 2;   // 16 KB: one bf16 piece of a [128 x 64] weight k-block
constexpr int B_TILE = 2 * NB * BK * 2; // 8 KB: both pieces of a [32 x 64] state k-block
constexpr int TMEM_KB = 6;              // weight k-blocks resident in tensor memory (6 x 64 columns + 2 x 64 accumulator columns)
constexpr int MAX_KB = 10;              // k-blocks per CTA (H/4/64): the rest (<= 4) stays in shared memory
constexpr int XCH = 4 * NB * UPC * 4;   // exchange slots [4 sources][32 b][32 u] fp32

struct Params {
    int T, B, BS, H, CPD, use_len, cell;
    const int *seq_len;
    float *gates;               // [T*BS, 2H]  fwd: P -> h;  bwd: h -> dz
    float *y;                   // [T*BS, 2H]  fwd out
    const float *dy;            // [T*BS, 2H]  bwd in
    __nv_bfloat16 *xbuf;        // [2 pieces][2 dirs][2 parity][32][H]   h_t / dz_t exchange
    unsigned int *counters;     // [2] step counters, [2] error flag
    const __nv_bfloat16 *wpack; // [2 pieces][2H rows][H]
    float *dbias;               // bwd: [2H] column sums of dz, or null
    int db_accum;
    unsigned long long *trace;  // optional [grid][64 steps][8 slots] globaltimer stamps (tools/lstm_trace.py)
};

__device__ __forceinline__ void stamp(const Params &p, int step, int slot)
{
    if (p.trace && step < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[((size_t)blockIdx.x * 64 + step) * 8 + slot] = t;
    }
}

__device__ __forceinline__ void cell_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

static size_t smem_bytes(int NKB)
{
    const int nsw = NKB > TMEM_KB ? NKB - TMEM_KB : 0;
    return (size_t)nsw * 2 * A_PIECE + (size_t)NKB * B_TILE + XCH + 256 + 1024;
}

template <bool FWD>
__global__ void __launch_bounds__(NTHREADS, 1)
rnn_rec_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const Params p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const int H = p.H, T = p.T, B = p.B;
    const int KQ = H / 4, NKB = KQ / BK, NSW = NKB > TMEM_KB ? NKB - TMEM_KB : 0;
    const uint32_t w_base = smem_base;                                   // shared-memory resident weight k-blocks
    const uint32_t b_base = w_base + (uint32_t)NSW * 2 * A_PIECE;        // state tiles of the current step
    const uint32_t xch_base = b_base + (uint32_t)NKB * B_TILE;
    const float *slots = reinterpret_cast<const float *>(smem_gen + (xch_base - smem_base));   // [4][32 b][32 u]
    const uint32_t bar_base = xch_base + XCH;
    auto fullB = [&](int kb) { return bar_base + 8u * kb; };
    const uint32_t wres = bar_base + 8u * MAX_KB, bfree = wres + 8, tfull = wres + 16, tempty = wres + 24, xfull = wres + 32;
    const uint32_t tmem_slot = wres + 40;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = (int)ptx::cluster_ctarank();                  // K-quarter of this CTA
    const int cid = blockIdx.x >> 2;
    const int UBD = H / 128;                                    // unit blocks (clusters) per direction
    const int d = cid / UBD, ub = cid % UBD;

    if (threadIdx.x == 0) {
        for (int g = 0; g < (NKB + 3) / 4; ++g) ptx::mbar_init(fullB(g), 1);
        ptx::mbar_init(wres, 1); ptx::mbar_init(bfree, 2);           // bfree / tfull: one commit per MMA issuer
        ptx::mbar_init(tfull, 2); ptx::mbar_init(tempty, 4); ptx::mbar_init(xfull, 4);
        ptx::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { ptx::tma_prefetch_desc(&mapW); ptx::tma_prefetch_desc(&mapX); }
    if (warp == 6) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // every CTA's barriers exist before anyone arrives remotely
    ptx::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;
    const uint32_t tmem_w = tmem_d + 128;       // resident weights: k-block kb, piece pc at column 128 + (kb*2+pc)*32
    if (warp < 4) {
        // my unit row's weights for the first k-blocks of the K-quarter: 32 columns (= 64 bf16) per k-block and piece
        const int row = d * H + ub * 128 + warp * 32 + lane;
        const int nres = NKB < TMEM_KB ? NKB : TMEM_KB;
        for (int kb = 0; kb < nres; ++kb)
            for (int pc = 0; pc < 2; ++pc) {
                const uint4 *src = reinterpret_cast<const uint4 *>(p.wpack + ((size_t)pc * 2 * H + row) * H + (size_t)q * KQ + (size_t)kb * BK);
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = __ldg(src + j);
                    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
                }
                ptx::tmem_st32(tmem_w + ((uint32_t)(warp * 32) << 16) + (uint32_t)((kb * 2 + pc) * 32), r);
            }
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (warp == 4) {
        if (lane == 0) {
            if (NSW > 0) {          // the k-blocks that do not fit in tensor memory: loaded once, never recycled
                ptx::mbar_expect_tx(wres, (uint32_t)NSW * 2 * A_PIECE);
                for (int kb = TMEM_KB; kb < NKB; ++kb)
                    ptx::tma_load_3d(w_base + (uint32_t)(kb - TMEM_KB) * 2 * A_PIECE, &mapW, q * KQ + kb * BK, d * H + ub * 128, 0, wres);
            }
            for (int n = 0; n < T; ++n) {
                rec::wait_counter(p.counters + d, (unsigned)(p.CPD * n), p.counters + 2);
                ptx::fence_proxy_async();
                stamp(p, n, 0);
                if (n > 0) ptx::mbar_wait(bfree, (uint32_t)((n - 1) & 1));      // the previous step's MMAs have read the tiles
                const int row0 = (d * 2 + (n & 1)) * NB;
                // one barrier per group of 4 k-blocks: a satisfied mbarrier wait costs the MMA-issuing thread as much as
                // two of its MMAs (tools/ubench/mma_loop.cu)
                for (int kb = 0; kb < NKB; ++kb) {
                    if ((kb & 3) == 0) ptx::mbar_expect_tx(fullB(kb >> 2), (uint32_t)((NKB - kb < 4 ? NKB - kb : 4) * B_TILE));
                    ptx::tma_load_3d(b_base + (uint32_t)kb * B_TILE, &mapX, q * KQ + kb * BK, row0, 0, fullB(kb >> 2));   // both pieces
                }
                stamp(p, n, 1);
            }
        }
    } else if (warp == 6 || warp == 7) {
        // two MMA issuers (even / odd k-blocks, own accumulators): every MMA of this kernel is a small one whose cost is its
        // ~45-cycle issue, not its math (tools/ubench/mma_loop.cu), and two threads issue in parallel
        if (lane == 0) {
            const int me = warp - 6;
            const uint32_t acc = tmem_d + (uint32_t)(me * 64);
            const uint32_t idesc32 = ptx::make_idesc_bf16(128, NB, 0, 0), idesc64 = ptx::make_idesc_bf16(128, 2 * NB, 0, 0);
            if (NSW > 0) ptx::mbar_wait(wres, 0);
            for (int n = 0; n < T; ++n) {
                ptx::mbar_wait(tempty, (uint32_t)((n & 1) ^ 1));
                ptx::tc_fence_after();
                int group = -1;
                for (int kb = me; kb < NKB; kb += 2) {
                    if ((kb >> 2) != group) {
                        group = kb >> 2;
                        ptx::mbar_wait(fullB(group), (uint32_t)(n & 1));
                        ptx::tc_fence_after();
                    }
                    const bool first = kb == me;
                    // bf16x3 with 2 MMAs per k-step: the two pieces of x are consecutive rows of one K-major tile, so
                    // A_hi x [x_hi; x_lo] is ONE N = 64 MMA (columns 0-31: hi*hi, 32-63: hi*lo) and A_lo x x_hi
                    // accumulates into columns 0-31; the epilogue adds the column groups.
                    const uint64_t bd = ptx::make_smem_desc(b_base + (uint32_t)kb * B_TILE, 16, 1024, 2);
                    if (kb < TMEM_KB) {
                        const uint32_t ta_hi = tmem_w + (uint32_t)((kb * 2 + 0) * 32), ta_lo = ta_hi + 32;
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j) {
                            ptx::mma_bf16_ts(acc, ta_hi + 8 * j, bd + (uint64_t)(2 * j), idesc64, !(first && j == 0));
                            ptx::mma_bf16_ts(acc, ta_lo + 8 * j, bd + (uint64_t)(2 * j), idesc32, 1);
                        }
                    } else {
                        const uint32_t wa = w_base + (uint32_t)(kb - TMEM_KB) * 2 * A_PIECE;
                        const uint64_t ad_hi = ptx::make_smem_desc(wa, 16, 1024, 2), ad_lo = ptx::make_smem_desc(wa + A_PIECE, 16, 1024, 2);
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j) {
                            ptx::mma_bf16(acc, ad_hi + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc64, !(first && j == 0));
                            ptx::mma_bf16(acc, ad_lo + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc32, 1);
                        }
                    }
                }
                ptx::mma_commit(bfree);
                ptx::mma_commit(tfull);
                if (me == 0) stamp(p, n, 2);
            }
        }
    } else if (warp < 4 || warp >= 8) {
        const bool reader = warp < 4;                                       // warps 0-3 also ship the accumulator rows
        const int tid = threadIdx.x;
        const int e = reader ? tid : tid - 128;                             // 0..255
        const int cu = e & 31, bg = e >> 5;                                 // cell ownership: rows bg*4 .. bg*4+3
        const int unit = ub * 128 + q * UPC + cu;                           // the 32 units whose cells this CTA owns
        const size_t col = (size_t)d * H + unit;                            // column in the [., 2H] buffers
        const bool tanh_cell = p.cell == CTCASR_CELL_RNN_TANH;
        float dbacc = 0.f;
        int len4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int b = bg * 4 + j; len4[j] = b < B ? (p.use_len ? min(p.seq_len[b], T) : T) : 0; }
        // destination of my TMEM rows: CTA `warp` of the cluster, slot q, [b][lane]
        const uint32_t remote_slot = ptx::mapa(xch_base + (uint32_t)(q * NB * UPC) * 4u, (uint32_t)(warp & 3));
        const uint32_t remote_bar = ptx::mapa(xfull, (uint32_t)(warp & 3));
        const size_t piece = (size_t)2 * 2 * NB * H;
        for (int n = 0; n < T; ++n) {
            const int i = FWD ? n : T - 1 - n;                              // processing step of the forward pass
            const int tt = d == 0 ? i : T - 1 - i;
            const uint32_t tphase = (uint32_t)(n & 1);
            // everything the cell needs that does not depend on the recurrence
            float a[4], g[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = bg * 4 + j;
                a[j] = g[j] = 0.f;
                if (b < B) {
                    a[j] = p.gates[((size_t)tt * p.BS + b) * 2 * H + col];             // fwd: P;  bwd: h
                    if (!FWD) g[j] = p.dy[((size_t)tt * p.BS + b) * 2 * H + col];
                }
            }
            if (reader) {
                ptx::mbar_wait(tfull, tphase);
                if (tid == 0) stamp(p, n, 3);
                ptx::tc_fence_after();
                uint32_t r[32], r2[32];
                float acc[NB];
                ptx::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16), r);      // rows = units 32*warp + lane of the block
                ptx::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + 32, r2);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[b] = __uint_as_float(r[b]) + __uint_as_float(r2[b]);
                if (NKB > 1) {                                                  // the odd k-blocks: the second issuer's accumulator
                    ptx::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + 64, r);
                    ptx::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + 96, r2);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int b = 0; b < NB; ++b) acc[b] += __uint_as_float(r[b]) + __uint_as_float(r2[b]);
                }
                ptx::tc_fence_before();
#pragma unroll
                for (int b = 0; b < NB; ++b)
                    ptx::st_cluster_f32(remote_slot + (uint32_t)(b * UPC + lane) * 4u, acc[b]);
                __syncwarp();
                if (lane == 0) { ptx::mbar_arrive(tempty); ptx::mbar_arrive_remote(remote_bar); }
            }
            ptx::mbar_wait_cluster(xfull, tphase);                          // the four partials of my units have landed
            if (tid == 0) stamp(p, n, 4);
            __nv_bfloat16 *xb = p.xbuf + ((size_t)(d * 2 + ((n + 1) & 1)) * NB) * H + unit;
            float out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = bg * 4 + j;
                const bool live = tt < len4[j];
                const int o = b * UPC + cu;
                const float s = (slots[o] + slots[NB * UPC + o]) + (slots[2 * NB * UPC + o] + slots[3 * NB * UPC + o]);
                float v = 0.f;
                if (live) {
                    if (FWD) {
                        const float z = a[j] + s;
                        v = tanh_cell ? rec::tanhf_(z) : fmaxf(z, 0.f);
                    } else {
                        const float dh = g[j] + s;
                        v = tanh_cell ? dh * (1.f - a[j] * a[j]) : (a[j] > 0.f ? dh : 0.f);
                    }
                }
                out[j] = v;
                if (!FWD) dbacc += v;
                __nv_bfloat16 hi, lo;                   // the bf16 pieces are what the other CTAs wait for
                rec::split2(v, hi, lo);
                xb[(size_t)b * H] = hi;
                xb[piece + (size_t)b * H] = lo;
            }
            // every writing thread orders its own pieces for the other CTAs' TMA (async proxy) reads (see lstm_tc.cu)
            if (tid == 0) stamp(p, n, 5);
            rec::fence_release_gpu();
            ptx::fence_proxy_async();
            cell_bar();
            if (tid == 0) { rec::signal_counter(p.counters + d); stamp(p, n, 6); }
#pragma unroll
            for (int j = 0; j < 4; ++j) {               // fp32 results for the other passes, after the signal
                const int b = bg * 4 + j;
                if (b < B) {
                    p.gates[((size_t)tt * p.BS + b) * 2 * H + col] = out[j];
                    if (FWD) p.y[((size_t)tt * p.BS + b) * 2 * H + col] = out[j];
                }
            }
        }
        if (!FWD && p.dbias) {
            // column sums of dz for my 32 units: the 8 row groups meet in the (now idle) exchange slots
            float *red = const_cast<float *>(slots);                        // [8 bg][32 cu]
            cell_bar();
            red[bg * UPC + cu] = dbacc;
            cell_bar();
            if (e < UPC) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) v += red[k * UPC + cu];
                float *dst = p.dbias + col;
                *dst = p.db_accum ? *dst + v : v;
            }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // no CTA exits while a peer may still write into its shared memory
    if (warp == 6) ptx::tmem_dealloc(tmem_d, 512);
}

// forward weights: Wh fp32 [2][H k][H u] -> [2 pieces][2H rows (d, u)][H k]: the K-major A operand of the swap-AB MMA
__global__ void pack_t_kernel(const float *__restrict__ wh, __nv_bfloat16 *__restrict__ wp, int H)
{
    __shared__ float tile[32][33];
    const int d = blockIdx.z, k0 = blockIdx.y * 32, u0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;                                // 32 x 8
    for (int j = ty; j < 32; j += 8) tile[j][tx] = wh[((size_t)d * H + k0 + j) * H + u0 + tx];
    __syncthreads();
    const size_t piece = (size_t)2 * H * H;
    for (int j = ty; j < 32; j += 8) {
        __nv_bfloat16 hi, lo;
        rec::split2(tile[tx][j], hi, lo);
        const size_t o = ((size_t)d * H + u0 + j) * H + k0 + tx;
        wp[o] = hi;
        wp[piece + o] = lo;
    }
}
// backward weights: plain 2-piece split of Wh viewed as [2H rows (d, k = unit of dh)][H columns of dz]
__global__ void split2_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        __nv_bfloat16 hi, lo;
        rec::split2(x[i], hi, lo);
        out[i] = hi;
        out[n + i] = lo;
    }
}

static unsigned long long *g_trace = nullptr;
struct WsLayout { size_t wpack, xbuf, counters, total; };
static WsLayout ws_layout(int H)
{
    WsLayout w;
    w.wpack = 0;
    w.xbuf = align_up((size_t)2 * 2 * H * H * 2, 1024);                     // 2 pieces x [2H][H] bf16
    w.counters = w.xbuf + align_up((size_t)2 * 2 * 2 * NB * H * 2, 1024);
    w.total = w.counters + 1024;
    return w;
}

template <bool FWD>
static int launch(const int *seq_len, const float *wh, float *gates, float *y, const float *dy, float *dbias,
                  int T, int B, int H, int cell, int use_len, void *ws, cudaStream_t stream)
{
    char *base = reinterpret_cast<char *>(align_up((size_t)(uintptr_t)ws, 1024));
    const WsLayout L = ws_layout(H);
    __nv_bfloat16 *wp = reinterpret_cast<__nv_bfloat16 *>(base + L.wpack);
    __nv_bfloat16 *xbuf = reinterpret_cast<__nv_bfloat16 *>(base + L.xbuf);
    unsigned int *ctr = reinterpret_cast<unsigned int *>(base + L.counters);
    const int NKB = H / 4 / BK, grid = 2 * (H / 128) * 4, CPD = grid / 2;
    const int smem = (int)smem_bytes(NKB);
    auto kernel = rnn_rec_kernel<FWD>;

    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    static int ok_grid[2] = {0, 0};
    if (ok_grid[FWD] != grid) {         // the step barrier spins: every cluster must be resident at once
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int nclusters = 0;
        CTCASR_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg));
        if (nclusters * 4 < grid) return fail(CTCASR_ERR_UNSUPPORTED, "rec_tc: %d clusters cannot be co-resident (%d)", grid / 4, nclusters);
        ok_grid[FWD] = grid;
    }
    if (FWD) pack_t_kernel<<<dim3(H / 32, H / 32, 2), dim3(32, 8), 0, stream>>>(wh, wp, H);
    else split2_kernel<<<148 * 4, 256, 0, stream>>>(wh, wp, (size_t)2 * H * H);
    CTCASR_LAUNCH_CHECK();
    CUtensorMap mapW, mapX;
    int rc = rec::make_map(&mapW, wp, (uint64_t)H, (uint64_t)2 * H, 128);
    if (rc) return rc;
    rc = rec::make_map(&mapX, xbuf, (uint64_t)H, (uint64_t)2 * 2 * NB, NB);
    if (rc) return rc;
    // batches above 32 rows run as consecutive launches over 32-row slices of the same buffers
    for (int b0 = 0; b0 < B; b0 += NB) {
        CTCASR_CUDA_CHECK(cudaMemsetAsync(xbuf, 0, (size_t)2 * 2 * 2 * NB * H * 2, stream));    // h_{-1} = 0 / no gradient into the last step
        CTCASR_CUDA_CHECK(cudaMemsetAsync(ctr, 0, 64, stream));
        Params p;
        p.T = T; p.B = B - b0 < NB ? B - b0 : NB; p.BS = B; p.H = H; p.CPD = CPD; p.use_len = use_len; p.cell = cell;
        p.seq_len = seq_len ? seq_len + b0 : nullptr;
        p.gates = gates + (size_t)b0 * 2 * H; p.y = y ? y + (size_t)b0 * 2 * H : nullptr; p.dy = dy ? dy + (size_t)b0 * 2 * H : nullptr;
        p.xbuf = xbuf; p.counters = ctr; p.wpack = wp; p.dbias = dbias; p.db_accum = b0 > 0;
        p.trace = FWD ? g_trace : nullptr;
        ProfScope prof(FWD ? PROF_LSTM_FWD : PROF_LSTM_BWD, stream);
        CTCASR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, mapW, mapX, p));
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
    }
    return CTCASR_OK;
}

}  // namespace rnn1

void rec_tc_set_trace(unsigned long long *buf) { rnn1::g_trace = buf; }

bool rec_tc_eligible(int T, int B, int H, int cell)
{
    return (cell == CTCASR_CELL_RNN_TANH || cell == CTCASR_CELL_RNN_RELU) && T >= 1 && B >= 1 &&
           H >= 256 && H % 256 == 0 && H / 256 <= rnn1::MAX_KB && H / 16 <= 148;
}

// state tiles of every step over all CTAs (the weights are loaded once per launch: 8 H^2 bytes)
double rec_tc_stream_bytes(int T, int H)
{
    const int NKB = H / 4 / rnn1::BK;
    return ((double)NKB * rnn1::B_TILE * T + (double)NKB * 2 * rnn1::A_PIECE) * (H / 16.0);
}

size_t rec_tc_workspace_bytes(int H)
{
    if (H < 256 || H % 256) return 0;
    return rnn1::ws_layout(H).total + 1024;
}

int rec_tc_fwd(const int *seq_len, const float *wh, float *gates, float *y, int T, int B, int H, int cell, int use_len,
               void *ws, cudaStream_t stream)
{
    return rnn1::launch<true>(seq_len, wh, gates, y, nullptr, nullptr, T, B, H, cell, use_len, ws, stream);
}

int rec_tc_bwd(const int *seq_len, const float *wh, float *gates, const float *dy, float *dbias,
               int T, int B, int H, int cell, int use_len, void *ws, cudaStream_t stream)
{
    return rnn1::launch<false>(seq_len, wh, gates, nullptr, dy, dbias, T, B, H, cell, use_len, ws, stream);
}

}  // namespace ctcasr
