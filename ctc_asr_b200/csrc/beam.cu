// beam.cu — CTC prefix beam search for sm_100a: replaces tf.nn.ctc_beam_search_decoder(inputs,
// sequence_length, beam_width=FLAGS.beam_width (1024, asr/params.py:85), top_paths=1,
// merge_repeated=False) as decode_fn calls it (asr/model.py:292-296).
//
// One CTA (1024 threads) per utterance walks the frames; the beam (at most `beam_width` prefixes:
// node id, parent id, last label, log p_blank / p_label / p_total) lives in shared memory.  Per frame:
//   1. y = logits[t] - max (TF r1.12's Step() shifts by the frame maximum; the decoded ids do not
//      depend on the per-frame constant);
//   2. every prefix in the beam is re-scored:  label = LSE(label, parent in beam ? (same label as
//      parent ? parent.blank : parent.total) : -inf) + y[label],  blank = total + y[blank];
//      a shared-memory hash (node id -> slot) finds the parent and marks which children are present;
//   3. every absent child (prefix + c) is a candidate with label mass y[c] + (c == last ? blank : total);
//   4. the beam_width best of the  W + W (V-1)  candidates survive: exact radix select on the
//      order-preserving integer image of the scores (4 x 8 bits, warp-aggregated histogram), ties
//      resolved by candidate index (re-scored prefixes first), then a stable block-scan compaction;
//   5. surviving children get their node id from a per-utterance hash table in HBM keyed by
//      (parent id, label): a prefix that leaves the beam and later comes back is the same node, so
//      its still-active descendants find it again (TF keeps the tree object for the same reason).
// After the last frame the best prefix is read off the parent chain.
// Steps 3-4 define the beam order-independently: the beam_width best of {re-scored prefixes} U {absent children}.  TF
// r1.12's Step() visits the prefixes in descending order of their previous score and grows the beam sequentially; the
// result is the same set EXCEPT for one order-dependent side effect (oracle/beam_search.h, `reoffer_wipe`): a prefix X
// that is pushed out of the beam in the middle of the grow loop and then re-offered (and rejected) as a child of its
// parent P, when P's turn comes before X's own, no longer offers its own children.  Step 4b detects the frames in which
// that can change the beam (X does not survive, P ranks before X, X has a child that would survive) and replays TF's
// sequential loop for them, heap and all, in one thread; every other frame takes the parallel path.  Scores use the same
// fixed sequence of IEEE operations as the oracle (softplus_neg below), so transcripts can be compared bit for bit with
// the TF-faithful oracle mode.
#include "common.cuh"

#include <math.h>

namespace ctcasr {
namespace beam {

constexpr int NT = 1024;
constexpr int MAXW = 1024;
constexpr int MAXV = 32;
constexpr int SLOTS = 2048;             // shared-memory hash: node id -> beam slot

struct Params {
    const float *logits; int T, B, V, blank; const int *seq_len;
    int W, merge_repeated;
    int *out_ids; int *out_len; float *out_logp;
    int2 *nodes; size_t nodes_per_utt;          // {parent, label}
    int2 *table; size_t table_per_utt;          // {key = parent * 32 + label, node}; key -1 = empty
};

// log(1 + exp(-d)), d >= 0 — operation for operation oracle/beam_search.h:bs_softplus_neg
__device__ __forceinline__ float softplus_neg(float d)
{
    if (!(d < 87.0f)) return 0.0f;
    const float x = -d;
    const float n = rintf(__fmul_rn(x, 1.44269504f));
    float r = __fmaf_rn(n, -0.693145751953125f, x);
    r = __fmaf_rn(n, -1.42860677e-6f, r);
    float p = 2.48015873e-5f;
    p = __fmaf_rn(p, r, 1.98412698e-4f);
    p = __fmaf_rn(p, r, 1.38888889e-3f);
    p = __fmaf_rn(p, r, 8.33333333e-3f);
    p = __fmaf_rn(p, r, 4.16666667e-2f);
    p = __fmaf_rn(p, r, 1.66666667e-1f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    const float sc = __uint_as_float((uint32_t)((int)n + 127) << 23);
    const float e = __fmul_rn(p, sc);
    const float s = __fdiv_rn(e, __fadd_rn(2.0f, e));
    const float s2 = __fmul_rn(s, s);
    float q = 7.69230769e-2f;
    q = __fmaf_rn(q, s2, 9.09090909e-2f);
    q = __fmaf_rn(q, s2, 1.11111111e-1f);
    q = __fmaf_rn(q, s2, 1.42857143e-1f);
    q = __fmaf_rn(q, s2, 2.0e-1f);
    q = __fmaf_rn(q, s2, 3.33333333e-1f);
    q = __fmaf_rn(q, s2, 1.0f);
    return __fmul_rn(__fmul_rn(2.0f, s), q);
}
__device__ __forceinline__ float lse(float a, float b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const float m = a > b ? a : b, lo = a > b ? b : a;
    return __fadd_rn(m, softplus_neg(__fsub_rn(m, lo)));
}
// order-preserving image of a float; -inf -> 0x007fffff, every finite value is larger
__device__ __forceinline__ uint32_t okey(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
constexpr uint32_t KEY_NEG_INF = 0x007fffffu;

// exclusive scan of one int per thread over the CTA (NT = 1024); `red` holds 33 ints; returns the
// exclusive prefix and writes the total to *total
__device__ __forceinline__ int block_exscan(int v, int *red, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = red[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        red[lane] = w;                          // inclusive over warps
        if (lane == 31) red[32] = w;
    }
    __syncthreads();
    const int base = warp ? red[warp - 1] : 0;
    *total = red[32];
    __syncthreads();                            // red is reused by the next scan
    return base + inc - v;
}

struct Beam {                                   // one buffer of beam state in shared memory
    int *id, *par, *lab;
    float *pb, *pl, *pt;
};

__global__ void __launch_bounds__(NT, 1) beam_search_kernel(const Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int W = p.W, V = p.V, blank = p.blank, VC = V - 1;      // VC children per prefix (labels 0..V-2)
    const int NC = W * V;                                         // candidates: W re-scored + W * VC children
    // ---- shared memory carve-up ----
    int *ip = reinterpret_cast<int *>(smem);
    Beam bm[2];
    for (int k = 0; k < 2; ++k) {
        bm[k].id = ip; ip += MAXW; bm[k].par = ip; ip += MAXW; bm[k].lab = ip; ip += MAXW;
        bm[k].pb = reinterpret_cast<float *>(ip); ip += MAXW;
        bm[k].pl = reinterpret_cast<float *>(ip); ip += MAXW;
        bm[k].pt = reinterpret_cast<float *>(ip); ip += MAXW;
    }
    float *nbl = reinterpret_cast<float *>(ip); ip += MAXW;       // re-scored blank
    float *nlb = reinterpret_cast<float *>(ip); ip += MAXW;       // re-scored label
    uint32_t *present = reinterpret_cast<uint32_t *>(ip); ip += MAXW;
    int *psl = ip; ip += MAXW;                                    // slot of the parent prefix in the beam, or -1
    int *chead = ip; ip += MAXW;                                  // step 4b: list of the children of a prefix that are in the beam
    int *cnext = ip; ip += MAXW;
    unsigned char *gone = reinterpret_cast<unsigned char *>(ip); ip += MAXW / 4;      // step 4b: pushed out of the beam
    unsigned char *wiped = reinterpret_cast<unsigned char *>(ip); ip += MAXW / 4;     //          lost its right to grow
    uint32_t *keep = reinterpret_cast<uint32_t *>(ip); ip += (MAXW * MAXV) / 32;      //          members of the final heap, one bit per candidate
    int *hkey = ip; ip += SLOTS;
    int *hval = ip; ip += SLOTS;
    int *hist = ip; ip += 256;
    int *red = ip; ip += 40;
    float *y = reinterpret_cast<float *>(ip); ip += MAXV;
    int *misc = ip; ip += 8;                                      // [0] node counter, [1] prefix, [2] k remaining
    float *cand = reinterpret_cast<float *>(ip);                  // [NC]

    const int Tb = min(p.seq_len[b], p.T);
    int2 *nodes = p.nodes + (size_t)b * p.nodes_per_utt;
    int2 *table = p.table + (size_t)b * p.table_per_utt;
    const uint32_t tcap = (uint32_t)p.table_per_utt;

    int cur = 0, n = 1;
    if (tid == 0) {
        bm[0].id[0] = 0; bm[0].par[0] = -1; bm[0].lab[0] = -1;
        bm[0].pb[0] = 0.f; bm[0].pl[0] = -INFINITY; bm[0].pt[0] = 0.f;     // the empty prefix: log 1
        nodes[0] = make_int2(-1, -1);
        misc[0] = 1;
    }
    __syncthreads();

    for (int t = 0; t < Tb; ++t) {
        const Beam &B0 = bm[cur], &B1 = bm[cur ^ 1];
        // ---- 1. frame scores relative to the maximum ----
        if (tid < V) {
            const float *x = p.logits + ((size_t)t * p.B + b) * V;
            float m = x[0];
            for (int k = 1; k < V; ++k) m = fmaxf(m, x[k]);
            y[tid] = __fsub_rn(x[tid], m);
        }
        for (int i = tid; i < SLOTS; i += NT) hkey[i] = -1;
        if (tid < W) present[tid] = 0u;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        // ---- 2a. node id -> slot ----
        if (tid < n) {
            const int id = B0.id[tid];
            uint32_t h = ((uint32_t)id * 2654435761u) >> 21;              // 11 bits
            for (;;) {
                const int old = atomicCAS(&hkey[h], -1, id);
                if (old == -1) { hval[h] = tid; break; }
                h = (h + 1) & (SLOTS - 1);
            }
        }
        __syncthreads();
        // ---- 2b. re-score the prefixes in the beam ----
        float my_tot = -INFINITY;
        if (tid < n) {
            const int lab = B0.lab[tid], par = B0.par[tid];
            float prev = -INFINITY;
            int pslot = -1;
            if (par >= 0) {
                uint32_t h = ((uint32_t)par * 2654435761u) >> 21;
                for (;;) {
                    const int k = hkey[h];
                    if (k == par) {
                        const int ps = hval[h];
                        pslot = ps;
                        prev = lab == B0.lab[ps] ? B0.pb[ps] : B0.pt[ps];
                        atomicOr(&present[ps], 1u << lab);
                        break;
                    }
                    if (k == -1) break;
                    h = (h + 1) & (SLOTS - 1);
                }
            }
            const float nl = lab >= 0 ? __fadd_rn(lse(B0.pl[tid], prev), y[lab]) : -INFINITY;
            const float nb = __fadd_rn(B0.pt[tid], y[blank]);
            my_tot = lse(nb, nl);
            nbl[tid] = nb; nlb[tid] = nl;
            psl[tid] = pslot;
        }
        if (tid < W) cand[tid] = my_tot;
        __syncthreads();
        // ---- 3. absent children ----
        if (tid < W) {
            float *row = cand + W + tid * VC;
            if (tid < n) {
                const uint32_t pres = present[tid];
                const int lab = B0.lab[tid];
                const float pbv = B0.pb[tid], ptv = B0.pt[tid];
                for (int c = 0; c < VC; ++c)
                    row[c] = (pres >> c) & 1u ? -INFINITY : __fadd_rn(y[c], c == lab ? pbv : ptv);
            } else {
                for (int c = 0; c < VC; ++c) row[c] = -INFINITY;
            }
        }
        __syncthreads();
        // ---- 4. threshold = W-th largest finite score (exact radix select) ----
        int finite = 0;
        for (int i = tid; i < NC; i += NT) finite += okey(cand[i]) > KEY_NEG_INF;
        int M;
        block_exscan(finite, red, &M);
        uint32_t thr = KEY_NEG_INF + 1u;        // accept every finite candidate
        int k_eq = 0x7fffffff;                  // ... and every tie with the threshold
        if (M > W) {
            uint32_t prefix = 0;
            int k = W;
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 24 - 8 * pass;
                const uint32_t himask = pass ? ~((1u << (shift + 8)) - 1u) : 0u;
                for (int i0 = 0; i0 < NC; i0 += NT) {             // uniform trip count: match_any needs whole warps
                    const int i = i0 + tid;
                    int bin = 256;
                    if (i < NC) {
                        const uint32_t key = okey(cand[i]);
                        if (key > KEY_NEG_INF && (key & himask) == prefix) bin = (key >> shift) & 255;
                    }
                    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                    if (bin < 256 && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
                }
                __syncthreads();
                if (tid < 32) {                 // find the bin holding the k-th largest: lanes own 8 bins each, top down
                    int own = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) own += hist[255 - (lane * 8 + j)];
                    int inc = own;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += u;
                    }
                    const uint32_t hit = __ballot_sync(0xffffffffu, inc >= k);
                    const int owner = __ffs(hit) - 1;             // exists: the matching candidates number >= k
                    if (lane == owner) {
                        int above = inc - own;
                        for (int j = 0; j < 8; ++j) {
                            const int bn = 255 - (lane * 8 + j);
                            if (above + hist[bn] >= k) { misc[1] = (int)(prefix | ((uint32_t)bn << shift)); misc[2] = k - above; break; }
                            above += hist[bn];
                        }
                    }
                }
                __syncthreads();
                prefix = (uint32_t)misc[1]; k = misc[2];
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
            }
            thr = prefix; k_eq = k;
        }
        // ---- 4b. TF's order-dependent side effect (header comment): detect, and replay sequentially if it can matter ----
        {
            bool susp = false;
            if (tid < n && M > W) {
                const int ps = psl[tid];
                if (ps >= 0 && B0.pt[ps] >= B0.pt[tid] && okey(cand[tid]) <= thr) {
                    const float *row = cand + W + tid * VC;
                    float mx = -INFINITY;
                    for (int c = 0; c < VC; ++c) mx = fmaxf(mx, row[c]);
                    susp = mx > -INFINITY && okey(mx) >= thr;
                }
            }
            if (__syncthreads_or(susp)) {
                // the hash table is free now: heap of at most W entries {score, push order, candidate index} + rank -> slot
                float *hk = reinterpret_cast<float *>(hkey);
                int *hs = hkey + MAXW, *hr = hval, *byrank = hval + MAXW;
                if (tid < n) {      // TF's visiting order: descending previous score, earlier entry first on ties
                    const float me = B0.pt[tid];
                    int r = 0;
                    for (int j = 0; j < n; ++j) { const float o = B0.pt[j]; r += (o > me) || (o == me && j < tid); }
                    byrank[r] = tid;
                    chead[tid] = -1; gone[tid] = 0; wiped[tid] = 0;
                }
                for (int i = tid; i < (NC + 31) / 32; i += NT) keep[i] = 0u;
                __syncthreads();
                if (tid < n && psl[tid] >= 0) cnext[tid] = atomicExch(&chead[psl[tid]], tid);
                __syncthreads();
                if (tid < 32) {
                    // Warp 0 replays TF's grow loop.  The beam members live in an unsorted array (lane l owns entries
                    // l, l+32, ...); "bottom" = the minimum by (score, later push first), kept as a per-lane minimum +
                    // a shuffle reduction.  A child that fails against the current bottom fails against every later
                    // one (the bottom only rises), so the 28 children of a prefix are screened with one ballot and only
                    // the survivors are taken one by one, in class order.
                    int hn = n, seq = n;
                    for (int k = lane; k < n; k += 32) { const int sl = byrank[k]; hk[k] = cand[sl]; hs[k] = k; hr[k] = sl; }
                    __syncwarp();
                    float lmk; int lms, lmi;                        // my lowest entry
                    auto lower = [](float ka, int sa, float kb, int sb) { return ka < kb || (ka == kb && sa > sb); };
                    auto local_min = [&]() {
                        lmk = INFINITY; lms = -1; lmi = -1;
                        for (int k = lane; k < hn; k += 32)
                            if (lmi < 0 || lower(hk[k], hs[k], lmk, lms)) { lmk = hk[k]; lms = hs[k]; lmi = k; }
                    };
                    float bk; int bs, bi;                           // the bottom of the beam
                    auto bottom = [&]() {
                        bk = lmk; bs = lms; bi = lmi;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const float ok2 = __shfl_xor_sync(0xffffffffu, bk, o);
                            const int os = __shfl_xor_sync(0xffffffffu, bs, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
                            if (oi >= 0 && (bi < 0 || lower(ok2, os, bk, bs))) { bk = ok2; bs = os; bi = oi; }
                        }
                    };
                    local_min();
                    bottom();
                    for (int r = 0; r < n; ++r) {
                        const int i = byrank[r];
                        __syncwarp();
                        if (wiped[i]) continue;
                        const float pti = B0.pt[i];
                        if (!(pti > -INFINITY && (hn < W || pti > bk))) continue;
                        const int lab = B0.lab[i];
                        const uint32_t pres = present[i];
                        float sc = -INFINITY;
                        int sx = -1;
                        const int c = lane;
                        if (c < VC) {
                            if ((pres >> c) & 1u) {     // this child is a prefix of the beam: only of interest once it has been pushed out
                                sx = chead[i];
                                while (sx >= 0 && B0.lab[sx] != c) sx = cnext[sx];
                                if (sx >= 0 && gone[sx]) sc = __fadd_rn(y[c], c == lab ? B0.pb[i] : pti);
                                else sx = -1;
                            } else {
                                sc = cand[W + i * VC + c];
                            }
                        }
                        const bool pass = sc > -INFINITY && (hn < W || sc > bk);
                        if (sx >= 0 && !pass) wiped[sx] = 1;            // re-offered and rejected
                        uint32_t todo = __ballot_sync(0xffffffffu, pass);
                        while (todo) {
                            const int cc = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const float s2 = __shfl_sync(0xffffffffu, sc, cc);
                            const int sx2 = __shfl_sync(0xffffffffu, sx, cc);
                            const int idx = W + i * VC + cc;
                            if (hn < W || s2 > bk) {
                                int at = hn;
                                if (hn == W) {
                                    at = bi;
                                    if (lane == 0) {
                                        const int e = hr[bi];
                                        if (e < W) {
                                            gone[e] = 1;
                                            // a child of THIS prefix that is still ahead in class order is re-offered right away: rejected
                                            // (its fresh score cannot exceed the re-scored one that has just been the bottom)
                                            if (psl[e] == i && B0.lab[e] > cc) wiped[e] = 1;
                                        }
                                    }
                                    __syncwarp();           // the read of the evicted entry before its slot is overwritten
                                } else {
                                    ++hn;
                                }
                                if (lane == (at & 31)) { hk[at] = s2; hs[at] = seq; hr[at] = idx; if (sx2 >= 0) cand[idx] = s2; }
                                ++seq;
                                __syncwarp();
                                if (lane == (at & 31)) local_min();
                                bottom();
                            } else if (sx2 >= 0 && lane == 0) {
                                wiped[sx2] = 1;
                            }
                        }
                    }
                    __syncwarp();
                    for (int k = lane; k < hn; k += 32) atomicOr(&keep[hr[k] >> 5], 1u << (hr[k] & 31));
                }
                __syncthreads();
                for (int i = tid; i < NC; i += NT)
                    if (!((keep[i >> 5] >> (i & 31)) & 1u)) cand[i] = -INFINITY;
                thr = KEY_NEG_INF + 1u; k_eq = 0x7fffffff;                  // the survivors are exactly the finite candidates
                __syncthreads();
            }
        }
        // ---- stable compaction: a contiguous run of candidates per thread ----
        const int per = (NC + NT - 1) / NT;
        const int lo = tid * per, hi = min(NC, lo + per);
        int gt = 0, eq = 0;
        for (int i = lo; i < hi; ++i) {
            const uint32_t key = okey(cand[i]);
            gt += key > thr; eq += key == thr;
        }
        int tot_eq, n_new;
        const int eq_before = block_exscan(eq, red, &tot_eq);
        int eq_take = k_eq - eq_before;
        eq_take = eq_take < 0 ? 0 : (eq_take > eq ? eq : eq_take);
        int slot = block_exscan(gt + eq_take, red, &n_new);
        for (int i = lo; i < hi; ++i) {
            const float s = cand[i];
            const uint32_t key = okey(s);
            bool take = key > thr;
            if (key == thr && eq_take > 0) { take = true; --eq_take; }
            if (!take) continue;
            if (i < W) {
                B1.id[slot] = B0.id[i]; B1.par[slot] = B0.par[i]; B1.lab[slot] = B0.lab[i];
                B1.pb[slot] = nbl[i]; B1.pl[slot] = nlb[i]; B1.pt[slot] = s;
            } else {
                const int e = i - W, src = e / VC, c = e - src * VC;
                const int pid = B0.id[src];
                const int key2 = pid * 32 + c;
                uint32_t h = ((uint32_t)key2 * 2654435761u) % tcap;
                int node;
                for (;;) {
                    const int old = atomicCAS(&table[h].x, -1, key2);
                    if (old == -1) {                               // first time this prefix enters a beam
                        node = atomicAdd(&misc[0], 1);
                        table[h].y = node;
                        nodes[node] = make_int2(pid, c);
                        break;
                    }
                    if (old == key2) { node = table[h].y; break; } // written in an earlier frame
                    h = h + 1 == tcap ? 0 : h + 1;
                }
                B1.id[slot] = node; B1.par[slot] = pid; B1.lab[slot] = c;
                B1.pb[slot] = -INFINITY; B1.pl[slot] = s; B1.pt[slot] = s;
            }
            ++slot;
        }
        n = n_new;
        cur ^= 1;
        __syncthreads();
    }

    // ---- best prefix: largest total, lowest slot on ties ----
    const Beam &BF = bm[cur];
    float best = tid < n ? BF.pt[tid] : -INFINITY;
    int best_i = tid < n ? tid : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    float *redf = reinterpret_cast<float *>(hist);
    if (lane == 0) { redf[tid >> 5] = best; hist[64 + (tid >> 5)] = best_i; }
    __syncthreads();
    int *row = p.out_ids + (size_t)b * p.T;
    if (tid == 0) {
        for (int w = 1; w < NT / 32; ++w) {
            const float ov = redf[w]; const int oi = hist[64 + w];
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        int len = 0;
        for (int node = BF.id[best_i]; node > 0; node = nodes[node].x) ++len;
        int pos = len;
        for (int node = BF.id[best_i]; node > 0; node = nodes[node].x) row[--pos] = nodes[node].y;
        if (p.merge_repeated) {
            int m = 0;
            for (int i = 0; i < len; ++i) if (i == 0 || row[i] != row[i - 1]) row[m++] = row[i];
            len = m;
        }
        p.out_len[b] = len;
        if (p.out_logp) p.out_logp[b] = best;
        misc[3] = len;
    }
    __syncthreads();
    for (int i = misc[3] + tid; i < p.T; i += NT) row[i] = -1;
}

static size_t nodes_per_utt(int T, int W) { return (size_t)T * W + 1; }
static size_t table_per_utt(int T, int W) { return 2 * ((size_t)T * W + 1) + 1; }
static size_t smem_bytes(int W, int V)
{
    return sizeof(int) * ((size_t)12 * MAXW + 3 * MAXW + 3 * MAXW + MAXW / 2 + MAXW * MAXV / 32 + 2 * SLOTS + 256 + 40 + MAXV + 8) +
           sizeof(float) * (size_t)W * V;
}

}  // namespace beam
}  // namespace ctcasr

using namespace ctcasr;

extern "C" size_t ctcasr_beam_search_workspace_bytes(int T, int B, int V, int beam_width)
{
    if (T < 0 || B < 1 || V < 2 || V > beam::MAXV || beam_width < 1 || beam_width > beam::MAXW) return 0;
    const size_t per = (beam::nodes_per_utt(T, beam_width) + beam::table_per_utt(T, beam_width)) * sizeof(int2);
    return align_up(per, 256) * B;
}

extern "C" int ctcasr_beam_search(const float *logits, int T, int B, int V, int blank, const int32_t *seq_len,
                                  int beam_width, int merge_repeated, int32_t *out_ids, int32_t *out_len,
                                  float *out_logp, void *ws, size_t ws_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CTCASR_REQUIRE(logits && seq_len && out_ids && out_len && T >= 0 && B >= 1, "beam_search: bad args");
    CTCASR_REQUIRE(V >= 2 && V <= beam::MAXV && blank == V - 1, "beam_search: need 2 <= num_classes <= %d and blank = num_classes - 1 (got V=%d blank=%d)",
                   beam::MAXV, V, blank);
    CTCASR_REQUIRE(beam_width >= 1 && beam_width <= beam::MAXW, "beam_search: beam_width %d not in 1..%d", beam_width, beam::MAXW);
    CTCASR_REQUIRE((size_t)T * beam_width < ((size_t)1 << 26), "beam_search: T * beam_width too large");
    const size_t need = ctcasr_beam_search_workspace_bytes(T, B, V, beam_width);
    if (!ws || ws_bytes < need) return fail(CTCASR_ERR_WORKSPACE, "beam_search: workspace %zu < %zu", ws_bytes, need);
    beam::Params p;
    p.logits = logits; p.T = T; p.B = B; p.V = V; p.blank = blank; p.seq_len = seq_len;
    p.W = beam_width; p.merge_repeated = merge_repeated;
    p.out_ids = out_ids; p.out_len = out_len; p.out_logp = out_logp;
    p.nodes_per_utt = beam::nodes_per_utt(T, beam_width);
    p.table_per_utt = beam::table_per_utt(T, beam_width);
    // layout: all node arrays, then all tables (the tables are cleared to key = -1 with one memset)
    p.nodes = reinterpret_cast<int2 *>(ws);
    p.table = p.nodes + p.nodes_per_utt * B;
    CTCASR_CUDA_CHECK(cudaMemsetAsync(p.table, 0xff, p.table_per_utt * B * sizeof(int2), stream));
    const size_t smem = beam::smem_bytes(beam_width, V);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        CTCASR_CUDA_CHECK(cudaFuncSetAttribute(beam::beam_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    beam::beam_search_kernel<<<B, beam::NT, smem, stream>>>(p);
    CTCASR_LAUNCH_CHECK();
    return CTCASR_OK;
}
