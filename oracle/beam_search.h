/*
 * beam_search.h — TEST INFRASTRUCTURE ONLY (included by oracle.c).  "parity unpinned": see oracle_impl.h.
 *
 * Restates tf.nn.ctc_beam_search_decoder(inputs, sequence_length, beam_width, top_paths=1,
 * merge_repeated) as the reference calls it at asr/model.py:292-296 (beam_width = FLAGS.beam_width =
 * 1024, asr/params.py:85; merge_repeated=False).  The arithmetic lives in TensorFlow r1.12
 * (tensorflow/core/util/ctc/ctc_beam_search.h, ctc_beam_entry.h, kernels/ctc_decoder_ops.cc
 * [TF-RECALL]); this file follows that algorithm object for object:
 *   - a prefix tree of BeamEntry {parent, label, children, oldp, newp} with BeamProbability
 *     {total, blank, label} in the log domain, kLogZero = -inf; the root has total = blank = log 1;
 *   - `leaves`: a TopN of at most beam_width entries ordered by newp.total;
 *   - Step(frame): input -= max(input); every leaf's oldp = newp; every leaf is re-scored
 *       label: LSE(newp.label, parent active ? (label == parent.label ? parent.oldp.blank
 *                                                                     : parent.oldp.total) : -inf) + input[label]
 *       blank: oldp.total + input[blank];  total: LSE(blank, label)
 *     and pushed back; then, in descending oldp order, every leaf whose oldp is still a candidate
 *     offers its INACTIVE children: label = input[c] + (c == leaf.label ? oldp.blank : oldp.total),
 *     blank = -inf; a child enters iff its total is > -inf and (the beam is not full or it beats the
 *     bottom, which is then evicted and reset); rejected children are reset;
 *   - after the last frame the best leaf's label sequence is read off the parent chain
 *     (merge_repeated: drop a label equal to its predecessor).
 * Documented choices where TF leaves freedom: (1) r1.12's Step only subtracts the frame maximum
 * (later releases subtract the log-sum-exp: a per-frame constant common to every candidate, so the
 * decoded ids are the same and only the discarded log-probabilities shift); (2) exactly equal float
 * scores: TF's order is whatever its heap produces — here the earlier-inserted entry ranks higher;
 * (3) log(1 + exp(-d)) is evaluated with the fixed sequence of IEEE operations of `bs_softplus_neg`
 * (rint, fma, one division), repeated verbatim in ctc_asr_b200/csrc/beam.cu, so that CPU and GPU
 * scores agree to the last bit and the decoded ids can be compared exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float bs_softplus_neg(float d)     /* log(1 + exp(-d)), d >= 0 */
{
    if (!(d < 87.0f)) return 0.0f;
    const float x = -d;
    const float n = rintf(x * 1.44269504f);
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.42860677e-6f, r);
    float p = 2.48015873e-5f;                     /* 1/8! .. Horner to 1 */
    p = fmaf(p, r, 1.98412698e-4f);
    p = fmaf(p, r, 1.38888889e-3f);
    p = fmaf(p, r, 8.33333333e-3f);
    p = fmaf(p, r, 4.16666667e-2f);
    p = fmaf(p, r, 1.66666667e-1f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    union { uint32_t u; float f; } sc;
    sc.u = (uint32_t)((int)n + 127) << 23;        /* 2^n, n in [-126, 0] */
    const float e = p * sc.f;                     /* exp(-d) in (0, 1] */
    const float s = e / (2.0f + e);               /* log1p(e) = 2 atanh(e / (2 + e)) */
    const float s2 = s * s;
    float q = 7.69230769e-2f;                     /* 1/13 */
    q = fmaf(q, s2, 9.09090909e-2f);
    q = fmaf(q, s2, 1.11111111e-1f);
    q = fmaf(q, s2, 1.42857143e-1f);
    q = fmaf(q, s2, 2.0e-1f);
    q = fmaf(q, s2, 3.33333333e-1f);
    q = fmaf(q, s2, 1.0f);
    return (2.0f * s) * q;
}

static inline float bs_lse(float a, float b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const float m = a > b ? a : b, lo = a > b ? b : a;
    return m + bs_softplus_neg(m - lo);
}

typedef struct { float total, blank, label; } BsProb;
typedef struct { int parent, label, children, branch_step; uint64_t seq; BsProb oldp, newp; } BsEntry;

typedef struct {
    BsEntry *e; size_t n, cap;
    int *heap; int hn;             /* min-heap of entry indices: the bottom of the beam on top */
    uint64_t seq;
} BsState;

static void bs_reset_prob(BsProb *p) { p->total = p->blank = p->label = -INFINITY; }
/* a ranks below b */
static int bs_below(const BsState *s, int a, int b)
{
    const BsEntry *x = &s->e[a], *y = &s->e[b];
    if (x->newp.total != y->newp.total) return x->newp.total < y->newp.total;
    return x->seq > y->seq;
}
static void bs_heap_push(BsState *s, int idx)
{
    s->e[idx].seq = s->seq++;
    int i = s->hn++;
    s->heap[i] = idx;
    while (i > 0) {
        const int p = (i - 1) / 2;
        if (!bs_below(s, s->heap[i], s->heap[p])) break;
        const int t = s->heap[i]; s->heap[i] = s->heap[p]; s->heap[p] = t;
        i = p;
    }
}
static int bs_heap_pop(BsState *s)
{
    const int top = s->heap[0];
    s->heap[0] = s->heap[--s->hn];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < s->hn && bs_below(s, s->heap[l], s->heap[m])) m = l;
        if (r < s->hn && bs_below(s, s->heap[r], s->heap[m])) m = r;
        if (m == i) break;
        const int t = s->heap[i]; s->heap[i] = s->heap[m]; s->heap[m] = t;
        i = m;
    }
    return top;
}
static int bs_new_entries(BsState *s, int count)
{
    if (s->n + count > s->cap) {
        while (s->n + count > s->cap) s->cap *= 2;
        s->e = (BsEntry *)realloc(s->e, s->cap * sizeof(BsEntry));
    }
    const int first = (int)s->n;
    s->n += count;
    return first;
}

/* logits [T,B,V] time-major float; out_ids [B, T] (-1 padded), out_len [B], out_logp [B] (nullable) */
/* reoffer_wipe != 0: TF r1.12 as recalled, including an order-dependent artifact of its Step(): a leaf
 * that is evicted in the middle of the grow loop and then re-offered (and rejected) as a "new" child of
 * its parent has its oldp reset, so when the loop reaches that leaf it no longer offers its own children.
 * reoffer_wipe == 0: the same algorithm without that side effect on leaves of the current step; the beam
 * after every frame is then exactly the beam_width best of {re-scored leaves} U {absent children of leaves},
 * independent of visiting order — the formulation the CUDA kernel implements. */
int oracle_ctc_beam_search(const float *logits, int T, int B, int V, int blank, const int *seq_len,
                           int beam_width, int merge_repeated, int reoffer_wipe,
                           int *out_ids, int *out_len, float *out_logp)
{
    if (beam_width < 1 || V < 2 || blank != V - 1) return -1;     /* TF: blank = num_classes - 1 */
    float *input = (float *)malloc(sizeof(float) * V);
    int *branches = (int *)malloc(sizeof(int) * beam_width);
    for (int b = 0; b < B; ++b) {
        BsState s;
        s.cap = 1024; s.n = 0; s.e = (BsEntry *)malloc(s.cap * sizeof(BsEntry));
        s.heap = (int *)malloc(sizeof(int) * (beam_width + 1)); s.hn = 0; s.seq = 0;
        const int root = bs_new_entries(&s, 1);
        s.e[root].parent = -1; s.e[root].label = -1; s.e[root].children = -1; s.e[root].branch_step = -1;
        bs_reset_prob(&s.e[root].oldp); bs_reset_prob(&s.e[root].newp);
        s.e[root].newp.total = 0.f; s.e[root].newp.blank = 0.f;    /* log 1 */
        bs_heap_push(&s, root);
        const int Tb = seq_len[b] < T ? seq_len[b] : T;
        for (int t = 0; t < Tb; ++t) {
            const float *x = logits + ((size_t)t * B + b) * V;
            float mx = x[0];
            for (int k = 1; k < V; ++k) mx = x[k] > mx ? x[k] : mx;
            for (int k = 0; k < V; ++k) input[k] = x[k] - mx;
            /* branches = leaves.Extract(): descending newp.total */
            const int nb = s.hn;
            for (int i = nb - 1; i >= 0; --i) branches[i] = bs_heap_pop(&s);
            for (int i = 0; i < nb; ++i) { s.e[branches[i]].oldp = s.e[branches[i]].newp; s.e[branches[i]].branch_step = t; }
            for (int i = 0; i < nb; ++i) {
                BsEntry *e = &s.e[branches[i]];
                if (e->parent >= 0) {
                    const BsEntry *p = &s.e[e->parent];
                    if (p->newp.total != -INFINITY) {             /* parent->Active() */
                        const float previous = e->label == p->label ? p->oldp.blank : p->oldp.total;
                        e->newp.label = bs_lse(e->newp.label, previous);
                    }
                    e->newp.label += input[e->label];
                }
                e->newp.blank = e->oldp.total + input[blank];
                e->newp.total = bs_lse(e->newp.blank, e->newp.label);
                bs_heap_push(&s, branches[i]);
            }
            for (int i = 0; i < nb; ++i) {
                const int bi = branches[i];
                {
                    const BsProb *pr = &s.e[bi].oldp;
                    if (!(pr->total > -INFINITY && (s.hn < beam_width || pr->total > s.e[s.heap[0]].newp.total))) continue;
                }
                if (s.e[bi].children < 0) {                       /* PopulateChildren(num_classes - 1) */
                    const int first = bs_new_entries(&s, V - 1);
                    for (int c = 0; c < V - 1; ++c) {
                        BsEntry *ch = &s.e[first + c];
                        ch->parent = bi; ch->label = c; ch->children = -1; ch->seq = 0; ch->branch_step = -1;
                        bs_reset_prob(&ch->oldp); bs_reset_prob(&ch->newp);
                    }
                    s.e[bi].children = first;
                }
                for (int c = 0; c < V - 1; ++c) {
                    const int ci = s.e[bi].children + c;
                    BsEntry *ch = &s.e[ci];
                    const BsEntry *e = &s.e[bi];
                    if (ch->newp.total != -INFINITY) continue;    /* active: already in the beam */
                    ch->newp.blank = -INFINITY;
                    const float previous = c == e->label ? e->oldp.blank : e->oldp.total;
                    ch->newp.label = input[c] + previous;
                    ch->newp.total = ch->newp.label;
                    if (ch->newp.total > -INFINITY && (s.hn < beam_width || ch->newp.total > s.e[s.heap[0]].newp.total)) {
                        if (s.hn == beam_width) {
                            const int bottom = bs_heap_pop(&s);
                            bs_reset_prob(&s.e[bottom].newp);
                        }
                        bs_heap_push(&s, ci);
                    } else {
                        if (reoffer_wipe || ch->branch_step != t) bs_reset_prob(&ch->oldp);
                        bs_reset_prob(&ch->newp);
                    }
                }
            }
        }
        /* TopPaths(1): the best leaf */
        int best = s.heap[0];
        for (int i = 1; i < s.hn; ++i) if (bs_below(&s, best, s.heap[i])) best = s.heap[i];
        int n = 0;
        int *row = out_ids + (size_t)b * T;
        for (int c = best; s.e[c].parent >= 0; c = s.e[c].parent) row[n++] = s.e[c].label;   /* reversed */
        for (int i = 0; i < n / 2; ++i) { const int tmp = row[i]; row[i] = row[n - 1 - i]; row[n - 1 - i] = tmp; }
        if (merge_repeated) {
            int m = 0;
            for (int i = 0; i < n; ++i) if (i == 0 || row[i] != row[i - 1]) row[m++] = row[i];
            n = m;
        }
        for (int i = n; i < T; ++i) row[i] = -1;
        out_len[b] = n;
        if (out_logp) out_logp[b] = s.e[best].newp.total;
        free(s.e); free(s.heap);
    }
    free(input); free(branches);
    return 0;
}
