/*
 * oracle.c — TEST INFRASTRUCTURE ONLY.  See oracle_impl.h for scope, provenance and the
 * "parity unpinned" statement.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this file.
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

int oracle_operand_rounding = 0;
void oracle_set_operand_rounding(int mode) { oracle_operand_rounding = mode; }

#define REAL float
#define SUF(name) CAT(name, _f32)
#include "oracle_impl.h"
#undef REAL
#undef SUF
#undef NEG_INF
#undef RB

#define oracle_hash32 oracle_hash32_d
#define oracle_drop_keep oracle_drop_keep_d
#define REAL double
#define SUF(name) CAT(name, _f64)
#include "oracle_impl.h"
#undef REAL
#undef SUF

/* tf.edit_distance(decoded, labels, normalize) at asr/model.py:338: plain row-by-row Levenshtein. */
int oracle_edit_distance(const int *hyp, int hstride, const int *hyp_len, const int *truth, int tstride,
                         const int *truth_len, int B, int normalize, float *out)
{
    for (int b = 0; b < B; ++b) {
        const int n = hyp_len[b], m = truth_len[b];
        const int *h = hyp + (size_t)b * hstride, *t = truth + (size_t)b * tstride;
        int *prev = (int *)malloc(sizeof(int) * (m + 1)), *cur = (int *)malloc(sizeof(int) * (m + 1));
        for (int j = 0; j <= m; ++j) prev[j] = j;
        for (int i = 1; i <= n; ++i) {
            cur[0] = i;
            for (int j = 1; j <= m; ++j) {
                int v = prev[j - 1] + (h[i - 1] != t[j - 1]);
                if (prev[j] + 1 < v) v = prev[j] + 1;
                if (cur[j - 1] + 1 < v) v = cur[j - 1] + 1;
                cur[j] = v;
            }
            int *tmp = prev; prev = cur; cur = tmp;
        }
        const int dist = prev[m];
        out[b] = normalize ? (m > 0 ? (float)dist / (float)m : (n > 0 ? INFINITY : 0.f)) : (float)dist;
        free(prev); free(cur);
    }
    return 0;
}

#include "beam_search.h"

int oracle_abi_version(void) { return 1; }
