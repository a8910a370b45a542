/* CPU check of emfusion_b200/csrc/seq_add.h against the sequential loop (tests/test_seq_add.py). */
#include <stdio.h>
#include <stdlib.h>
static long g_skipped = 0;
#define EMF_SEQ_ADD_COUNT g_skipped
#include "seq_add.h"

static uint64_t s = 88172645463325252ull;
static uint32_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }

static float seq(float t, float step, int n) {
    for (int i = 0; i < n; ++i) { volatile float r = t + step; t = r; }
    return t;
}

int main(int argc, char** argv) {
    long cases = argc > 1 ? atol(argv[1]) : 200000;
    long bad = 0, total = 0;
    for (long c = 0; c < cases; ++c) {
        float t, step;
        int n = 1 + (int)(rnd() % 3000);
        const int kind = (int)(rnd() % 6);
        if (kind == 0) {            /* metric ray lengths and voxel steps */
            t = 0.01f + (float)(rnd() % 100000) * 1e-4f;
            step = 0.0005f + (float)(rnd() % 1000) * 1e-5f;
        } else if (kind == 1) {     /* ties: step = (q + 1/2) ulp(t) */
            const int e = 110 + (int)(rnd() % 30);
            t = emf_u2f(((uint32_t)e << 23) | (rnd() & 0x7fffff));
            const float u = emf_u2f((uint32_t)(e - 23) << 23);
            step = ((float)(rnd() % 4096) + 0.5f) * u;
        } else if (kind == 2) {     /* negative start crossing zero */
            t = -(float)(rnd() % 1000) * 1e-3f;
            step = 0.001f + (float)(rnd() % 100) * 1e-4f;
        } else if (kind == 3) {     /* random bit patterns, moderate exponents */
            t = emf_u2f(((uint32_t)(100 + rnd() % 50) << 23) | (rnd() & 0x7fffff));
            step = emf_u2f(((uint32_t)(90 + rnd() % 50) << 23) | (rnd() & 0x7fffff));
        } else if (kind == 4) {     /* step larger than t, tiny t */
            t = emf_u2f(((uint32_t)(1 + rnd() % 20) << 23) | (rnd() & 0x7fffff));
            step = emf_u2f(((uint32_t)(100 + rnd() % 30) << 23) | (rnd() & 0x7fffff));
        } else {                    /* exact binary fractions (no rounding at all) */
            t = (float)(rnd() % 4096) / 256.0f;
            step = (float)(1 + rnd() % 64) / 1024.0f;
        }
        const float a = seq(t, step, n), b = emf_seq_add(t, step, n);
        if (emf_f2u(a) != emf_f2u(b)) {
            if (bad < 10) printf("MISMATCH t=%a step=%a n=%d seq=%a closed=%a\n", t, step, n, a, b);
            ++bad;
        }
        total += n;
    }
    printf("cases %ld bad %ld steps %ld covered_by_closed_form %ld\n", cases, bad, total, g_skipped);
    return bad != 0 || g_skipped * 2 < total;
}
