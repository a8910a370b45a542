/* seq_add.h -- n repeated fp32 additions of a constant in O(1).
 *
 * The reference raycast advances its ray parameter with `raylength += raystep` once per march sample
 * (reference src/core/cuda/TSDF.cu:523); every addition rounds, so the value after n samples is NOT
 * t + n*step.  When the samples themselves are provably no-ops (raycast.cu, "jumps") the n additions still
 * have to be reproduced bit for bit.  Inside one binade [2^e, 2^(e+1)) every float is a multiple of
 * u = 2^(e-23), so RN(t + step) = t + inc with a constant inc (a multiple of u, |step - inc| <= u/2) for
 * every t of the binade -- including the tie case once two consecutive additions agree (if inc/u were odd
 * the two would differ).  seq_add therefore takes true single steps until two consecutive increments
 * agree inside one binade, then covers the rest of that binade with one exact fma.
 *
 * Plain C so that the CPU test (tests/test_seq_add.py) can check it exhaustively against the sequential loop.
 */
#ifndef EMF_SEQ_ADD_H
#define EMF_SEQ_ADD_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define EMF_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define EMF_HD static inline
#endif

EMF_HD uint32_t emf_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
EMF_HD float emf_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
EMF_HD float emf_add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
EMF_HD float emf_fma_rn(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

/* t after n executions of `t = t + step` (round to nearest even each time).  Any t, step, n >= 0. */
EMF_HD float emf_seq_add(float t, float step, int n) {
    while (n > 0) {
        const float t1 = emf_add_rn(t, step);
        --n;
        if (n == 0) return t1;
        const uint32_t b0 = emf_f2u(t) >> 23, b1 = emf_f2u(t1) >> 23;     /* sign + exponent */
        /* closed form only for positive normal t in one binade with 0 < step <= t */
        if (b0 == b1 && b0 > 0 && b0 < 254 && step > 0.0f && step <= t) {
            const float inc = t1 - t;                                  /* exact (same binade) */
            const float t2 = emf_add_rn(t1, step);
            if ((emf_f2u(t2) >> 23) == b1 && t2 - t1 == inc && inc > 0.0f) {
                const float top = emf_u2f((b1 + 1) << 23);              /* 2^(e+1) */
                /* largest k with t1 + k*inc < top, minus a safety step for the rounding of the quotient */
                const float kq = (top - t1) / inc;
                int k = kq > 16777216.0f ? 16777216 : (int)kq;
                k -= 2;
                if (k > n) k = n;
                if (k > 0) {
#ifdef EMF_SEQ_ADD_COUNT
                    EMF_SEQ_ADD_COUNT += k;
#endif
                    t = emf_fma_rn((float)k, inc, t1);                   /* exact: a multiple of u below 2^(e+1) */
                    n -= k;
                    continue;
                }
            }
        }
        t = t1;
    }
    return t;
}

#endif /* EMF_SEQ_ADD_H */
