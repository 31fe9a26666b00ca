/*
 * glc_detmath.h -- portable, bit-reproducible elementary functions (exp, log, pow, atan, cbrt).
 *
 * WHY: the reference's RHS nests loose-tolerance iterative solvers (Brent at 1e-2, a fixed-point
 * structure solve at 1e-2, adaptive quadrature at 1e-3).  Their discrete decisions (accept / stop /
 * bisect) flip on last-bit differences, so two implementations that differ only in the rounding of
 * exp/log/pow drift apart at the 1e-3 level on a fraction of nodes.  Built only from correctly rounded
 * IEEE-754 operations (+ - * / fma and bit manipulation), these functions return identical bits on the
 * GPU and on the host, which lets the parity tests assert bit-exact agreement between the CUDA path
 * and the CPU checker.  Fused multiply-adds are EXPLICIT (DM_FMA: fma() is correctly rounded on both
 * sides -- DFMA on the device, vfmadd on the host with -mfma, glibc's exact fma otherwise); the compilers'
 * own contraction stays off (-fmad=false / -ffp-contract=off) because each would fuse different pairs.
 * The polynomial kernels are where a GPU lane spends a quarter of its instructions (profiles/r02c):
 * fused, a Horner/Estrin step is one dependent instruction instead of two.
 *
 * Accuracy (checked against numpy/glibc in tests/test_detmath.py): exp, log, atan <= 2 ulp;
 * cbrt <= 1 ulp; pow(x,y) relative error <= ~(2 + |y ln x|) ulp.
 *
 * Algorithms: argument reduction + polynomial kernels in the style of Sun's fdlibm (public
 * domain algorithms; coefficients are the standard Taylor / minimax values).
 */
#ifndef GLC_DETMATH_H
#define GLC_DETMATH_H

#if defined(__CUDACC__)
#define GLC_HD __host__ __device__ __forceinline__
/* the big kernels call exp/log/pow/atan/cbrt as real functions: inlining their polynomial kernels at every
   call site blew the evolve kernel up to 665 KB of SASS and made it instruction-fetch bound */
#ifdef GLC_INLINE_MATH
#define GLC_HD_BIG __host__ __device__ __forceinline__
#else
#define GLC_HD_BIG __host__ __device__ __noinline__
#endif
#else
#define GLC_HD static inline
#define GLC_HD_BIG static inline
#endif

typedef union {
    double d;
    unsigned long long u;
} glc_dm_bits;

#if defined(__CUDACC__)
#define DM_FMA(a, b, c) fma((a), (b), (c))
#else
#define DM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

GLC_HD double dm_from_bits(unsigned long long u) {
    glc_dm_bits b;
    b.u = u;
    return b.d;
}
GLC_HD unsigned long long dm_to_bits(double d) {
    glc_dm_bits b;
    b.d = d;
    return b.u;
}

/* 2^k for integer k, exact (handles the subnormal range by two-step scaling) */
GLC_HD double dm_scale2(double x, int k) {
    if (k > 1023) {
        x = x * dm_from_bits(0x7fe0000000000000ULL); /* 2^1023 */
        k -= 1023;
        if (k > 1023) k = 1023;
    } else if (k < -1022) {
        x = x * dm_from_bits(0x0010000000000000ULL); /* 2^-1022 */
        k += 1022;
        if (k < -1022) k = -1022;
    }
    return x * dm_from_bits((unsigned long long)(k + 1023) << 52);
}

GLC_HD_BIG double dm_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double sqrt2 = 1.41421356237309514547;
    unsigned long long u;
    int e;
    double m, s, z, p;
    if (x != x) return x;
    if (x < 0.0) return dm_from_bits(0x7ff8000000000000ULL);
    if (x == 0.0) return -dm_from_bits(0x7ff0000000000000ULL);
    u = dm_to_bits(x);
    if (u == 0x7ff0000000000000ULL) return x;
    e = 0;
    if ((u >> 52) == 0) { /* subnormal */
        x = x * 18014398509481984.0; /* 2^54 */
        u = dm_to_bits(x);
        e = -54;
    }
    e += (int)((u >> 52) & 0x7ff) - 1023;
    m = dm_from_bits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > sqrt2) {
        m = m * 0.5;
        e += 1;
    }
    s = (m - 1.0) / (m + 1.0);
    z = s * s;
    /* 2 atanh(s) = 2 s (1 + z/3 + z^2/5 + ...): the series without its leading 1, p = z Q(z) with Q of degree 11,
       evaluated as four interleaved Horner chains in z^4 (Estrin) with fused steps: a dependency chain of 7 fused
       operations instead of 24 -- these functions run latency-bound on a GPU lane */
    {
        const double z2 = z * z, z4 = z2 * z2;
        const double q0 = DM_FMA(DM_FMA(1.0 / 19.0, z4, 1.0 / 11.0), z4, 1.0 / 3.0);
        const double q1 = DM_FMA(DM_FMA(1.0 / 21.0, z4, 1.0 / 13.0), z4, 1.0 / 5.0);
        const double q2 = DM_FMA(DM_FMA(1.0 / 23.0, z4, 1.0 / 15.0), z4, 1.0 / 7.0);
        const double q3 = DM_FMA(DM_FMA(1.0 / 25.0, z4, 1.0 / 17.0), z4, 1.0 / 9.0);
        p = z * DM_FMA(z2, DM_FMA(z, q3, q2), DM_FMA(z, q1, q0));
    }
    {
        const double two_s = 2.0 * s;
        const double de = (double)e;
        /* log = e ln2_hi + (2s + (2s*p + e ln2_lo)) */
        return DM_FMA(de, ln2_hi, two_s + DM_FMA(two_s, p, de * ln2_lo));
    }
}

GLC_HD_BIG double dm_exp(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double inv_ln2 = 1.44269504088896338700e+00;
    double r, p, fk;
    int k;
    if (x != x) return x;
    if (x > 709.782712893384) return dm_from_bits(0x7ff0000000000000ULL);
    if (x < -745.2) return 0.0;
    fk = x * inv_ln2;
    k = (int)(fk + (fk < 0.0 ? -0.5 : 0.5));
    fk = (double)k;
    r = DM_FMA(-fk, ln2_lo, DM_FMA(-fk, ln2_hi, x));
    /* Taylor series of e^r, |r| <= 0.3466, degree 14: e^r = 1 + r (1 + r (1/2 + r Q(r))), Q of degree 11 evaluated
       as four interleaved Horner chains in r^4 (Estrin; dependency chain 16 instead of 30), the three leading terms
       by Horner so that their rounding is as before */
    {
        const double r2 = r * r, r4 = r2 * r2;
        const double q0 = DM_FMA(DM_FMA(1.0 / 39916800.0, r4, 1.0 / 5040.0), r4, 1.0 / 6.0);
        const double q1 = DM_FMA(DM_FMA(1.0 / 479001600.0, r4, 1.0 / 40320.0), r4, 1.0 / 24.0);
        const double q2 = DM_FMA(DM_FMA(1.0 / 6227020800.0, r4, 1.0 / 362880.0), r4, 1.0 / 120.0);
        const double q3 = DM_FMA(DM_FMA(1.0 / 87178291200.0, r4, 1.0 / 3628800.0), r4, 1.0 / 720.0);
        p = DM_FMA(r2, DM_FMA(r, q3, q2), DM_FMA(r, q1, q0));
    }
    p = DM_FMA(p, r, 0.5);
    p = DM_FMA(p, r, 1.0);
    p = DM_FMA(p, r, 1.0);
    return dm_scale2(p, k);
}

GLC_HD_BIG double dm_pow(double x, double y) {
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (x != x || y != y) return x + y;
    if (x == 0.0) return (y > 0.0) ? 0.0 : dm_from_bits(0x7ff0000000000000ULL);
    if (x < 0.0) return dm_from_bits(0x7ff8000000000000ULL); /* not needed on this path */
    if (y == 1.0) return x;
    if (y == 2.0) return x * x;
    if (y == 0.5) return sqrt(x);
    return dm_exp(y * dm_log(x));
}

GLC_HD_BIG double dm_atan(double x) {
    const double atanhi0 = 4.63647609000806093515e-01, atanhi1 = 7.85398163397448278999e-01,
                 atanhi2 = 9.82793723247329054082e-01, atanhi3 = 1.57079632679489655800e+00;
    const double atanlo0 = 2.26987774529616870924e-17, atanlo1 = 3.06161699786838301793e-17,
                 atanlo2 = 1.39033110312309984516e-17, atanlo3 = 6.12323399573676603587e-17;
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01,
                 aT2 = 1.42857142725034663711e-01, aT3 = -1.11111104054623557880e-01,
                 aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02,
                 aT8 = 4.97687799461593236017e-02, aT9 = -3.65315727442169155270e-02,
                 aT10 = 1.62858201153657823623e-02;
    double ax, z, w, s1, s2, hi, lo, res;
    int id, neg;
    if (x != x) return x;
    neg = x < 0.0;
    ax = neg ? -x : x;
    if (ax >= 7.378697629483821e19) { /* 2^66 */
        res = atanhi3 + atanlo3;
        return neg ? -res : res;
    }
    if (ax < 0.4375) {
        if (ax < 1.862645149230957e-9) return x; /* 2^-29 */
        id = -1;
    } else if (ax < 1.1875) {
        if (ax < 0.6875) {
            id = 0;
            ax = (2.0 * ax - 1.0) / (2.0 + ax);
        } else {
            id = 1;
            ax = (ax - 1.0) / (ax + 1.0);
        }
    } else if (ax < 2.4375) {
        id = 2;
        ax = (ax - 1.5) / (1.0 + 1.5 * ax);
    } else {
        id = 3;
        ax = -1.0 / ax;
    }
    z = ax * ax;
    w = z * z;
    s1 = z * DM_FMA(w, DM_FMA(w, DM_FMA(w, DM_FMA(w, DM_FMA(w, aT10, aT8), aT6), aT4), aT2), aT0);
    s2 = w * DM_FMA(w, DM_FMA(w, DM_FMA(w, DM_FMA(w, aT9, aT7), aT5), aT3), aT1);
    if (id < 0) {
        res = DM_FMA(-ax, s1 + s2, ax);
        return neg ? -res : res;
    }
    hi = (id == 0) ? atanhi0 : (id == 1) ? atanhi1 : (id == 2) ? atanhi2 : atanhi3;
    lo = (id == 0) ? atanlo0 : (id == 1) ? atanlo1 : (id == 2) ? atanlo2 : atanlo3;
    res = hi - (DM_FMA(ax, s1 + s2, -lo) - ax);
    return neg ? -res : res;
}

GLC_HD_BIG double dm_cbrt(double x) {
    double a, t;
    int neg;
    if (x != x || x == 0.0) return x;
    neg = x < 0.0;
    a = neg ? -x : x;
    if (dm_to_bits(a) == 0x7ff0000000000000ULL) return x;
    t = dm_exp(dm_log(a) / 3.0);
    /* two Newton steps on t^3 = a */
    t = t - DM_FMA(t * t, t, -a) / (3.0 * t * t);
    t = t - DM_FMA(t * t, t, -a) / (3.0 * t * t);
    return neg ? -t : t;
}

#endif /* GLC_DETMATH_H */
