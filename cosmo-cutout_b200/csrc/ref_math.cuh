/*
 * ref_math.cuh -- the reference's per-particle arithmetic, restated op for op for sm_100a.
 *
 * The reference computes (processLC.cpp:282-284, :899-908, :1425-1435)
 *     d     = (float)sqrt(x*x + y*y + z*z)            all-float products and sums, float sqrt
 *     theta = acos(z/d)  * 180.0 / PI * ARCSEC        float acosf from glibc, then double scaling
 *     phi   = atan(y/x)  * 180.0 / PI * ARCSEC        float atanf from glibc, then double scaling
 * with PI = 3.14159265 (processLC.h:35), ARCSEC = 3600.0 (processLC.h:36), and no FMA contraction
 * (x86-64 -O3 without -march).  glibc 2.39's acosf/atanf are the fdlibm float routines
 * (sysdeps/ieee754/flt-32/e_acosf.c, s_atanf.c); they are NOT correctly rounded, so the only way
 * to reproduce every theta/phi bit is to execute the same operation sequence.  Every arithmetic
 * step below therefore goes through an explicitly rounded, never-contracted primitive
 * (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn, __dmul_rn, __ddiv_rn on the device; plain IEEE
 * operators on the host, where this header is also compiled -- with -ffp-contract=off -- so the
 * restatement can be checked against the host libm on all 2^32 inputs without a GPU).
 *
 * Nothing here may be "simplified": association order, the double division by the truncated PI and
 * the final narrowing are all part of the result.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define CC_HD __host__ __device__ __forceinline__
#define CC_FMUL(a, b) __fmul_rn((a), (b))
#define CC_FADD(a, b) __fadd_rn((a), (b))
#define CC_FSUB(a, b) __fsub_rn((a), (b))
#define CC_FDIV(a, b) __fdiv_rn((a), (b))
#define CC_FSQRT(a) __fsqrt_rn((a))
#define CC_DMUL(a, b) __dmul_rn((a), (b))
#define CC_DDIV(a, b) __ddiv_rn((a), (b))
#define CC_D2F(a) __double2float_rn((a))
#define CC_F2BITS(a) __float_as_int((a))
#define CC_BITS2F(a) __int_as_float((a))
#else
#include <math.h>
#include <string.h>
#if defined(__CUDACC__)
#define CC_HD __host__ __device__ inline
#else
#define CC_HD static inline
#endif
#define CC_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define CC_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define CC_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#define CC_FDIV(a, b) ((float)((float)(a) / (float)(b)))
#define CC_FSQRT(a) sqrtf((a))
#define CC_DMUL(a, b) ((double)(a) * (double)(b))
#define CC_DDIV(a, b) ((double)(a) / (double)(b))
#define CC_D2F(a) ((float)(a))
static inline int32_t cc_f2bits_host(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float cc_bits2f_host(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
#define CC_F2BITS(a) cc_f2bits_host((a))
#define CC_BITS2F(a) cc_bits2f_host((a))
#endif

namespace ccref {

#define CC_W(u) CC_BITS2F((int32_t)(u))

/* glibc 2.39 __ieee754_acosf (flt-32/e_acosf.c); constants are the IEEE bit patterns of its .rodata */
CC_HD float acosf_ref(float x) {
    const float one = 1.0f;
    const float pi = CC_W(0x40490fdau), pio2_hi = CC_W(0x3fc90fdau), pio2_lo = CC_W(0x33a22168u);
    const float pS0 = CC_W(0x3e2aaaabu), pS1 = CC_W(0xbea6b090u), pS2 = CC_W(0x3e4e0aa8u),
                pS3 = CC_W(0xbd241146u), pS4 = CC_W(0x3a4f7f04u), pS5 = CC_W(0x3811ef08u);
    const float qS1 = CC_W(0xc019d139u), qS2 = CC_W(0x4001572du), qS3 = CC_W(0xbf303361u),
                qS4 = CC_W(0x3d9dc62eu);
    const int32_t hx = CC_F2BITS(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) { /* |x| == 1 */
        if (hx > 0) return 0.0f;
        return CC_FADD(pi, CC_FMUL(2.0f, pio2_lo));
    } else if (ix > 0x3f800000) { /* |x| > 1 or NaN */
        const float t = CC_FSUB(x, x);
        return CC_FDIV(t, t);
    }
    float z, s;
    if (ix < 0x3f000000) { /* |x| < 0.5 */
        if (ix <= 0x32800000) return CC_FADD(pio2_hi, pio2_lo);
        z = CC_FMUL(x, x);
    } else if (hx < 0) {
        z = CC_FMUL(CC_FADD(one, x), 0.5f);
    } else {
        z = CC_FMUL(CC_FSUB(one, x), 0.5f);
    }
    /* p = z*(pS0+z*(pS1+z*(pS2+z*(pS3+z*(pS4+z*pS5))))), q = 1+z*(qS1+z*(qS2+z*(qS3+z*qS4))) */
    float p = CC_FADD(pS4, CC_FMUL(z, pS5));
    p = CC_FADD(pS3, CC_FMUL(z, p));
    p = CC_FADD(pS2, CC_FMUL(z, p));
    p = CC_FADD(pS1, CC_FMUL(z, p));
    p = CC_FADD(pS0, CC_FMUL(z, p));
    p = CC_FMUL(z, p);
    float q = CC_FADD(qS3, CC_FMUL(z, qS4));
    q = CC_FADD(qS2, CC_FMUL(z, q));
    q = CC_FADD(qS1, CC_FMUL(z, q));
    q = CC_FADD(one, CC_FMUL(z, q));
    const float r = CC_FDIV(p, q);
    if (ix < 0x3f000000) {
        /* pio2_hi - (x - (pio2_lo - x*r)) */
        return CC_FSUB(pio2_hi, CC_FSUB(x, CC_FSUB(pio2_lo, CC_FMUL(x, r))));
    } else if (hx < 0) {
        s = CC_FSQRT(z);
        const float w = CC_FSUB(CC_FMUL(r, s), pio2_lo);
        return CC_FSUB(pi, CC_FMUL(2.0f, CC_FADD(s, w)));
    } else {
        s = CC_FSQRT(z);
        const float df = CC_BITS2F(CC_F2BITS(s) & (int32_t)0xfffff000);
        const float c = CC_FDIV(CC_FSUB(z, CC_FMUL(df, df)), CC_FADD(s, df));
        const float w = CC_FADD(CC_FMUL(r, s), c);
        return CC_FMUL(2.0f, CC_FADD(df, w));
    }
}

/* glibc 2.39 __atanf (flt-32/s_atanf.c) */
CC_HD float atanf_ref(float x) {
    const float one = 1.0f;
    const int32_t hx = CC_F2BITS(x);
    const int32_t ix = hx & 0x7fffffff;
    float hi, lo;
    int id;
    if (ix >= 0x4c000000) { /* |x| >= 2^25, inf, NaN */
        if (ix > 0x7f800000) return CC_FADD(x, x);
        const float a3 = CC_W(0x3fc90fdau), l3 = CC_W(0x33a22168u);
        if (hx > 0) return CC_FADD(a3, l3);
        return CC_FSUB(-a3, l3);
    }
    if (ix < 0x3ee00000) { /* |x| < 0.4375 */
        if (ix < 0x31000000) return x; /* |x| < 2^-29 */
        id = -1;
        hi = 0.0f;
        lo = 0.0f;
    } else {
        x = CC_BITS2F(ix); /* fabsf */
        if (ix < 0x3f980000) {     /* |x| < 1.1875 */
            if (ix < 0x3f300000) { /* 7/16 <= |x| < 11/16 */
                id = 0; hi = CC_W(0x3eed6338u); lo = CC_W(0x31ac3769u);
                x = CC_FDIV(CC_FSUB(CC_FMUL(2.0f, x), one), CC_FADD(2.0f, x));
            } else { /* 11/16 <= |x| < 19/16 */
                id = 1; hi = CC_W(0x3f490fdau); lo = CC_W(0x33222168u);
                x = CC_FDIV(CC_FSUB(x, one), CC_FADD(x, one));
            }
        } else {
            if (ix < 0x401c0000) { /* |x| < 2.4375 */
                id = 2; hi = CC_W(0x3f7b985eu); lo = CC_W(0x33140fb4u);
                x = CC_FDIV(CC_FSUB(x, 1.5f), CC_FADD(one, CC_FMUL(1.5f, x)));
            } else { /* 2.4375 <= |x| < 2^25 */
                id = 3; hi = CC_W(0x3fc90fdau); lo = CC_W(0x33a22168u);
                x = CC_FDIV(-1.0f, x);
            }
        }
    }
    const float aT0 = CC_W(0x3eaaaaabu), aT1 = CC_W(0xbe4ccccdu), aT2 = CC_W(0x3e124925u),
                aT3 = CC_W(0xbde38e38u), aT4 = CC_W(0x3dba2e6eu), aT5 = CC_W(0xbd9d8795u),
                aT6 = CC_W(0x3d886b35u), aT7 = CC_W(0xbd6ef16bu), aT8 = CC_W(0x3d4bda59u),
                aT9 = CC_W(0xbd15a221u), aT10 = CC_W(0x3c8569d7u);
    const float z = CC_FMUL(x, x);
    const float w = CC_FMUL(z, z);
    /* s1 = z*(aT0+w*(aT2+w*(aT4+w*(aT6+w*(aT8+w*aT10))))) ; s2 = w*(aT1+w*(aT3+w*(aT5+w*(aT7+w*aT9)))) */
    float s1 = CC_FADD(aT8, CC_FMUL(w, aT10));
    s1 = CC_FADD(aT6, CC_FMUL(w, s1));
    s1 = CC_FADD(aT4, CC_FMUL(w, s1));
    s1 = CC_FADD(aT2, CC_FMUL(w, s1));
    s1 = CC_FADD(aT0, CC_FMUL(w, s1));
    s1 = CC_FMUL(z, s1);
    float s2 = CC_FADD(aT7, CC_FMUL(w, aT9));
    s2 = CC_FADD(aT5, CC_FMUL(w, s2));
    s2 = CC_FADD(aT3, CC_FMUL(w, s2));
    s2 = CC_FADD(aT1, CC_FMUL(w, s2));
    s2 = CC_FMUL(w, s2);
    const float xs = CC_FMUL(x, CC_FADD(s1, s2));
    if (id < 0) return CC_FSUB(x, xs);
    const float r = CC_FSUB(hi, CC_FSUB(CC_FSUB(xs, lo), x));
    return (hx < 0) ? -r : r;
}

/* radians (float, from acosf/atanf) -> arcseconds, processLC.cpp:283-284:
 *   float = t * 180.0 / PI * ARCSEC   ==  (float)((((double)t * 180.0) / 3.14159265) * 3600.0)
 * This is the SPECIFICATION form: a true, correctly rounded fp64 division. */
CC_HD float to_arcsec(float t) {
    return CC_D2F(CC_DMUL(CC_DDIV(CC_DMUL((double)t, 180.0), 3.14159265), 3600.0));
}

#if defined(__CUDACC__)
/* Device fast form of to_arcsec: the division by the truncated PI is replaced by a multiplication with the correctly
 * rounded reciprocal and one fma correction step,
 *     q0 = a * RPI;   r = fma(-q0, PI, a);   q = fma(r, RPI, q0)          (a = t * 180.0 is exact in fp64)
 * which yields the SAME correctly rounded quotient as a / PI for every finite non-zero float t: verified on all 2^32
 * inputs on the CPU (tests/native/check_arcsec_fast.cpp, run by tests/test_host_cpu.py) and on the device against the
 * specification form (tests/test_gpu_parity.py).  Zero (sign of zero) and non-finite inputs take the specification
 * form.  Saves the ~25-instruction fp64 division sequence twice per candidate. */
__device__ __forceinline__ float to_arcsec_fast(float t) {
    const unsigned int mag = (unsigned int)__float_as_int(t) & 0x7fffffffu;
    if (mag - 1u >= 0x7f7fffffu) return to_arcsec(t); /* +-0, inf, NaN */
    const double PI = 3.14159265, RPI = 1.0 / 3.14159265;
    const double a = __dmul_rn((double)t, 180.0);
    const double q0 = __dmul_rn(a, RPI);
    const double r = __fma_rn(-q0, PI, a);
    const double q = __fma_rn(r, RPI, q0);
    return __double2float_rn(__dmul_rn(q, 3600.0));
}
#endif
#if defined(__CUDA_ARCH__)
#define CC_TO_ARCSEC(t) to_arcsec_fast((t))
#else
#define CC_TO_ARCSEC(t) to_arcsec((t))
#endif

/* processLC.cpp:282 / :899 / :1425 */
CC_HD float radius(float x, float y, float z) {
    return CC_FSQRT(CC_FADD(CC_FADD(CC_FMUL(x, x), CC_FMUL(y, y)), CC_FMUL(z, z)));
}

/* processLC.cpp:283 / :900 / :1426 */
CC_HD float theta_arcsec(float x, float y, float z) {
    return CC_TO_ARCSEC(acosf_ref(CC_FDIV(z, radius(x, y, z))));
}

/* processLC.cpp:284 (use case 1: no guards, first-octant particles only) */
CC_HD float phi_arcsec_plain(float x, float y) { return CC_TO_ARCSEC(atanf_ref(CC_FDIV(y, x))); }

/* processLC.cpp:903-908 / :1430-1435 (use case 2: x == 0 guards; 90.0*ARCSEC == 324000.0f exactly) */
CC_HD float phi_arcsec_guarded(float x, float y) {
    if (x == 0.0f && y > 0.0f) return 324000.0f;
    if (x == 0.0f && y < 0.0f) return -324000.0f;
    return CC_TO_ARCSEC(atanf_ref(CC_FDIV(y, x)));
}

/* util.cpp:620-643 matVecMul: ans[r] starts at 0.0f and adds R[r][m]*v[m] for m = 0,1,2 in float */
CC_HD void mat_vec(const float* R, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = CC_FADD(CC_FADD(CC_FADD(0.0f, CC_FMUL(R[0], x)), CC_FMUL(R[1], y)), CC_FMUL(R[2], z));
    oy = CC_FADD(CC_FADD(CC_FADD(0.0f, CC_FMUL(R[3], x)), CC_FMUL(R[4], y)), CC_FMUL(R[5], z));
    oz = CC_FADD(CC_FADD(CC_FADD(0.0f, CC_FMUL(R[6], x)), CC_FMUL(R[7], y)), CC_FMUL(R[8], z));
}

/* util.cpp:492-500 aToZ */
CC_HD float a_to_z(float a) { return CC_FSUB(CC_FDIV(1.0f, a), 1.0f); }

}  // namespace ccref
