// tests/native/check_libm_restate.cpp -- TEST INFRASTRUCTURE.
// Compares the product's fdlibm restatement (cosmo-cutout_b200/csrc/ref_math.cuh, host build) with the
// host glibc acosf/atanf (what the reference links) on every one of the 2^32 float bit patterns, or on a
// strided subset (argv[1] = stride, default 1).  NaN results compare equal regardless of payload.
// Prints "acosf mismatches=<n> atanf mismatches=<n> checked=<n>" and exits non-zero on any mismatch.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ref_math.cuh"

static inline float W(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t B(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    uint64_t bad_acos = 0, bad_atan = 0, checked = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad_acos, bad_atan, checked)
    for (uint64_t u = 0; u < (1ull << 32); u += stride) {
        const float x = W((uint32_t)u);
        volatile float xv = x; /* keep the compiler from constant-folding libm calls */
        const float a0 = acosf(xv), a1 = ccref::acosf_ref(x);
        const float t0 = atanf(xv), t1 = ccref::atanf_ref(x);
        if (!(B(a0) == B(a1) || (a0 != a0 && a1 != a1))) ++bad_acos;
        if (!(B(t0) == B(t1) || (t0 != t0 && t1 != t1))) ++bad_atan;
        ++checked;
    }
    printf("acosf mismatches=%llu atanf mismatches=%llu checked=%llu\n", (unsigned long long)bad_acos,
           (unsigned long long)bad_atan, (unsigned long long)checked);
    return (bad_acos || bad_atan) ? 1 : 0;
}
