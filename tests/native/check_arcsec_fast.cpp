// tests/native/check_arcsec_fast.cpp -- TEST INFRASTRUCTURE.
// Proof by exhaustion for ccref::to_arcsec_fast (cosmo-cutout_b200/csrc/ref_math.cuh): for every finite non-zero
// float t, the division-free form
//     a = (double)t * 180.0;  q0 = a * RPI;  r = fma(-q0, PI, a);  q = fma(r, RPI, q0);  (float)(q * 3600.0)
// equals the reference's (float)(((double)t * 180.0) / 3.14159265 * 3600.0) bit for bit, and even the fp64 quotient
// q equals a / PI.  Prints "checked=<n> quotient_mismatches=<n> float_mismatches=<n>"; exit 1 on any mismatch.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main() {
    const double PI = 3.14159265;
    const double RPI = 1.0 / 3.14159265;
    uint64_t bad_q = 0, bad_f = 0, n = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad_q, bad_f, n)
    for (uint64_t u = 0; u < (1ull << 32); ++u) {
        const uint32_t b = (uint32_t)u;
        if (((b & 0x7fffffffu) - 1u) >= 0x7f7fffffu) continue; /* +-0, inf, NaN take the specification form */
        float t;
        memcpy(&t, &b, 4);
        volatile double a = (double)t * 180.0;
        const double q = a / PI;
        const double q0 = a * RPI;
        const double r = __builtin_fma(-q0, PI, a);
        const double q1 = __builtin_fma(r, RPI, q0);
        if (memcmp(&q, &q1, 8) != 0) ++bad_q;
        const float f0 = (float)(q * 3600.0), f1 = (float)(q1 * 3600.0);
        if (memcmp(&f0, &f1, 4) != 0) ++bad_f;
        ++n;
    }
    printf("checked=%llu quotient_mismatches=%llu float_mismatches=%llu\n", (unsigned long long)n,
           (unsigned long long)bad_q, (unsigned long long)bad_f);
    return (bad_q || bad_f) ? 1 : 0;
}
