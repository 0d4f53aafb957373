// tests/native/check_window_filters.cpp -- TEST INFRASTRUCTURE.
// Property test of the conservative windows in cosmo-cutout_b200/csrc/prefilter.h against the exact arithmetic of
// cosmo-cutout_b200/csrc/ref_math.cuh (host build), for the three filters of the cell-index halo kernel:
//   outer (c, s) window  (uc2_bounds_cells)        exact rough test passes  =>  cheap (c~, s~) is inside
//   inner (c, s) window  (uc2_bounds_cells_inner)   cheap (c~, s~) is inside =>  exact rough test passes
//   outer (c, t) window  (uc2_bounds)               exact strict test passes =>  cheap (c~, t~) is inside
// The device computes c~ = z*rsqrt.approx(r2), s~ = y*rsqrt.approx(rxy2)*sign(x), t~ = y*rcp.approx(x) in float; their
// errors are bounded by 4e-7 absolute (c, s) and 5e-7 relative (t) (prefilter.h), so the properties are checked for the
// float evaluation AND for both ends of that error interval around the true (double) value.
// Points are aimed at every bound of every window (+- a few float ulps of the arcsecond value) and spread inside and
// outside.  Prints "windows=<n> points=<n> inner_decided=<n> violations=<n>"; exit 1 on any violation.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "prefilter.h"
#include "ref_math.cuh"

static uint64_t s_rng = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() { s_rng ^= s_rng << 13; s_rng ^= s_rng >> 7; s_rng ^= s_rng << 17; return s_rng; }
static inline double uni(double a, double b) { return a + (b - a) * (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }

int main() {
    const double conv = 3.14159265 / 180.0 / 3600.0;
    uint64_t windows = 0, points = 0, inner_decided = 0, bad = 0;
    for (int w = 0; w < 4000; ++w) {
        // rough-like windows anywhere on the sphere; every 4th one is a narrow fine-like window at theta ~ 90 deg, phi ~ 0
        const bool fine_like = (w % 4) == 3;
        const double tc = fine_like ? uni(89.5, 90.5) * 3600 : uni(0.3, 179.7) * 3600;
        const double pc = fine_like ? uni(-0.5, 0.5) * 3600 : uni(-89.0, 89.0) * 3600;
        const double tw = fine_like ? uni(30, 3000) : uni(100, 20000), pw = fine_like ? uni(30, 3000) : uni(100, 40000);
        const float th[2] = {(float)(tc - tw), (float)(tc + tw)}, ph[2] = {(float)(pc - pw), (float)(pc + pw)};
        float outer[4], inner[4], outer_t[4];
        ccpf::uc2_bounds_cells(th, ph, outer);
        ccpf::uc2_bounds_cells_inner(th, ph, inner);
        ccpf::uc2_bounds(th, ph, outer_t);
        ++windows;
        for (int k = 0; k < 600; ++k) {
            // angle pair: one angle on a bound (+- ulps) or both random around the window
            double t = uni(th[0] - 0.3 * tw, th[1] + 0.3 * tw), p = uni(ph[0] - 0.3 * pw, ph[1] + 0.3 * pw);
            const int aim = k % 6;
            if (aim < 4) {
                const float b = aim == 0 ? th[0] : aim == 1 ? th[1] : aim == 2 ? ph[0] : ph[1];
                const double jit = (double)((int)(rnd() % 9) - 4) * (double)(nextafterf(fabsf(b) + 1.0f, INFINITY) - (fabsf(b) + 1.0f));
                if (aim < 2) t = (double)b + jit; else p = (double)b + jit;
            }
            if (t <= 1.0 || t >= 647999.0 || fabs(p) >= 323999.0) continue;
            const double r = exp(uni(log(1.0), log(3000.0)));
            const double sgn = (rnd() & 1) ? 1.0 : -1.0; /* phi = atan(y/x): x < 0 mirrors (x, y) */
            const float x = (float)(sgn * r * sin(t * conv) * cos(p * conv)), y = (float)(sgn * r * sin(t * conv) * sin(p * conv)),
                        z = (float)(r * cos(t * conv));
            if (fabsf(x) < 1e-6f) continue; /* the kernels send |x| < 2^-30 to the exact path */
            ++points;
            const float thu = ccref::theta_arcsec(x, y, z), phu = ccref::phi_arcsec_guarded(x, y);
            const bool rough = thu >= th[0] && thu < th[1] && phu > ph[0] && phu < ph[1];
            const bool strict = thu > th[0] && thu < th[1] && phu > ph[0] && phu < ph[1];
            // true values of the float position, and the float evaluation the device resembles
            const double rxy2 = (double)x * x + (double)y * y, r2 = rxy2 + (double)z * z;
            const double c = z / sqrt(r2), s = (x < 0 ? -1.0 : 1.0) * y / sqrt(rxy2), tt = (double)y / (double)x;
            const float rxy2f = fmaf(y, y, x * x), r2f = fmaf(z, z, rxy2f);
            const float cf = z * (1.0f / sqrtf(r2f)), sf = (x < 0 ? -1.0f : 1.0f) * (y * (1.0f / sqrtf(rxy2f))), tf = y * (1.0f / x);
            const double cs[3] = {(double)cf, c - 4e-7, c + 4e-7}, ss[3] = {(double)sf, s - 4e-7, s + 4e-7};
            const double ts[3] = {(double)tf, tt - 5e-7 * fabs(tt), tt + 5e-7 * fabs(tt)};
            bool any_inner = false;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    const bool in_outer = cs[i] <= outer[0] && cs[i] >= outer[1] && ss[j] >= outer[2] && ss[j] <= outer[3];
                    const bool in_inner = cs[i] <= inner[0] && cs[i] >= inner[1] && ss[j] >= inner[2] && ss[j] <= inner[3];
                    const bool in_outer_t = cs[i] <= outer_t[0] && cs[i] >= outer_t[1] && ts[j] >= outer_t[2] && ts[j] <= outer_t[3];
                    if (rough && !in_outer) ++bad;
                    if (in_inner && !rough) ++bad;
                    if (strict && !in_outer_t) ++bad;
                    any_inner |= in_inner;
                }
            inner_decided += any_inner;
        }
    }
    printf("windows=%llu points=%llu inner_decided=%llu violations=%llu\n", (unsigned long long)windows,
           (unsigned long long)points, (unsigned long long)inner_decided, (unsigned long long)bad);
    return bad ? 1 : 0;
}
