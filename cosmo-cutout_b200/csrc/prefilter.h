/*
 * prefilter.h -- conservative pre-filter bounds (host, double precision).
 *
 * The exact membership test runs the reference's float arithmetic (ref_math.cuh), ~150 instructions and two
 * fp64 divides per particle: far too much for every particle of a bandwidth-bound scan.  The scan kernels
 * therefore first apply a cheap test in cos(theta) / tan(phi) space and run the exact arithmetic only on what
 * passes.  The cheap test must never reject a particle the exact test accepts.  Error model:
 *
 *   exact path:  c_f  = fl(z / fl(sqrt(fl(fl(x*x + y*y) + z*z))))  = c (1+e1),   |e1| <= 4 ulp  ~ 2.4e-7
 *                th_f = fl32(fl64(acosf(c_f)) * 648000/PI_ref)     = acos(c_f) * K (1+e2), |e2| <= 2 ulp
 *                t_f  = fl(y/x) = t (1+e3), |e3| <= 0.5 ulp;   ph_f = atan(t_f) * K (1+e4), |e4| <= 2 ulp
 *   cheap path:  its own c~ (or c~^2) and t~ carry <= 8 ulp ~ 5e-7 relative error (fma, rsqrt.approx,
 *                div.approx) as long as every magnitude is inside [2^-30, 2^40] -- outside that range the
 *                kernels skip the cheap test and go straight to the exact path.
 *
 * Every bound is therefore widened by CC_PF_MARGIN = 1e-5 relative (20x the worst case above) BOTH in angle
 * and in cos/tan space, plus a small absolute slack, and rounded outward to float.  K = 648000/PI_ref with
 * PI_ref = 3.14159265 converts the reference's arcsecond floats back to the radians its acosf/atanf
 * returned, so radians here are true radians of those results.
 */
#pragma once
#include <math.h>

#define CC_PF_MARGIN 1e-5
#define CC_PF_ABS_RAD 1e-9

namespace ccpf {

static inline double arcsec_to_rad(double a) { return a * 3.14159265 / 180.0 / 3600.0; }
static inline float round_down(double v) {
    if (isinf(v)) return (float)v;
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return nextafterf(f, -INFINITY);
}
static inline float round_up(double v) {
    if (isinf(v)) return (float)v;
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return nextafterf(f, INFINITY);
}
/* values so small that a float product with them could go subnormal are pushed outward to 0 / +-1e-30 */
static inline double widen_small_lo(double v) { return (fabs(v) < 1e-30) ? (v < 0 ? -1e-30 : 0.0) : v; }
static inline double widen_small_hi(double v) { return (fabs(v) < 1e-30) ? (v > 0 ? 1e-30 : 0.0) : v; }

/* theta in (th_lo, th_hi) arcsec  =>  cos(theta) in [clo, chi] */
static inline void cos_bounds(double th_lo, double th_hi, double* clo, double* chi) {
    const double a_lo = arcsec_to_rad(th_lo), a_hi = arcsec_to_rad(th_hi);
    const double lo = a_lo - CC_PF_MARGIN * fabs(a_lo) - CC_PF_ABS_RAD;
    const double hi = a_hi + CC_PF_MARGIN * fabs(a_hi) + CC_PF_ABS_RAD;
    if (!(th_lo == th_lo) || !(th_hi == th_hi)) { *clo = INFINITY; *chi = -INFINITY; return; } /* NaN bound: nothing passes */
    if (lo <= 0.0) *chi = INFINITY;
    else if (lo >= M_PI) *chi = -1.0 + CC_PF_MARGIN; /* acos never exceeds pi (+ rounding) */
    else { const double c = cos(lo); *chi = widen_small_hi(c + CC_PF_MARGIN * fabs(c)); }
    if (hi >= M_PI) *clo = -INFINITY;
    else if (hi <= 0.0) *clo = 1.0 - CC_PF_MARGIN;
    else { const double c = cos(hi); *clo = widen_small_lo(c - CC_PF_MARGIN * fabs(c)); }
}

/* phi in (ph_lo, ph_hi) arcsec, phi = atan(y/x)  =>  y/x in [tlo, thi] */
static inline void tan_bounds(double ph_lo, double ph_hi, double* tlo, double* thi) {
    const double b_lo = arcsec_to_rad(ph_lo), b_hi = arcsec_to_rad(ph_hi);
    const double lo = b_lo - CC_PF_MARGIN * fabs(b_lo) - CC_PF_ABS_RAD;
    const double hi = b_hi + CC_PF_MARGIN * fabs(b_hi) + CC_PF_ABS_RAD;
    if (!(ph_lo == ph_lo) || !(ph_hi == ph_hi)) { *tlo = INFINITY; *thi = -INFINITY; return; }
    const double half_pi = M_PI / 2;
    if (lo <= -half_pi) *tlo = -INFINITY;
    else if (lo >= half_pi) *tlo = INFINITY;
    else { const double t = tan(lo); *tlo = widen_small_lo(t - CC_PF_MARGIN * fabs(t)); }
    if (hi >= half_pi) *thi = INFINITY;
    else if (hi <= -half_pi) *thi = -INFINITY;
    else { const double t = tan(hi); *thi = widen_small_hi(t + CC_PF_MARGIN * fabs(t)); }
}

/* use case 1 works on squares (z > 0 there): z^2 in (clo2, chi2) * r^2 */
static inline void uc1_bounds(const float th[2], const float ph[2], float* chi2, float* clo2, float* tlo, float* thi) {
    double clo, chi, tl, th_;
    cos_bounds(th[0], th[1], &clo, &chi);
    tan_bounds(ph[0], ph[1], &tl, &th_);
    if (chi <= 0.0) *chi2 = 0.0f;                 /* theta_lo beyond 90 deg: no first-octant particle passes */
    else if (isinf(chi)) *chi2 = INFINITY;
    else *chi2 = round_up(chi * chi * (1.0 + 1e-6));
    if (clo <= 0.0) *clo2 = -1.0f;                /* theta_hi beyond 90 deg: no lower test */
    else if (isinf(clo)) *clo2 = INFINITY;
    else *clo2 = round_down(clo * clo * (1.0 - 1e-6));
    *tlo = round_down(tl);
    *thi = round_up(th_);
}

static inline void uc2_bounds(const float th[2], const float ph[2], float out4[4]) {
    double clo, chi, tl, th_;
    cos_bounds(th[0], th[1], &clo, &chi);
    tan_bounds(ph[0], ph[1], &tl, &th_);
    out4[0] = round_up(chi);
    out4[1] = round_down(clo);
    out4[2] = round_down(tl);
    out4[3] = round_up(th_);
}

/* cell-index variant (k2_cutout_halo_cells): the azimuth is tested as s = sin(phi) = y*sign(x)/sqrt(x^2+y^2), which
 * is monotone in phi = atan(y/x) on (-90, 90) deg and bounded, so it can index a uniform grid.  The kernel's s~ and c~
 * (one rsqrt.approx + two multiplies each) are within 4e-7 ABSOLUTE of the true values (|s|,|c| <= 1), so after the
 * relative widening of the angles both intervals get an absolute 2e-6 on top. */
static inline void uc2_bounds_cells(const float th[2], const float ph[2], float out4[4]) {
    double clo, chi, tl, th_;
    cos_bounds(th[0], th[1], &clo, &chi);
    tan_bounds(ph[0], ph[1], &tl, &th_);
    const double abs_slack = 2e-6;
    double slo, shi;
    if (isinf(tl)) slo = tl; else slo = tl / sqrt(1.0 + tl * tl) - abs_slack;
    if (isinf(th_)) shi = th_; else shi = th_ / sqrt(1.0 + th_ * th_) + abs_slack;
    if (shi >= 1.0) shi = INFINITY;   /* sin can reach 1 (x == 0 is handled by the range flag, but stay safe) */
    if (slo <= -1.0) slo = -INFINITY;
    if (!isinf(chi)) chi += abs_slack;
    if (!isinf(clo)) clo -= abs_slack;
    out4[0] = round_up(chi);
    out4[1] = round_down(clo);
    out4[2] = round_down(slo);
    out4[3] = round_up(shi);
}

/* The same model read the other way: an INNER window in (c, s) space.  A pair whose cheap (c~, s~) lies inside it is
 * certain to pass the exact rough test (lo <= theta_u < hi, lo < phi_u < hi), so its exact angles need not be
 * evaluated to count it.  Every step of the outer derivation is monotone and every error term is two-sided: shrink
 * the angles by the margins instead of widening them, shrink again in cos / tan space, take off the absolute slack,
 * round inward.  Unbounded sides stay unbounded exactly where the outer window has them (theta_u <= 0 or >= pi,
 * |phi_u| >= pi/2 cannot occur for an in-range position).  A window narrower than its margins comes out empty
 * (lo > hi): nothing is certain, everything goes to the exact path. */
static inline void uc2_bounds_cells_inner(const float th[2], const float ph[2], float out4[4]) {
    out4[0] = -INFINITY; out4[1] = INFINITY; out4[2] = INFINITY; out4[3] = -INFINITY; /* empty */
    if (!(th[0] == th[0]) || !(th[1] == th[1]) || !(ph[0] == ph[0]) || !(ph[1] == ph[1])) return;
    const double abs_slack = 2e-6;
    const double a_lo = arcsec_to_rad(th[0]), a_hi = arcsec_to_rad(th[1]);
    const double b_lo = arcsec_to_rad(ph[0]), b_hi = arcsec_to_rad(ph[1]);
    const double lo = a_lo + CC_PF_MARGIN * fabs(a_lo) + CC_PF_ABS_RAD, hi = a_hi - CC_PF_MARGIN * fabs(a_hi) - CC_PF_ABS_RAD;
    const double pl = b_lo + CC_PF_MARGIN * fabs(b_lo) + CC_PF_ABS_RAD, pu = b_hi - CC_PF_MARGIN * fabs(b_hi) - CC_PF_ABS_RAD;
    const double half_pi = M_PI / 2;
    double chi, clo, slo, shi;
    if (a_lo < 0.0) chi = INFINITY;               /* theta_u >= 0 > lo always */
    else if (lo >= M_PI) return;
    else { const double c = cos(lo); chi = c - CC_PF_MARGIN * fabs(c) - abs_slack; }
    if (a_hi > M_PI + 1e-3) clo = -INFINITY;      /* theta_u <= pi (+ rounding) < hi always */
    else if (hi <= 0.0) return;
    else { const double c = cos(hi < M_PI ? hi : M_PI); clo = c + CC_PF_MARGIN * fabs(c) + abs_slack; }
    if (b_lo < -half_pi - 1e-3) slo = -INFINITY;  /* |phi_u| <= pi/2 (+ rounding) */
    else if (pl >= half_pi) return;
    else {
        const double t = tan(pl > -half_pi ? pl : -half_pi + 1e-12), tt = t + CC_PF_MARGIN * fabs(t);
        slo = tt / sqrt(1.0 + tt * tt) + abs_slack;
    }
    if (b_hi > half_pi + 1e-3) shi = INFINITY;
    else if (pu <= -half_pi) return;
    else {
        const double t = tan(pu < half_pi ? pu : half_pi - 1e-12), tt = t - CC_PF_MARGIN * fabs(t);
        shi = tt / sqrt(1.0 + tt * tt) - abs_slack;
    }
    out4[0] = isinf(chi) ? (float)chi : round_down(chi);
    out4[1] = isinf(clo) ? (float)clo : round_up(clo);
    out4[2] = isinf(slo) ? (float)slo : round_up(slo);
    out4[3] = isinf(shi) ? (float)shi : round_down(shi);
}

}  // namespace ccpf
