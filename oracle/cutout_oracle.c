/*
 * oracle/cutout_oracle.c -- CPU restatement of the cosmo-cutout cutout hot path (see cutout_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Parity: pinned against oracle/_ref and tests/golden (header comment).
 *
 * Arithmetic contract: float products and sums in source order,
 * float sqrt / divide / acosf / atanf from the host libm (glibc 2.39 in this image), then the
 * arcsecond scaling ((double)t * 180.0) / 3.14159265 * 3600.0 in double, rounded to float.
 * Built with -ffp-contract=off and no -march so no FMA is formed, like the reference's -O3 build.
 */
#include "cutout_oracle.h"

#include <float.h>
#include <gnu/libc-version.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

_Static_assert(FLT_EVAL_METHOD == 0, "float expressions must evaluate in float");

/* acosf/atanf result -> arcsec, processLC.cpp:283-284: `t * 180.0 / PI * ARCSEC` assigned to float */
static inline float to_arcsec(float t) {
    return (float)((double)t * 180.0 / ORACLE_PI * ORACLE_ARCSEC);
}

/* processLC.cpp:282 (and :899, :1425): float d = (float)sqrt(x*x + y*y + z*z), all-float argument */
static inline float radius(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

/* processLC.cpp:903-908, :1430-1435 */
static inline float phi_guarded(float x, float y) {
    if (x == 0 && y > 0) return (float)(90.0 * ORACLE_ARCSEC);
    else if (x == 0 && y < 0) return (float)(-90.0 * ORACLE_ARCSEC);
    return to_arcsec(atanf(y / x));
}

void oracle_theta_phi_uc1(int64_t n, const float* x, const float* y, const float* z, float* theta, float* phi) {
    for (int64_t i = 0; i < n; ++i) {
        float d = radius(x[i], y[i], z[i]);
        theta[i] = to_arcsec(acosf(z[i] / d));
        phi[i] = to_arcsec(atanf(y[i] / x[i]));
    }
}

void oracle_theta_phi_uc2(int64_t n, const float* x, const float* y, const float* z, float* theta, float* phi) {
    for (int64_t i = 0; i < n; ++i) {
        float d = radius(x[i], y[i], z[i]);
        theta[i] = to_arcsec(acosf(z[i] / d));
        phi[i] = phi_guarded(x[i], y[i]);
    }
}

float oracle_aToZ(float a) { return (1.0f / a) - 1.0f; }

void oracle_matvec(const float* m, const float* v, float* out) {
    /* util.cpp:637-641: ans starts at 0.0f and accumulates m=0,1,2 in float */
    for (int r = 0; r < 3; ++r) {
        float acc = 0.0f;
        for (int c = 0; c < 3; ++c) acc += (m[3 * r + c] * v[c]);
        out[r] = acc;
    }
}

int oracle_zToStep(float z, int totSteps, float maxZ) {
    float amin = 1 / (maxZ + 1);
    float amax = 1.0;
    float adiff = (amax - amin) / (totSteps);
    float a = 1 / (1 + z);
    int step = (int)floorf((a - amin) / adiff);
    return step;
}

float oracle_stepToZ(float step, int totSteps, float maxZ) {
    float amin = 1 / (maxZ + 1);
    float amax = 1.0;
    float adiff = (amax - amin) / (totSteps);
    float a = amin + (adiff * step);
    float z = 1 / a - 1;
    return z;
}

void oracle_cli_bounds(float center_deg, float half_deg, float* out2) {
    /* main.cpp:265-268: strtof(...) * ARCSEC is a double product stored to float */
    float center = (float)((double)center_deg * ORACLE_ARCSEC);
    float half = (float)((double)half_deg * ORACLE_ARCSEC);
    out2[0] = center - half;
    out2[1] = center + half;
}

static void gather_one(const oracle_cols_in* in, int64_t n, float theta, float phi, int pos_only,
                       oracle_cols_out* out, int64_t k) {
    out->theta[k] = theta;
    out->phi[k] = phi;
    out->redshift[k] = oracle_aToZ(in->a[n]);
    out->x[k] = in->x[n];
    out->y[k] = in->y[n];
    out->z[k] = in->z[n];
    out->id[k] = in->id[n];
    if (!pos_only) {
        out->vx[k] = in->vx[n];
        out->vy[k] = in->vy[n];
        out->vz[k] = in->vz[n];
        out->rotation[k] = in->rotation[n];
        out->replication[k] = in->replication[n];
    }
}

int64_t oracle_cutout_const(const oracle_cols_in* in, const float* theta_cut, const float* phi_cut,
                            oracle_cols_out* out) {
    int64_t k = 0;
    for (int64_t n = 0; n < in->n; ++n) {
        const float x = in->x[n], y = in->y[n], z = in->z[n];
        if (x > 0.0 && y > 0.0 && z > 0.0) { /* :279 */
            float d = radius(x, y, z);
            float theta = to_arcsec(acosf(z / d));
            float phi = to_arcsec(atanf(y / x));
            if (theta > theta_cut[0] && theta < theta_cut[1] && phi > phi_cut[0] && phi < phi_cut[1]) { /* :287 */
                gather_one(in, n, theta, phi, 0, out, k);
                ++k;
            }
        }
    }
    return k;
}

/* ---------- per-halo set-up: processLC.cpp:512-733 ---------- */

/* std::inner_product(v,v,w,0.0): double accumulator, float products (util.cpp:659-661, :742) */
static double inner_d(const float* a, const float* b) {
    double acc = 0.0;
    for (int i = 0; i < 3; ++i) acc = acc + (double)(a[i] * b[i]);
    return acc;
}

static void sph_of_unit(const float* v, float* theta, float* phi) {
    /* :681-682: acos(v[2]/1) and atan(v[1]/v[0]), float libm, double scaling */
    *theta = to_arcsec(acosf(v[2] / 1));
    *phi = to_arcsec(atanf(v[1] / v[0]));
}

void oracle_halo_setup(const float* pos, float boxLength, oracle_halo* h) {
    memset(h, 0, sizeof(*h));
    for (int i = 0; i < 3; ++i) h->pos[i] = pos[i];

    h->halo_r = sqrtf(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]); /* :539 */

    float halfBoxLength = (float)(((boxLength / 2.0) / 60) * ORACLE_PI / 180.0); /* :545 */
    float dtheta = halfBoxLength, dphi = dtheta;
    h->half_box_rad = halfBoxLength;
    h->theta_cut[0] = (float)((ORACLE_PI / 2 - dtheta) * 180.0 / ORACLE_PI * ORACLE_ARCSEC); /* :550 */
    h->theta_cut[1] = (float)((ORACLE_PI / 2 + dtheta) * 180.0 / ORACLE_PI * ORACLE_ARCSEC);
    h->phi_cut[0] = (float)((0 - dphi) * 180.0 / ORACLE_PI * ORACLE_ARCSEC); /* :552 */
    h->phi_cut[1] = (float)((0 + dphi) * 180.0 / ORACLE_PI * ORACLE_ARCSEC);

    /* :570-581 rotation axis and angle taking pos onto (halo_r, 0, 0) */
    const float a[3] = {pos[0], pos[1], pos[2]};
    const float b[3] = {h->halo_r, 0, 0};
    float axb[3]; /* util.cpp:712-714 */
    axb[0] = a[1] * b[2] - a[2] * b[1];
    axb[1] = -(a[0] * b[2] - a[2] * b[0]);
    axb[2] = a[0] * b[1] - a[1] * b[0];
    float mag_axb = (float)sqrt(inner_d(axb, axb)); /* util.cpp:742 */
    for (int i = 0; i < 3; ++i) h->k[i] = (mag_axb == 0) ? 0 : axb[i] / mag_axb;

    float v1dv2 = (float)inner_d(a, b); /* util.cpp:659-663 */
    float mag_v1 = (float)sqrt(inner_d(a, a));
    float mag_v2 = (float)sqrt(inner_d(b, b));
    h->beta = acosf(v1dv2 / (mag_v1 * mag_v2));

    /* util.cpp:862-866 */
    const float* k = h->k;
    float K[9] = {0.0f, -k[2], k[1], k[2], 0.0f, -k[0], -k[1], k[0], 0.0f};
    memcpy(h->K, K, sizeof(K));

    /* util.cpp:888-897: R = I + sin(B) K + (1-cos(B)) K^2 */
    float K2[9], Ksin[9], K2cos[9];
    for (int n = 0; n < 3; ++n)
        for (int m = 0; m < 3; ++m) {
            float acc = 0.0f;
            for (int y = 0; y < 3; ++y) acc += K[3 * n + y] * K[3 * y + m];
            K2[3 * n + m] = acc;
        }
    const float s = sinf(h->beta);
    const float omc = 1 - cosf(h->beta);
    for (int i = 0; i < 9; ++i) { Ksin[i] = K[i] * s; K2cos[i] = K2[i] * omc; }
    for (int n = 0; n < 3; ++n)
        for (int m = 0; m < 3; ++m) {
            const int i = 3 * n + m;
            if (n == m) h->R[i] = (float)(1.0 + Ksin[i] + K2cos[i]); /* double sum, narrowed */
            else h->R[i] = Ksin[i] + K2cos[i];
        }

    /* util.cpp:778-839 inverse by scaled adjugate; determinant accumulates float terms in double */
    const float* m = h->R;
#define M(r, c) m[3 * (r) + (c)]
    double det;
    det = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1));
    det -= M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0));
    det += M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
    const double det_inv = 1.0 / det;
    const float sc = (float)det_inv; /* scale_adjoint_3x3 takes a float */
    float* ai = h->R_inv;
    ai[0] = sc * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1));
    ai[3] = sc * (M(1, 2) * M(2, 0) - M(1, 0) * M(2, 2));
    ai[6] = sc * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
    ai[1] = sc * (M(0, 2) * M(2, 1) - M(0, 1) * M(2, 2));
    ai[4] = sc * (M(0, 0) * M(2, 2) - M(0, 2) * M(2, 0));
    ai[7] = sc * (M(0, 1) * M(2, 0) - M(0, 0) * M(2, 1));
    ai[2] = sc * (M(0, 1) * M(1, 2) - M(0, 2) * M(1, 1));
    ai[5] = sc * (M(0, 2) * M(1, 0) - M(0, 0) * M(1, 2));
    ai[8] = sc * (M(0, 0) * M(1, 1) - M(0, 1) * M(1, 0));
#undef M

    /* :634-639 bounds back to radians (float) */
    float tr[2], pr[2];
    for (int i = 0; i < 2; ++i) {
        tr[i] = (float)((h->theta_cut[i] / ORACLE_ARCSEC) * ORACLE_PI / 180.0);
        pr[i] = (float)((h->phi_cut[i] / ORACLE_ARCSEC) * ORACLE_PI / 180.0);
    }
    /* :641-663 corners A(t1,p1) B(t0,p1) C(t0,p0) D(t1,p0) */
    const int ti[4] = {1, 0, 0, 1}, pi[4] = {1, 1, 0, 0};
    float th_c[4], ph_c[4];
    for (int c = 0; c < 4; ++c) {
        float v[3] = {sinf(tr[ti[c]]) * cosf(pr[pi[c]]), sinf(tr[ti[c]]) * sinf(pr[pi[c]]), cosf(tr[ti[c]])};
        oracle_matvec(h->R_inv, v, &h->corners_rot[3 * c]); /* :665-668 */
        sph_of_unit(&h->corners_rot[3 * c], &th_c[c], &ph_c[c]); /* :680-694 */
        h->corners_sph[2 * c] = th_c[c];
        h->corners_sph[2 * c + 1] = ph_c[c];
    }
    /* :714-724 (std::min_element / max_element: first minimum, first maximum) */
    const float ang_buffer = 600;
    float tmin = th_c[0], tmax = th_c[0], pmin = ph_c[0], pmax = ph_c[0];
    for (int c = 1; c < 4; ++c) {
        if (th_c[c] < tmin) tmin = th_c[c];
        if (tmax < th_c[c]) tmax = th_c[c];
        if (ph_c[c] < pmin) pmin = ph_c[c];
        if (pmax < ph_c[c]) pmax = ph_c[c];
    }
    h->theta_rough[0] = tmin - ang_buffer;
    h->theta_rough[1] = tmax + ang_buffer;
    h->phi_rough[0] = pmin - ang_buffer;
    h->phi_rough[1] = pmax + ang_buffer;
}

int64_t oracle_scatter_kept(int64_t Np, int numranks) {
    /* util.cpp:106-108: float np_per_rank = float(Np)/numranks; idxRemap = int(j / np_per_rank), j an int.
     * Kept = indices mapped to a rank < numranks (processLC.cpp:1004-1006 counts only those). The map is
     * monotone in j, so the kept set is a prefix; find its length by bisection on the same expression. */
    const float np_per_rank = (float)Np / numranks;
    int64_t lo = 0, hi = Np; /* invariant: all j < lo kept, all j >= hi dropped (or hi == Np) */
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        int r = (int)((int)mid / np_per_rank);
        if (r < numranks) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* stable merge sort of indices by key (std::stable_sort with `theta[n] < theta[m]`, processLC.cpp:1217) */
static void merge_sort_idx(int32_t* idx, int32_t* tmp, int64_t n, const float* key) {
    if (n < 2) return;
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t i = lo, j = mid, o = lo;
            while (i < mid && j < hi) {
                if (key[idx[j]] < key[idx[i]]) tmp[o++] = idx[j++]; else tmp[o++] = idx[i++];
            }
            while (i < mid) tmp[o++] = idx[i++];
            while (j < hi) tmp[o++] = idx[j++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int32_t));
    }
}

/* std::lower_bound on the sorted theta array (processLC.cpp:1334-1337) */
static int64_t lower_bound_sorted(const int32_t* idx, const float* key, int64_t n, float v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (key[idx[mid]] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

int64_t oracle_cutout_halo(const oracle_cols_in* in, const oracle_halo* h, int pos_only, int ref_numranks,
                           oracle_cols_out* out, int64_t* rough_count) {
    /* redistribution side effect (processLC.cpp:1003-1027): only a prefix of the read block survives */
    const int64_t Np = oracle_scatter_kept(in->n, ref_numranks);

    float* theta = (float*)malloc((size_t)(Np > 0 ? Np : 1) * sizeof(float));
    float* phi = (float*)malloc((size_t)(Np > 0 ? Np : 1) * sizeof(float));
    int32_t* idx = (int32_t*)malloc((size_t)(Np > 0 ? Np : 1) * sizeof(int32_t));
    int32_t* tmp = (int32_t*)malloc((size_t)(Np > 0 ? Np : 1) * sizeof(int32_t));
    oracle_theta_phi_uc2(Np, in->x, in->y, in->z, theta, phi); /* :897-909 */
    for (int64_t i = 0; i < Np; ++i) idx[i] = (int32_t)i;
    merge_sort_idx(idx, tmp, Np, theta); /* :1214-1221 */

    const int64_t minN = lower_bound_sorted(idx, theta, Np, h->theta_rough[0]); /* :1334-1340 */
    const int64_t maxN = lower_bound_sorted(idx, theta, Np, h->theta_rough[1]);

    int64_t k = 0, rough = 0;
    for (int64_t s = minN; s < maxN; ++s) { /* :1383 */
        const int64_t n = idx[s];
        const float p = phi[n];
        if (p > h->phi_rough[0] && p < h->phi_rough[1]) { /* :1400 */
            const float v[3] = {in->x[n], in->y[n], in->z[n]};
            float vr[3];
            oracle_matvec(h->R, v, vr); /* :1422 */
            float d = radius(vr[0], vr[1], vr[2]);
            float v_theta = to_arcsec(acosf(vr[2] / d));
            float v_phi = phi_guarded(vr[0], vr[1]);
            ++rough; /* :1436 */
            if (v_theta > h->theta_cut[0] && v_theta < h->theta_cut[1] && v_phi > h->phi_cut[0] &&
                v_phi < h->phi_cut[1]) { /* :1439 */
                gather_one(in, n, v_theta, v_phi, pos_only, out, k);
                ++k;
            }
        }
    }
    free(theta); free(phi); free(idx); free(tmp);
    if (rough_count) *rough_count = rough;
    return k;
}

void oracle_acosf_array(int64_t n, const float* in, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = acosf(in[i]);
}
void oracle_atanf_array(int64_t n, const float* in, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = atanf(in[i]);
}
const char* oracle_libc_version(void) { return gnu_get_libc_version(); }
