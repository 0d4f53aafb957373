/*
 * host_setup.cpp -- O(1)-per-halo host work of the cutout path (C ABI: cc_halo_setup, cc_cli_bounds,
 * cc_ref_scatter_kept).  It stays on the host, on the host libm, exactly like the reference
 * (processLC.cpp:512-733 with util.cpp:649-665, :695-751, :778-902): the bounds and the rotation matrix
 * must be the same bits the reference derives, and they depend on glibc's sinf/cosf/acosf/atanf.
 *
 * Compiled with -ffp-contract=off and no -march, like the reference (Makefile.cooley:2): no FMA is formed.
 * Every narrowing to float below mirrors an assignment to a `float` / `vector<float>` in the reference.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/cosmo_cutout_b200.h"

namespace ccint {
int fail_msg(int code, const char* msg); /* capi.cu: records the message cc_last_error() returns */
}

namespace {

const double kPi = 3.14159265; /* processLC.h:35 */
const double kArcsec = 3600.0; /* processLC.h:36 */

struct V3 { float v[3]; };
struct M3 { float m[3][3]; };

/* std::inner_product(a, a+3, b, 0.0): the accumulator is double, each product is a float product */
double dot_float_products(const V3& a, const V3& b) {
    double acc = 0.0;
    for (int i = 0; i < 3; ++i) {
        const float p = a.v[i] * b.v[i];
        acc = acc + p;
    }
    return acc;
}

/* util.cpp:620-643 */
V3 mat_vec(const M3& A, const V3& x) {
    V3 r = {{0.0f, 0.0f, 0.0f}};
    for (int n = 0; n < 3; ++n)
        for (int m = 0; m < 3; ++m) r.v[n] += (A.m[n][m] * x.v[m]);
    return r;
}

/* acosf/atanf result (radians, float) to arcsec float: processLC.cpp:681-682 `... * 180.0/PI * ARCSEC` */
float rad_to_arcsec(float t) { return (float)(t * 180.0 / kPi * kArcsec); }

}  // namespace

extern "C" {

int cc_halo_setup(const float pos[3], float boxLength, cc_halo_info* out) {
    if (!pos || !out) return CC_ERR_ARG;
    memset(out, 0, sizeof(*out));
    const V3 a = {{pos[0], pos[1], pos[2]}};
    for (int i = 0; i < 3; ++i) out->pos[i] = pos[i];

    /* processLC.cpp:539 -- float sum of float squares, float sqrt */
    const float halo_r = sqrtf(a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]);
    out->halo_r = halo_r;

    /* :545-553 -- box half-width in radians, then the equatorial bounds in arcsec */
    const float half = (float)(((boxLength / 2.0) / 60) * kPi / 180.0);
    out->half_box_rad = half;
    const float dtheta = half, dphi = half;
    cc_halo& H = out->halo;
    H.theta_cut[0] = (float)((kPi / 2 - dtheta) * 180.0 / kPi * kArcsec);
    H.theta_cut[1] = (float)((kPi / 2 + dtheta) * 180.0 / kPi * kArcsec);
    H.phi_cut[0] = (float)((0 - dphi) * 180.0 / kPi * kArcsec);
    H.phi_cut[1] = (float)((0 + dphi) * 180.0 / kPi * kArcsec);

    /* :570-581 -- axis k = (a x b)/|a x b| and angle beta between a and b = (halo_r, 0, 0) */
    const V3 b = {{halo_r, 0.0f, 0.0f}};
    V3 axb; /* util.cpp:712-714 */
    axb.v[0] = a.v[1] * b.v[2] - a.v[2] * b.v[1];
    axb.v[1] = -(a.v[0] * b.v[2] - a.v[2] * b.v[0]);
    axb.v[2] = a.v[0] * b.v[1] - a.v[1] * b.v[0];
    const float mag_axb = (float)sqrt(dot_float_products(axb, axb)); /* util.cpp:742 */
    V3 k;
    for (int i = 0; i < 3; ++i) k.v[i] = (mag_axb == 0) ? 0.0f : axb.v[i] / mag_axb; /* util.cpp:744-750 */
    const float adb = (float)dot_float_products(a, b);                               /* util.cpp:659 */
    const float mag_a = (float)sqrt(dot_float_products(a, a));
    const float mag_b = (float)sqrt(dot_float_products(b, b));
    const float beta = acosf(adb / (mag_a * mag_b)); /* util.cpp:663 */
    for (int i = 0; i < 3; ++i) out->k[i] = k.v[i];
    out->beta = beta;

    /* util.cpp:852-866 cross-product matrix */
    const M3 K = {{{0.0f, -k.v[2], k.v[1]}, {k.v[2], 0.0f, -k.v[0]}, {-k.v[1], k.v[0], 0.0f}}};
    /* util.cpp:877-902 Rodrigues: R = I + sin(beta) K + (1 - cos(beta)) K^2 */
    M3 K2;
    for (int n = 0; n < 3; ++n)
        for (int m = 0; m < 3; ++m) {
            K2.m[n][m] = 0.0f;
            for (int y = 0; y < 3; ++y) K2.m[n][m] += K.m[n][y] * K.m[y][m];
        }
    const float s = sinf(beta), omc = 1 - cosf(beta);
    M3 R;
    for (int n = 0; n < 3; ++n)
        for (int m = 0; m < 3; ++m) {
            const float ks = K.m[n][m] * s, kc = K2.m[n][m] * omc;
            R.m[n][m] = (n == m) ? (float)(1.0 + ks + kc) : (ks + kc); /* diagonal sums run in double */
        }

    /* util.cpp:778-839 inverse = adjugate scaled by (float)(1/det); det accumulates float terms in double */
    double det;
    det = R.m[0][0] * (R.m[1][1] * R.m[2][2] - R.m[1][2] * R.m[2][1]);
    det -= R.m[0][1] * (R.m[1][0] * R.m[2][2] - R.m[1][2] * R.m[2][0]);
    det += R.m[0][2] * (R.m[1][0] * R.m[2][1] - R.m[1][1] * R.m[2][0]);
    const float sc = (float)(1.0 / det);
    M3 Ri;
    Ri.m[0][0] = sc * (R.m[1][1] * R.m[2][2] - R.m[1][2] * R.m[2][1]);
    Ri.m[1][0] = sc * (R.m[1][2] * R.m[2][0] - R.m[1][0] * R.m[2][2]);
    Ri.m[2][0] = sc * (R.m[1][0] * R.m[2][1] - R.m[1][1] * R.m[2][0]);
    Ri.m[0][1] = sc * (R.m[0][2] * R.m[2][1] - R.m[0][1] * R.m[2][2]);
    Ri.m[1][1] = sc * (R.m[0][0] * R.m[2][2] - R.m[0][2] * R.m[2][0]);
    Ri.m[2][1] = sc * (R.m[0][1] * R.m[2][0] - R.m[0][0] * R.m[2][1]);
    Ri.m[0][2] = sc * (R.m[0][1] * R.m[1][2] - R.m[0][2] * R.m[1][1]);
    Ri.m[1][2] = sc * (R.m[0][2] * R.m[1][0] - R.m[0][0] * R.m[1][2]);
    Ri.m[2][2] = sc * (R.m[0][0] * R.m[1][1] - R.m[0][1] * R.m[1][0]);
    for (int n = 0; n < 3; ++n)
        for (int m = 0; m < 3; ++m) {
            out->K[3 * n + m] = K.m[n][m];
            out->R[3 * n + m] = R.m[n][m];
            out->R_inv[3 * n + m] = Ri.m[n][m];
            H.R[3 * n + m] = R.m[n][m];
        }

    /* :634-639 the fine bounds back in radians (float), :641-663 the four corner unit vectors
     * A(t1,p1) B(t0,p1) C(t0,p0) D(t1,p0), :665-668 pointed at the halo with R^-1, :680-694 their (theta,phi) */
    float t_rad[2], p_rad[2];
    for (int i = 0; i < 2; ++i) {
        t_rad[i] = (float)((H.theta_cut[i] / kArcsec) * kPi / 180.0);
        p_rad[i] = (float)((H.phi_cut[i] / kArcsec) * kPi / 180.0);
    }
    const int which_t[4] = {1, 0, 0, 1}, which_p[4] = {1, 1, 0, 0};
    float c_theta[4], c_phi[4];
    for (int c = 0; c < 4; ++c) {
        const float t = t_rad[which_t[c]], p = p_rad[which_p[c]];
        const V3 corner = {{sinf(t) * cosf(p), sinf(t) * sinf(p), cosf(t)}};
        const V3 rot = mat_vec(Ri, corner);
        c_theta[c] = rad_to_arcsec(acosf(rot.v[2] / 1));
        c_phi[c] = rad_to_arcsec(atanf(rot.v[1] / rot.v[0]));
        for (int i = 0; i < 3; ++i) out->corners_rot[3 * c + i] = rot.v[i];
        out->corners_sph[2 * c] = c_theta[c];
        out->corners_sph[2 * c + 1] = c_phi[c];
    }
    /* :714-724 min/max over the corners (std::min_element / max_element), widened by 600 arcsec */
    const float ang_buffer = 600;
    float t_min = c_theta[0], t_max = c_theta[0], p_min = c_phi[0], p_max = c_phi[0];
    for (int c = 1; c < 4; ++c) {
        if (c_theta[c] < t_min) t_min = c_theta[c];
        if (t_max < c_theta[c]) t_max = c_theta[c];
        if (c_phi[c] < p_min) p_min = c_phi[c];
        if (p_max < c_phi[c]) p_max = c_phi[c];
    }
    H.theta_rough[0] = t_min - ang_buffer;
    H.theta_rough[1] = t_max + ang_buffer;
    H.phi_rough[0] = p_min - ang_buffer;
    H.phi_rough[1] = p_max + ang_buffer;
    return CC_OK;
}

void cc_cli_bounds(float center_deg, float half_deg, float out[2]) {
    /* main.cpp:265-268: `float c = strtof(..) * ARCSEC` is a double product narrowed to float */
    const float center = (float)(center_deg * kArcsec);
    const float half = (float)(half_deg * kArcsec);
    out[0] = center - half;
    out[1] = center + half;
}

int64_t cc_ref_scatter_kept(int64_t np, int numranks) {
    /* util.cpp:106-108: idxRemap[j] = int(j / (float(Np)/numranks)) with j an int converted to float.
     * processLC.cpp:1003-1006 counts only entries < numranks, so entries that round up to numranks are never
     * sent.  float(j) is monotone in j, hence the kept set is a prefix: bisect for its length. */
    if (np <= 0 || numranks <= 0) return np < 0 ? 0 : np;
    if (np > 2147483647ll) return np; /* the reference counts particles in `int` (processLC.cpp:897): no such run exists */
    const float np_per_rank = (float)((size_t)np) / numranks;
    int64_t lo = 0, hi = np;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        const int r = (int)((int)mid / np_per_rank);
        if (r < numranks) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

void cc_shard_range(int64_t n, int nranks, int rank, int64_t* lo, int64_t* hi) {
    if (n < 0) n = 0;
    if (nranks < 1) nranks = 1;
    /* 128-bit product: n * rank can exceed 2^63 only for absurd sizes, but stay exact anyway */
    *lo = (int64_t)(((__int128)n * rank) / nranks);
    *hi = (int64_t)(((__int128)n * (rank + 1)) / nranks);
}

int cc_merge_order(int nranks, const int64_t* counts, const float* const* theta_u, const int64_t* const* index,
                   int32_t* src_rank, int64_t* src_row) {
    if (nranks < 1 || !counts || !theta_u || !index) return ccint::fail_msg(CC_ERR_ARG, "cc_merge_order: bad argument");
    int64_t total = 0;
    for (int r = 0; r < nranks; ++r) {
        if (counts[r] < 0 || (counts[r] > 0 && (!theta_u[r] || !index[r])))
            return ccint::fail_msg(CC_ERR_ARG, "cc_merge_order: a rank with survivors has no key columns");
        total += counts[r];
    }
    if (total == 0) return CC_OK; /* a halo without survivors in this step: nothing to order (outputs may be NULL) */
    if (!src_rank || !src_row) return ccint::fail_msg(CC_ERR_ARG, "cc_merge_order: output arrays are NULL");
    /* k-way merge; theta_u >= 0 and never NaN for a survivor, so its bit pattern orders like its value */
    struct Head { uint32_t key; int64_t idx; int32_t r; int64_t row; };
    std::vector<Head> heap;
    auto less = [](const Head& a, const Head& b) { /* min-heap through std::*_heap's max-heap convention */
        return a.key != b.key ? a.key > b.key : a.idx > b.idx;
    };
    auto head_of = [&](int r, int64_t row) {
        Head h;
        memcpy(&h.key, &theta_u[r][row], 4);
        h.idx = index[r][row];
        h.r = r;
        h.row = row;
        return h;
    };
    for (int r = 0; r < nranks; ++r)
        if (counts[r] > 0) heap.push_back(head_of(r, 0));
    std::make_heap(heap.begin(), heap.end(), less);
    int64_t o = 0;
    while (!heap.empty()) {
        std::pop_heap(heap.begin(), heap.end(), less);
        const Head h = heap.back();
        heap.pop_back();
        src_rank[o] = h.r;
        src_row[o] = h.row;
        ++o;
        if (h.row + 1 < counts[h.r]) {
            heap.push_back(head_of(h.r, h.row + 1));
            std::push_heap(heap.begin(), heap.end(), less);
        }
    }
    return CC_OK;
}

}  // extern "C"
