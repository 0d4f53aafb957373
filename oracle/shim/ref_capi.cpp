/*
 * oracle/shim/ref_capi.cpp -- C entry points onto the REAL reference, TEST INFRASTRUCTURE ONLY.
 *
 * Compiled by oracle/Makefile together with /root/reference/src/{processLC,util}.cpp (taken where
 * they lie; processLC.cpp gets the three one-line fixes of SURVEY.md 8c on the fly) into
 * oracle/_ref/libcutout_ref.so.  These wrappers only marshal C arrays into the std::vector
 * arguments of the reference's own functions (processLC.h:41-48, util.h:140-203); they add no
 * arithmetic of their own.
 */
#include "processLC.h"

#include <iostream>
#include <sstream>

namespace {
struct QuietCout {
    std::streambuf* old;
    std::ostringstream sink;
    bool on;
    explicit QuietCout(bool quiet) : old(0), on(quiet) { if (on) old = std::cout.rdbuf(sink.rdbuf()); }
    ~QuietCout() { if (on) std::cout.rdbuf(old); }
};
std::vector<std::string> to_strings(const char** s, int n) {
    std::vector<std::string> v;
    for (int i = 0; i < n; ++i) v.push_back(std::string(s[i]));
    return v;
}
std::vector<std::vector<float> > to_mat(const float* m) {
    std::vector<std::vector<float> > M(3, std::vector<float>(3));
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] = m[3 * r + c];
    return M;
}
void from_mat(const std::vector<std::vector<float> >& M, float* m) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[3 * r + c] = M[r][c];
}
}  // namespace

extern "C" {

/* use case 1: processLC.h:41-43.  returns 0, or 100+code when the reference called MPI_Abort. */
int ref_processLC_const(const char* dir_name, const char* out_dir, const char** steps, int nsteps,
                        float theta_lo, float theta_hi, float phi_lo, float phi_hi, int quiet) {
    QuietCout q(quiet != 0);
    shim_set_abort_throws(1);
    std::vector<float> th(2), ph(2);
    th[0] = theta_lo; th[1] = theta_hi; ph[0] = phi_lo; ph[1] = phi_hi;
    try {
        processLC(std::string(dir_name), std::string(out_dir), to_strings(steps, nsteps), th, ph, 0, 1,
                  false, false, false, false);
    } catch (shim_abort& e) {
        return 100 + e.code;
    }
    return 0;
}

/* use case 2: processLC.h:45-48 */
int ref_processLC_halo(const char* dir_name, const char** out_dirs, int nhalos, const char** steps, int nsteps,
                       const float* halo_pos, const float* halo_props, int nprops_total, float boxLength,
                       int verbose, int timeit, int overwrite, int positionOnly, int forceWriteProps,
                       int propsOnly, int quiet) {
    QuietCout q(quiet != 0);
    shim_set_abort_throws(1);
    std::vector<float> pos(halo_pos, halo_pos + 3 * nhalos);
    std::vector<float> props(halo_props, halo_props + nprops_total);
    try {
        processLC(std::string(dir_name), to_strings(out_dirs, nhalos), to_strings(steps, nsteps), pos, props,
                  boxLength, 0, 1, verbose != 0, timeit != 0, overwrite != 0, positionOnly != 0,
                  forceWriteProps != 0, propsOnly != 0);
    } catch (shim_abort& e) {
        return 100 + e.code;
    }
    return 0;
}

/* util.h:140-146 */
float ref_aToZ(float a) { return aToZ(a); }
int ref_zToStep(float z) { return zToStep(z); }
float ref_stepToZ(float step) { return stepToZ(step); }

/* util.h:155-203: rotation set-up exactly as processLC.cpp:570-592 chains it.
 * in: a[3] (halo position), b[3] (target).  out: k[3], beta, K[9], R[9], Rinv[9]. */
void ref_rotation_setup(const float* a, const float* b, float* k_out, float* beta_out, float* K_out,
                        float* R_out, float* Rinv_out) {
    std::vector<float> va(a, a + 3), vb(b, b + 3), k;
    normCross(va, vb, k);
    float beta = vecPairAngle(va, vb);
    std::vector<std::vector<float> > K, R;
    cross_prod_matrix(k, K);
    rotation_matrix(K, beta, R);
    std::vector<std::vector<float> > Rinv = invert_3x3(R);
    for (int i = 0; i < 3; ++i) k_out[i] = k[i];
    *beta_out = beta;
    from_mat(K, K_out);
    from_mat(R, R_out);
    from_mat(Rinv, Rinv_out);
}

/* util.cpp:620-643 */
void ref_matVecMul(const float* m, const float* v, float* out) {
    std::vector<float> vv(v, v + 3);
    std::vector<float> r = matVecMul(to_mat(m), vv);
    for (int i = 0; i < 3; ++i) out[i] = r[i];
}

/* util.cpp:97-110: number of indices j < Np that comp_rank_scatter maps to ranks < numranks,
 * i.e. the particles that survive the use-case-2 redistribution. */
long long ref_scatter_kept(long long Np, int numranks) {
    std::vector<int> remap;
    comp_rank_scatter((size_t)Np, remap, numranks);
    long long kept = 0;
    for (size_t j = 0; j < remap.size(); ++j) kept += (remap[j] < numranks);
    return kept;
}

int ref_getLCSteps(int maxStep, int minStep, const char* dir, char* out, int outcap) {
    std::vector<std::string> s;
    getLCSteps(maxStep, minStep, std::string(dir), s);
    std::string joined;
    for (size_t i = 0; i < s.size(); ++i) { if (i) joined += " "; joined += s[i]; }
    if ((int)joined.size() + 1 > outcap) return -1;
    strcpy(out, joined.c_str());
    return (int)s.size();
}

}  // extern "C"
