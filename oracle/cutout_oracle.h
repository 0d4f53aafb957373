/*
 * oracle/cutout_oracle.h -- CPU restatement of the cosmo-cutout lightcone cutout hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; nothing under cosmo-cutout_b200/ links or calls it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so the
 * pin is the reference itself: oracle/_ref (the real sources compiled by oracle/Makefile) reproduces
 * the three known answers of SURVEY.md App. B.4, and tests/test_oracle_vs_ref.py checks this
 * restatement byte-for-byte against oracle/_ref; tests/golden/ holds vectors generated from
 * oracle/_ref (script committed beside them) for boxes where /root/reference is absent.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference/src).
 */
#ifndef CUTOUT_ORACLE_H
#define CUTOUT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_PI 3.14159265 /* processLC.h:35 (truncated literal, used in double) */
#define ORACLE_ARCSEC 3600.0 /* processLC.h:36 */

/* input columns of one lightcone step (util.h:33-47); vx..replication may be NULL when pos_only */
typedef struct {
    int64_t n;
    const float *x, *y, *z, *vx, *vy, *vz, *a;
    const int64_t* id;
    const int32_t *rotation, *replication;
} oracle_cols_in;

/* output columns (util.h:50-64); each must have room for the worst case (n) */
typedef struct {
    float *x, *y, *z, *vx, *vy, *vz, *redshift, *theta, *phi;
    int64_t* id;
    int32_t *rotation, *replication;
} oracle_cols_out;

/* everything processLC.cpp:512-733 derives per halo */
typedef struct {
    float pos[3];
    float halo_r;         /* :539 */
    float half_box_rad;   /* :545 */
    float theta_cut[2];   /* :550-551 (arcsec) */
    float phi_cut[2];     /* :552-553 */
    float k[3];           /* :580 */
    float beta;           /* :581 */
    float K[9];           /* :588 */
    float R[9];           /* :589 */
    float R_inv[9];       /* :592 */
    float corners_rot[12];/* A,B,C,D rotated unit vectors, :665-668 */
    float corners_sph[8]; /* (theta,phi) of A,B,C,D, :680-694 */
    float theta_rough[2]; /* :718-719 */
    float phi_rough[2];   /* :723-724 */
} oracle_halo;

/* processLC.cpp:282-284 -- no x==0 guards (use case 1 applies it to first-octant particles only) */
void oracle_theta_phi_uc1(int64_t n, const float* x, const float* y, const float* z, float* theta, float* phi);
/* processLC.cpp:899-908 and :1425-1435 -- with the x==0 guards */
void oracle_theta_phi_uc2(int64_t n, const float* x, const float* y, const float* z, float* theta, float* phi);
/* util.cpp:492-500 */
float oracle_aToZ(float a);
/* util.cpp:620-643 */
void oracle_matvec(const float* m9, const float* v3, float* out3);
/* util.cpp:506-527, :533-552 */
int oracle_zToStep(float z, int totSteps, float maxZ);
float oracle_stepToZ(float step, int totSteps, float maxZ);
/* main.cpp:265-274: "-t center half" in degrees -> {lo,hi} in arcsec */
void oracle_cli_bounds(float center_deg, float half_deg, float* out2);

/* use case 1 hot loop, processLC.cpp:276-310.  returns K. */
int64_t oracle_cutout_const(const oracle_cols_in* in, const float* theta_cut2, const float* phi_cut2,
                            oracle_cols_out* out);

/* processLC.cpp:512-733 (+ util.cpp:649-665, :695-751, :778-902) */
void oracle_halo_setup(const float* pos3, float boxLength, oracle_halo* out);

/* util.cpp:97-110 as used at processLC.cpp:1003-1006: how many leading particles survive the
 * redistribution when the reference runs on `numranks` ranks with all Np particles read by one rank
 * (SURVEY.md App. A Q7). */
int64_t oracle_scatter_kept(int64_t Np, int numranks);

/* use case 2, processLC.cpp:897-909 + :1214-1221 + :1334-1340 + :1383-1481 for one halo.
 * ref_numranks: rank count of the reference run being modelled (1 = the parity target).
 * returns K; *rough_count = rough_cutout_size (:1436). theta/phi out = rotated angles (:1446-1447). */
int64_t oracle_cutout_halo(const oracle_cols_in* in, const oracle_halo* halo, int pos_only, int ref_numranks,
                           oracle_cols_out* out, int64_t* rough_count);

/* glibc passthroughs so GPU tests can compare the device restatement of acosf/atanf with the libm
 * the reference links. */
void oracle_acosf_array(int64_t n, const float* in, float* out);
void oracle_atanf_array(int64_t n, const float* in, float* out);
const char* oracle_libc_version(void);

#ifdef __cplusplus
}
#endif
#endif
