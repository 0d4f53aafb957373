/* tests/native/abi_stub.c -- TEST INFRASTRUCTURE.
 * The binding a maintainer of the reference would add, written in plain C against
 * include/cosmo_cutout_b200.h: proves the header is a C ABI (no C++ types) and that the documented call sequence
 * type-checks.  Compiled and linked against the product library by tests/test_host_cpu.py; run without a GPU it
 * must fail at cc_create with a message, which is what main() checks. */
#include <stdio.h>
#include <string.h>
#include "cosmo_cutout_b200.h"

static int use_case_1(cc_ctx* ctx, int64_t n, const cc_cols_in* in, const float th[2], const float ph[2], cc_cols_out* out,
                      int64_t out_rows, int64_t* k) {
    int64_t counts[1], offset[1];
    if (cc_step_upload(ctx, n, in, 0)) return 1;
    if (cc_cutout_const(ctx, th, ph)) return 1;
    if (cc_exchange_counts(ctx, counts, NULL, offset)) return 1; /* one rank: counts[0], offset 0 */
    *k = counts[0];
    if (*k > out_rows) return 2;
    return cc_cutout_fetch(ctx, 0, 0, *k, out);
}

static int use_case_2(cc_ctx* ctx, int64_t n, const cc_cols_in* in, int nhalos, const float* halo_pos, float box,
                      cc_cols_out* out_all, int64_t* counts, int64_t* rough) {
    cc_halo halos[16];
    int64_t total = 0;
    if (nhalos > 16) return 2;
    for (int h = 0; h < nhalos; ++h) {
        cc_halo_info info;
        if (cc_halo_setup(&halo_pos[3 * h], box, &info)) return 1;
        halos[h] = info.halo;
    }
    if (cc_step_upload(ctx, n, in, 0)) return 1;
    if (cc_cutout_halo(ctx, nhalos, halos, 0, cc_ref_scatter_kept(n, 1))) return 1;
    if (cc_cutout_sync(ctx, counts, rough)) return 1;
    for (int h = 0; h < nhalos; ++h) total += counts[h];
    return cc_cutout_fetch_all(ctx, total, out_all);
}

int main(void) {
    cc_ctx* ctx = NULL;
    float b[2];
    cc_cli_bounds(87.0f, 2.0f, b);
    printf("version %s, devices %d, theta bounds %.1f %.1f\n", cc_version(), cc_device_count(), b[0], b[1]);
    if (cc_device_count() == 0) {
        const int rc = cc_create(0, &ctx);
        printf("cc_create without a GPU: rc=%d (%s)\n", rc, cc_last_error());
        return rc != CC_OK && ctx == NULL && strlen(cc_last_error()) > 0 ? 0 : 1;
    }
    (void)use_case_1; (void)use_case_2;
    return 0;
}
