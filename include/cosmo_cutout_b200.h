/*
 * cosmo_cutout_b200.h -- C ABI of the B200-native lightcone cutout path.
 *
 * This is the drop-in boundary for the hot path of jhollowed/cosmo-cutout.  The reference has no FFI
 * layer: its boundary is the two C++ overloads `processLC(...)` (reference src/processLC.h:41-48) called
 * from `main` (src/main.cpp:386-393).  Both overloads are re-implemented in C++ in
 * cosmo-cutout_b200/host/processLC.cpp with the reference's signatures, and they reach the GPU only through
 * the entry points below (plain pointers and sizes, no C++/torch types).  Each entry point cites the piece
 * of the reference it replaces (paths relative to the reference's src/).
 *
 * Conventions
 *   - every function returns CC_OK (0) or a negative CC_ERR_* code; cc_last_error() gives the message of
 *     the calling thread's last failure.  Nothing throws across this boundary.
 *   - the caller owns every buffer it passes in; a context owns its device memory and staging until
 *     cc_destroy().  A context is bound to ONE GPU and is not thread-safe; use one context per GPU.
 *   - there is NO CPU fallback: every compute entry point fails with CC_ERR_CUDA when no sm_100 device is
 *     usable.
 *   - all angles are arcseconds held in `float`, exactly as the reference holds them (processLC.h:35-36).
 */
#ifndef COSMO_CUTOUT_B200_H
#define COSMO_CUTOUT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CC_OK 0
#define CC_ERR_ARG (-1)      /* bad argument */
#define CC_ERR_CUDA (-2)     /* CUDA runtime / driver failure, or no usable GPU */
#define CC_ERR_STATE (-3)    /* call order (no step resident, no cutout pending, ...) */
#define CC_ERR_NCCL (-4)     /* NCCL failure or libnccl not loadable */
#define CC_ERR_NOMEM (-5)    /* device or pinned-host allocation failed */

typedef struct cc_ctx cc_ctx;

/* Input columns of one lightcone step: reference util.h:33-47 (struct Buffers_read).
 * vx,vy,vz,rotation,replication may be NULL when only position-only halo cutouts will be requested
 * (processLC.cpp:877-883 reads them only if !positionOnly). */
typedef struct cc_cols_in {
    const float *x, *y, *z, *vx, *vy, *vz, *a;
    const int64_t* id;
    const int32_t *rotation, *replication;
} cc_cols_in;

/* Output columns: reference util.h:50-64 (struct Buffers_write) in the order survivors are written.
 * Any pointer may be NULL (that column is skipped).  Two extra columns the reference keeps only
 * implicitly: `theta_u` = un-rotated theta of use case 2 (the stable-sort key, processLC.cpp:1214-1221)
 * and `index` = global input index of the survivor. */
typedef struct cc_cols_out {
    float *x, *y, *z, *vx, *vy, *vz, *redshift, *theta, *phi;
    int64_t* id;
    int32_t *rotation, *replication;
    float* theta_u;
    int64_t* index;
} cc_cols_out;

/* Everything processLC.cpp:512-733 derives for one halo that the per-particle loop consumes. */
typedef struct cc_halo {
    float R[9];           /* row-major rotation matrix, util.cpp:877-902 */
    float theta_cut[2];   /* fine bounds, strict, processLC.cpp:550-551, :1439 */
    float phi_cut[2];     /* processLC.cpp:552-553, :1440 */
    float theta_rough[2]; /* rough bounds: lo <= theta_u < hi, processLC.cpp:718-719, :1334-1337 */
    float phi_rough[2];   /* rough bounds, strict, processLC.cpp:723-724, :1400 */
} cc_halo;

/* Full per-halo derivation, for logging / properties.csv (processLC.cpp:539-724). */
typedef struct cc_halo_info {
    float pos[3], halo_r, half_box_rad, k[3], beta, K[9], R[9], R_inv[9];
    float corners_rot[12], corners_sph[8];
    cc_halo halo;
} cc_halo_info;

/* ------------------------------------------------------------------ library / context */
const char* cc_version(void);
const char* cc_last_error(void);
int cc_device_count(void);
/* Create a context on CUDA device `device`.  Fails (CC_ERR_CUDA) unless the device is sm_100. */
int cc_create(int device, cc_ctx** out);
void cc_destroy(cc_ctx* ctx);
/* the context's stream as a cudaStream_t (so a caller can order its own work against it) */
void* cc_stream(cc_ctx* ctx);

/* ------------------------------------------------------------------ host-side helpers (O(1) per halo)
 * Per-halo bounds + rotation set-up: replaces processLC.cpp:512-733 and util.cpp:649-665, :695-751,
 * :778-902.  Runs on the host with the host libm, like the reference, so the bounds are the same bits. */
int cc_halo_setup(const float pos[3], float box_length_arcmin, cc_halo_info* out);
/* "-t center half" / "-p center half" in degrees -> {lo,hi} in arcsec: main.cpp:265-274 */
void cc_cli_bounds(float center_deg, float half_deg, float out[2]);
/* Side effect of the reference's use-case-2 load balancer (util.cpp:97-110, processLC.cpp:1003-1027):
 * with Np particles read by a rank and `numranks` ranks, only the first cc_ref_scatter_kept(Np,numranks)
 * particles survive the redistribution (the float rank map overflows for Np > 2^24). */
int64_t cc_ref_scatter_kept(int64_t np, int numranks);

/* ------------------------------------------------------------------ step residency (replaces the
 * GenericIO fill of Buffers_read, processLC.cpp:102-139 / :861-889: columns go to HBM instead) */
/* Copy n particles of host columns to HBM (pinned or pageable host memory). `index_base` is the global
 * index of element 0 of this shard (0 on a single GPU). */
int cc_step_upload(cc_ctx* ctx, int64_t n, const cc_cols_in* host, int64_t index_base);
/* Adopt caller-owned DEVICE columns without copying (they must outlive the step). */
int cc_step_attach(cc_ctx* ctx, int64_t n, const cc_cols_in* dev, int64_t index_base);
/* Fill a resident step with the counter-based synthetic lightcone (bit-identical to cc_synth_host for the
 * same seed and global indices).  kind 0: first-octant cube (0,L]^3, kind 1: full cube (-L,L)^3. */
int cc_step_synth(cc_ctx* ctx, int64_t n, uint64_t seed, int64_t index_base, int kind, float box);
int cc_synth_host(int64_t n, uint64_t seed, int64_t index_base, int kind, float box, float* x, float* y,
                  float* z, float* vx, float* vy, float* vz, float* a, int64_t* id, int32_t* rotation,
                  int32_t* replication);
/* Copy [first, first+count) of the resident step back to host columns (NULL columns skipped). */
int cc_step_download(cc_ctx* ctx, int64_t first, int64_t count, float* x, float* y, float* z, float* vx,
                     float* vy, float* vz, float* a, int64_t* id, int32_t* rotation, int32_t* replication);
int64_t cc_step_size(cc_ctx* ctx);
int cc_step_release(cc_ctx* ctx);
/* A context can keep SEVERAL steps resident, in numbered slots, and cut any of them: cc_step_select makes `slot` the
 * current step (slot 0 at creation; a slot never used is empty).  Upload / attach / synth / release and every cutout
 * act on the current slot; a streamed cutout replaces the current slot's resident step, so stream on a slot kept for
 * that.  This is what lets consecutive processLC calls (one boxLength each, processLC.h:45-48 -- BASELINE config 5)
 * find the lightcone steps of the previous call still in HBM instead of reading and uploading them again; it plays
 * the role of the reference's "sort once, cut many halos" index (processLC.cpp:1214-1221, :1237) across calls. */
int cc_step_select(cc_ctx* ctx, int slot);
/* free / total bytes of the context's device (capacity planning for resident steps: 44 B per particle and shard) */
int cc_device_mem(cc_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes);
/* input-column bytes this context has copied host -> device so far (cc_step_upload and the streamed cutouts) */
int64_t cc_bytes_uploaded(cc_ctx* ctx);

/* ------------------------------------------------------------------ the cutout
 * Use case 1 (constant bounds): replaces the loop processLC.cpp:276-310.  Survivors keep input order.
 * Asynchronous: kernels are enqueued on the context's stream. */
int cc_cutout_const(cc_ctx* ctx, const float theta_cut[2], const float phi_cut[2]);
/* Use case 2 (halo-centred, rotated): replaces processLC.cpp:897-909 + :1214-1221 + :1334-1340 +
 * :1383-1481 for `nhalos` halos in ONE pass over the resident step.  Survivors of each halo are ordered by
 * (un-rotated theta as float, input index), the order the reference's stable arg-sort produces.
 * n_limit_global: only particles with global index < n_limit_global are considered (pass
 * cc_ref_scatter_kept(N_total, 1) to reproduce the reference's tail drop, or a negative value for all). */
int cc_cutout_halo(cc_ctx* ctx, int nhalos, const cc_halo* halos, int pos_only, int64_t n_limit_global);
/* Wait for the pending cutout; counts[h] = survivors, rough_counts[h] = survivors of the rough cut
 * (processLC.cpp:1436; 0 for use case 1).  Either array may be NULL. */
int cc_cutout_sync(cc_ctx* ctx, int64_t* counts, int64_t* rough_counts);
/* Copy halo h's survivors (h = 0 for use case 1) to host columns, rows [first, first+count): the contents of the
 * reference's Buffers_write vectors (util.h:53-70) as filled by processLC.cpp:291-307 / :1446-1461. */
int cc_cutout_fetch(cc_ctx* ctx, int halo, int64_t first, int64_t count, const cc_cols_out* host);
/* Copy the survivors of ALL halos of the last cutout with one transfer per column: rows are packed halo after halo,
 * halo h at rows [sum(counts[0..h)), +counts[h]) (counts from cc_cutout_sync); `total` must be sum(counts).  With
 * hundreds of halos per pass (--haloFile, main.cpp:295-349) this replaces thousands of small copies; the rows are
 * what the reference hands to its per-halo writes, processLC.cpp:1628-1697. */
int cc_cutout_fetch_all(cc_ctx* ctx, int64_t total, const cc_cols_out* host);
/* Device pointers to the packed survivor columns of the last cutout (valid until the next cutout call):
 * halo h occupies rows [row0, row0+count). */
int cc_cutout_device_cols(cc_ctx* ctx, int halo, cc_cols_out* dev, int64_t* row0, int64_t* count);

/* ------------------------------------------------------------------ streamed step (replaces the GenericIO
 * read path for steps that are not resident: pinned host -> HBM double-buffered, processLC.cpp:102-139).
 * One call = upload x,y,z in tiles on a copy stream overlapped with the cutout kernel on the previous tile.
 * The other columns never cross the bus as a whole: cc_cutout_fetch / cc_cutout_fetch_all pick vx..replication of
 * the survivors from the caller's host columns by survivor index (reading them from the GPU 4 bytes at a time
 * over PCIe costs ~10 ns per value: 90 ms for 1.5 M survivors), so the host columns must stay valid and unchanged
 * until the survivors have been fetched, and cc_cutout_device_cols returns NULL for those columns.  Host columns
 * that are really device or managed memory, or CC_DEVICE_GATHER=1, keep the gather on the GPU.
 * Results are fetched with cc_cutout_sync / cc_cutout_fetch as usual. */
int cc_cutout_const_streamed(cc_ctx* ctx, int64_t n, const cc_cols_in* host, int64_t index_base,
                             const float theta_cut[2], const float phi_cut[2]);
int cc_cutout_halo_streamed(cc_ctx* ctx, int64_t n, const cc_cols_in* host, int64_t index_base, int nhalos,
                            const cc_halo* halos, int pos_only, int64_t n_limit_global);
/* pinned host allocations for the streamed path */
int cc_host_alloc(void** ptr, size_t bytes);
int cc_host_free(void* ptr);
int cc_host_register(void* ptr, size_t bytes);
int cc_host_unregister(void* ptr);

/* ------------------------------------------------------------------ multi-GPU: count exchange
 * Replaces MPI_Allgather(count) + serial prefix sum (processLC.cpp:322-337, :1546-1583): an NCCL
 * all-gather of every GPU's {count, rough_count} per halo over NVLink followed by a device exclusive scan.
 * One process per GPU: rank 0 obtains an id with cc_comm_unique_id, ships it to the other ranks by any
 * means, every rank calls cc_comm_init.  Single process, several GPUs: cc_comm_init_all. */
#define CC_COMM_ID_BYTES 128
int cc_comm_unique_id(void* id128);
int cc_comm_init(cc_ctx* ctx, int nranks, int rank, const void* id128);
int cc_comm_init_all(cc_ctx** ctxs, int n);
/* After cc_cutout_*: gathers counts of all ranks.  all_counts[r*nhalos+h], all_rough[r*nhalos+h] (host,
 * may be NULL); offsets[h] = sum of counts of ranks < this rank (the file offset, in rows). Implies a
 * cc_cutout_sync.  Without a communicator it degenerates to one rank.  The exchange is enqueued on the context's
 * second stream behind the device event "counts are final", so it runs next to the ordering / gather kernels of
 * the pending cutout; the cutout stream (cc_stream) waits for it.  Like any collective it must be entered by every
 * rank.  The communicator is created with one CTA (the payload is a few KB). */
int cc_exchange_counts(cc_ctx* ctx, int64_t* all_counts, int64_t* all_rough, int64_t* offsets);
int cc_comm_rank(cc_ctx* ctx);
int cc_comm_size(cc_ctx* ctx);
/* Optional fast path for the count exchange: peer stores over NVLink instead of a collective launch.  Every rank
 * allocates a mailbox and exports it (cc_comm_peer_handle, a cudaIpcMemHandle_t); the handles of all ranks, in rank
 * order, are handed to cc_comm_peer_attach.  cc_exchange_counts then uses one single-CTA kernel that stores this
 * GPU's counters into every peer's mailbox and waits for theirs (up to 2048 halos per cutout; more falls back to
 * NCCL).  cc_comm_init_all sets this up by itself when the GPUs of the process have peer access. */
#define CC_PEER_HANDLE_BYTES 64
int cc_comm_peer_handle(cc_ctx* ctx, int nranks, int rank, void* handle64);
int cc_comm_peer_attach(cc_ctx* ctx, const void* handles);

/* ------------------------------------------------------------------ multi-GPU: host-side plan (no GPU needed)
 * Contiguous index shard of rank r out of nranks over n particles: [lo, hi).  Replaces the reference's
 * per-rank GenericIO blocks + load-balancing redistribution (processLC.cpp:974-1125, util.cpp:97-110). */
void cc_shard_range(int64_t n, int nranks, int rank, int64_t* lo, int64_t* hi);
/* Use case 2, several ranks: every rank's survivors of a halo arrive ordered by (theta_u, index); a single-rank
 * reference run orders the WHOLE step that way (stable arg-sort, processLC.cpp:1214-1221), so the file order is
 * the merge of the per-rank lists on those keys.  theta_u[r], index[r]: rank r's key columns (counts[r] rows).
 * Output: for merged row i, src_rank[i] / src_row[i] say where it comes from.  Rows total = sum(counts). */
int cc_merge_order(int nranks, const int64_t* counts, const float* const* theta_u, const int64_t* const* index,
                   int32_t* src_rank, int64_t* src_row);

/* ------------------------------------------------------------------ measurement */
/* Device time (CUDA events on the context's stream) of the scan kernel of the last cutout, and of the
 * whole enqueued cutout (scan + sort + gather).  ms; negative if none -- the events are only recorded after
 * cc_set_option(ctx, "timing", "on") (three event records between the kernels of every cutout otherwise).
 * Implies a stream synchronize. */
double cc_last_scan_ms(cc_ctx* ctx);
double cc_last_total_ms(cc_ctx* ctx);
/* number of kernels this context has launched so far */
int64_t cc_launch_count(cc_ctx* ctx);
/* statistics of the last cutout: [0] particles scanned, [1] pre-filter candidates (exact path taken),
 * [2] survivors, [3] rough survivors */
int cc_last_stats(cc_ctx* ctx, int64_t out[4]);
/* the same and more, first n of: [0..3] as cc_last_stats; [4] use case 2: scanned particles whose un-rotated theta is
 * NaN (NaN coordinate, z = +-Inf, or x = y = z = 0).  The reference orders particles by that angle with
 * std::stable_sort and std::lower_bound (processLC.cpp:1214-1221, :1334-1337); a NaN key breaks their ordering
 * contract and what the reference then returns for the OTHER particles is unspecified, so parity is only claimed for
 * steps where this is 0 (lc_cutout warns otherwise); [5] 1 when the interleaved gather payload of the resident step
 * exists; [6] microseconds the last cutout call spent building it (0: it did not); [7] use case 2: particles of this
 * shard behind n_limit_global (the reference's tail drop, util.cpp:106-108); [8] which scan ran: 0 use case 1,
 * 1 use case 2 every particle against every halo (k2_cutout_halo), 2 the cell index walked by the staged kernels
 * (k2c_sift, k2c_match, k2c_exact), 3 by one kernel (k2_cutout_halo_cells), 4 the particle index of the resident
 * step (k2p_query); [9] particles the scan read: all scanned ones, or with [8] = 4 only the index rows under the
 * halos' windows; [10] 1 when the particle index of the resident step exists; [11] microseconds the last cutout call
 * spent building it (0: it did not). */
#define CC_STATS_EX_FIELDS 12
int cc_last_stats_ex(cc_ctx* ctx, int64_t* out, int n);

/* ------------------------------------------------------------------ tuning switches (defaults are the measured best)
 * key "payload":  "auto" (default) | "always" | "never" -- the 32 B/particle interleaved copy of the seven gather-only
 *                 columns of a resident step; auto builds it inside the first cutout that follows a cutout with at
 *                 least max(65536, n/512) survivors on the same resident step
 * key "exchange": "auto" (default: NVLink peer mailboxes when attached, else NCCL) | "nccl" | "peer" | "eager" (peer, and a
 *                 small use-case-2 cutout posts its counts from inside its last kernel: every rank must then follow
 *                 every cutout by cc_exchange_counts)
 * key "cells":    "auto" (default) | "staged" | "mono" -- many halos: the cell index is walked by three staged kernels or by
 *                 one; auto measures the bitmap hit rate of a halo set on its first pass and picks from it
 * key "index":    "auto" (default) | "always" | "never" -- the particle index of a resident step: {x, y, z, index} in
 *                 (cos theta_u, sin phi_u) cell order (16 B/particle).  A many-halo cutout (>= 8 halos) on an indexed
 *                 step reads only the cells under its halos' windows instead of the whole step -- what the reference's
 *                 theta sort + lower_bound does (processLC.cpp:1214-1221, :1334-1340).  auto builds it inside the second
 *                 many-halo cutout on the same resident step of >= 4 Mi particles; results are identical either way
 * key "timing":   "off" (default) | "on" -- record the per-kernel CUDA events behind cc_last_scan_ms / cc_last_total_ms
 * key "gather":   "auto" (default) | "ldst" | "bulk" -- use case 1 gather: plain coalesced stores, or rows staged in
 *                 shared memory and written with cp.async.bulk (TMA bulk stores); auto picks bulk stores for dense
 *                 cutouts (profiles/r02 has the A/B numbers) */
int cc_set_option(cc_ctx* ctx, const char* key, const char* value);

/* ------------------------------------------------------------------ test hooks (device arithmetic)
 * Evaluate the device restatement of the reference arithmetic element-wise on HOST arrays (copied to the
 * GPU and back): op 0 = acosf, 1 = atanf (in0 only); 2 = theta(x,y,z) -> out0, 3 = use-case-1 phi(x,y),
 * 4 = use-case-2 phi(x,y) with the x==0 guards; 5 = aToZ(in0); 6 = radians -> arcsec scaling of in0 (the device
 * evaluates its division-free form, cc_debug_eval_host the specification form with a true fp64 division). */
int cc_debug_eval(cc_ctx* ctx, int op, int64_t n, const float* in0, const float* in1, const float* in2,
                  float* out0);
/* host build of the same restatement (ref_math.cuh compiled for the CPU), same ops */
int cc_debug_eval_host(int op, int64_t n, const float* in0, const float* in1, const float* in2, float* out0);

#ifdef __cplusplus
}
#endif
#endif
