/*
 * capi.cu -- implementation of include/cosmo_cutout_b200.h: contexts, HBM column residency, kernel launches,
 * result staging and the NCCL count exchange.  No CPU compute path exists here: every cutout entry point
 * launches the sm_100a kernels of kernels.cuh or fails.
 */
#include "../../include/cosmo_cutout_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <omp.h> /* types and enums only; the library itself is dlopen()ed on first use */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "prefilter.h"

namespace {

thread_local std::string g_err;
/* the last failure of ANY thread: what cc_last_error returns to a thread that has none of its own (a caller that
 * checks the result of work it handed to another thread) */
std::mutex g_err_any_mutex;
std::string g_err_any;
thread_local std::string g_err_copy;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    {
        std::lock_guard<std::mutex> lock(g_err_any_mutex);
        g_err_any = buf;
    }
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA, "%s failed: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)

}  // namespace
namespace ccint {
int fail_msg(int code, const char* msg) { return fail(code, "%s", msg); } /* for host_setup.cpp */
}
namespace {

/* ---------------------------------------------------------------- NCCL, loaded lazily */
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr; /* optional */
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.handle) return CC_OK;
    const char* names[] = {getenv("CC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(CC_ERR_NCCL, "cannot dlopen libnccl.so.2 (%s); set CC_NCCL_LIB", dlerror());
#define SYM(field, name)                                                       \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                 \
    if (!g_nccl.field) return fail(CC_ERR_NCCL, "libnccl lacks symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommInitAll, "ncclCommInitAll");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    *(void**)(&g_nccl.CommInitRankConfig) = dlsym(h, "ncclCommInitRankConfig");
    g_nccl.handle = h;
    return CC_OK;
}
#define NC(call)                                                                                  \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess) return fail(CC_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

/* ---------------------------------------------------------------- device buffer that only grows */
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return CC_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(CC_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
        }
        bytes = need;
        return CC_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

enum { COL_X, COL_Y, COL_Z, COL_VX, COL_VY, COL_VZ, COL_A, COL_ID, COL_ROT, COL_REP, N_IN_COLS };
const size_t IN_COL_SIZE[N_IN_COLS] = {4, 4, 4, 4, 4, 4, 4, 8, 4, 4};
enum { OC_X, OC_Y, OC_Z, OC_VX, OC_VY, OC_VZ, OC_RED, OC_TH, OC_PH, OC_ID, OC_ROT, OC_REP, OC_THU, OC_IDX, N_OUT_COLS };
const size_t OUT_COL_SIZE[N_OUT_COLS] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 8, 4, 4, 4, 8};

enum Pending { P_NONE = 0, P_CONST = 1, P_HALO = 2 };

/* a resident step that is not the context's current one (cc_step_select) */
struct StepSlot {
    bool valid = false, owned = false;
    int64_t n = 0, index_base = 0;
    const void* col[N_IN_COLS] = {};
    DevBuf col_buf[N_IN_COLS];
    DevBuf payload;
    bool payload_valid = false;
    int64_t last_survivors = -1;
    DevBuf pidx_sorted, pidx_start; /* particle index of the step (kernels.cuh k2p_*) */
    bool pidx_valid = false;
    uint64_t pidx_gen = 0;
    int64_t many_halo_cutouts = 0;
    void release() { for (auto& b : col_buf) b.release(); payload.release(); pidx_sorted.release(); pidx_start.release(); *this = StepSlot(); }
};

}  // namespace

struct cc_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_scan0 = nullptr, ev_scan1 = nullptr, ev_total1 = nullptr;
    /* the survivor counts are final long before the survivors are ordered and gathered: the count exchange runs on
     * its own stream from that point on, next to the ordering / gather kernels */
    cudaEvent_t ev_counts = nullptr, ev_xchg = nullptr;
    cudaStream_t xchg_stream = nullptr;
    int64_t launches = 0;
    int64_t last_const_count = -1;
    int k2_occupancy[2] = {0, 0};
    int cells_mode = 0;        /* many halos: 0 auto (per halo set, from the measured bitmap hit rate), 1 staged
                                  (k2c_sift / k2c_match / k2c_exact), 2 one kernel (k2_cutout_halo_cells); cc_set_option
                                  "cells" / CC_CELLS */
    bool cells_mono_ready = false;
    bool used_mono = false;    /* the pending cutout ran the one-kernel scan */
    DevBuf cand, cand_count, pair_p, pair_h, pair_count;
    unsigned int cand_region_cap = 0; /* candidates per warp region (k2c_sift) */
    int64_t pair_cap = 0;
    int sift_occupancy[2] = {0, 0};
    bool k3_smem_set = false, k3s_smem_set = false;
    bool k2_tma = false, k2t_ready = false;
    int k2t_occ = 0;
    int64_t last_nrec = -1;       /* survivors of the previous use-case-2 cutout on this context */
    bool force_general = false, used_small = false;
    unsigned int region_cap_w = 0; /* temp records per WARP region (use case 1) */
    int gather_mode = 0;         /* use case 1 gather: 0 auto, 1 k1s_gather (plain stores), 2 k1s_gather_bulk (cp.async.bulk
                                    stores from shared memory); cc_set_option "gather" / CC_GATHER */
    bool gather_bulk_ready = false;
    int k1s_occupancy[2] = {0, 0};
    DevBuf t_idx, t_th, t_ph, t_x, t_y, t_z;

    /* resident steps: the fields below are the CURRENT one; the others wait in `slots` (cc_step_select) */
    std::vector<StepSlot> slots;
    int cur_slot = 0;
    int64_t bytes_uploaded = 0; /* input-column bytes copied host -> device so far (cc_bytes_uploaded) */
    bool step_valid = false, step_owned = false;
    int64_t n = 0, index_base = 0;
    const void* col[N_IN_COLS] = {};
    DevBuf col_buf[N_IN_COLS];
    DevBuf payload;          /* interleaved gather columns (kernels.cuh Payload): built lazily, see maybe_build_payload */
    bool payload_valid = false;
    int payload_policy = 0;  /* 0 auto (second dense cutout on a resident step), 1 always, 2 never; cc_set_option "payload" */
    int64_t step_last_survivors = -1; /* survivors of the previous cutout on the resident step (-1: none yet) */
    cudaEvent_t ev_pay0 = nullptr, ev_pay1 = nullptr;
    bool payload_timed = false;       /* the pending cutout built the payload: ev_pay0..ev_pay1 is its duration */
    /* Particle index of the resident step (kernels.cuh k2p_*): {x, y, z, index} in (cos theta_u, sin phi_u) cell order +
     * the first row of every cell.  A many-halo cutout on an indexed step reads only the cells under its halos' windows.
     * Built lazily like the payload: policy auto = inside the SECOND many-halo cutout on the same resident step (a step
     * that is cut once is cheaper to scan); 16 B/particle + 2 x 16 MB of HBM; dropped when another step becomes resident. */
    DevBuf pidx_sorted, pidx_start, pidx_tmp, pidx_partial, pidx_work;
    bool pidx_valid = false;
    uint64_t pidx_gen = 0, pidx_gen_clock = 0; /* which build of the index this is */
    int index_policy = 0;             /* 0 auto, 1 always, 2 never; cc_set_option "index" / CC_INDEX */
    int64_t step_many_halo_cutouts = 0;
    cudaEvent_t ev_idx0 = nullptr, ev_idx1 = nullptr;
    bool index_timed = false;         /* the pending cutout built the index: ev_idx0..ev_idx1 is its duration */
    bool used_index = false;          /* the pending cutout walked the particle index */
    unsigned long long small_seq = 0; /* k3_small launches so far: the value its last launch writes to h_small[7] */
    bool sync_whole_stream = false;   /* cc_cutout_sync: wait for the stream even when the flag would do (cc_exchange_counts) */
    int64_t particles_read = 0;       /* by the scan kernel(s) of the last cutout: all scanned ones, or the index rows */
    int xchg_mode = 0;       /* 0 auto (peer mailboxes when attached, else NCCL), 1 NCCL, 2 peer, 3 eager: peer, and a
                                small use-case-2 cutout does the exchange inside its last kernel; cc_set_option "exchange" */
    bool eager_done = false; /* the pending cutout's last kernel carried the count exchange */

    /* survivors */
    DevBuf out_buf[N_OUT_COLS];
    int64_t out_cap = 0;

    /* scan scratch: [tile_state ... | ticket | result(2) ] */
    DevBuf scratch;
    /* use case 2 scratch */
    DevBuf rec_a, rec_b, rank_buf, place_buf, counts_buf;
    DevBuf hstate; /* use case 2 device state, one memset / one read-back: [misc 8 | counts 2H | seg H+1 | cursor 16H] u64 */
    /* Device-resident parameters (+ cell index) of the halo sets recently cut on this context.  A processLC call cuts
     * every step with ONE halo set, and a batch job (lc_cutout -f with several box sizes, BASELINE config 5) alternates
     * between a few: each set is prepared once (1000 halos: ~1 ms of trigonometry + the rasterisation of the cell
     * index) and found again by comparing the cc_halo arrays.  Least recently used set is replaced. */
    struct HaloSet {
        std::vector<cc_halo> key;
        bool bins = false;
        DevBuf pre_buf, halo_buf;
        DevBuf bin_start_buf, bin_bits_buf, bin_items_buf; /* cell index: {first,end} per cell, occupancy bitmap, halo lists */
        std::vector<size_t> bin_items_off;                 /* per launch batch: first item */
        DevBuf pidx_items_buf;                             /* particle-index walker: (halo, grid line, columns) items */
        std::vector<size_t> pidx_items_off;                /* per launch batch: first item, number of items */
        std::vector<unsigned int> pidx_items_cnt;
        int walker = 0;     /* 0: not measured yet (the staged kernels measure), 1: staged, 2: one kernel */
        uint64_t last_use = 0;
        void release() { pre_buf.release(); halo_buf.release(); bin_start_buf.release(); bin_bits_buf.release(); bin_items_buf.release(); pidx_items_buf.release(); key.clear(); }
    };
    static const int MAX_HALO_SETS = 8;
    HaloSet halo_sets[MAX_HALO_SETS];
    HaloSet* hs = nullptr;   /* the set of the pending cutout */
    uint64_t halo_set_clock = 0;
    bool use_bins = false;
    int64_t rec_cap = 0;

    /* pinned host mirror for small read-backs */
    unsigned long long* h_small = nullptr; /* [16 + 2*H_cap + ...] */
    size_t h_small_elems = 0;

    /* pending cutout */
    Pending pending = P_NONE;
    bool synced = false;
    float p_th[2] = {}, p_ph[2] = {};
    std::vector<cc_halo> p_halos;
    int p_pos_only = 0;
    int64_t p_n_limit = -1;
    int nhalos = 0; /* of the pending/last cutout (1 for use case 1) */
    std::vector<int64_t> counts, rough, seg_begin;
    int64_t stats[4] = {};
    int64_t nan_theta = 0;  /* use case 2: scanned particles whose un-rotated theta is NaN (see cc_last_stats_ex) */
    bool have_times = false;
    int prefetch = -1;       /* scan kernels ask the L2 for the warp's next tile while they work on this one: -1 auto, 0 off,
                                1 on (CC_PREFETCH).  Measured (profiles/r02): it hides DRAM latency where the scan is issue- or
                                latency-limited (wide field +6 %, 10 M particles +18 %, halo scan +3 %, sift +2 %) and costs
                                5 % where the scan already saturates HBM (1 B particles, narrow window) */
    bool timing = false;     /* record the per-kernel timing events (cc_last_scan_ms / cc_last_total_ms); cc_set_option
                                "timing": three event records per cutout sit between the kernels, so it is off unless asked */
    bool hstate_clean = false; /* use case 2 device state was zeroed by the previous cutout's last kernel (k3_small) */
    int hstate_clean_H = -1;
    DevBuf xchg_send;          /* use case 2, k3_small path: {counts, rough counts} saved for the count exchange before the
                                  device state is zeroed */
    const void* xchg_src = nullptr; /* what cc_exchange_counts sends for the pending use-case-2 cutout */

    /* streamed step (pinned host -> HBM double buffering) */
    bool streamed = false;
    bool host_gather = false;    /* streamed step: vx..replication of the survivors are picked from the host columns by
                                    the host when they are fetched, not read across PCIe 4 bytes at a time by the GPU */
    std::vector<int64_t> fetch_index; /* scratch of fetch_rows */
    cc_cols_in host_cols = {};
    cck::ColsIn gcols = {};      /* where survivor columns are gathered from */
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {}, ev_consumed[2] = {};
    DevBuf stage[2][3];
    int64_t stream_chunk = 16 << 20; /* particles per staged chunk */

    /* communicator */
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    DevBuf gather_buf, offsets_buf;
    /* peer-store exchange (kernels.cuh k4_exchange_peer) */
    DevBuf mailbox;
    unsigned long long* peer_mb[cck::XCHG_MAX_RANKS] = {};
    bool peer_opened[cck::XCHG_MAX_RANKS] = {};
    bool peers_ready = false;
    unsigned long long xchg_epoch = 0;
};

namespace {

cck::ColsIn cols_in_of(const cc_ctx* c) {
    cck::ColsIn ci;
    ci.x = (const float*)c->col[COL_X]; ci.y = (const float*)c->col[COL_Y]; ci.z = (const float*)c->col[COL_Z];
    ci.vx = (const float*)c->col[COL_VX]; ci.vy = (const float*)c->col[COL_VY]; ci.vz = (const float*)c->col[COL_VZ];
    ci.a = (const float*)c->col[COL_A]; ci.id = (const long long*)c->col[COL_ID];
    ci.rot = (const int*)c->col[COL_ROT]; ci.rep = (const int*)c->col[COL_REP];
    return ci;
}
cck::ColsOut cols_out_of(const cc_ctx* c) {
    cck::ColsOut o;
    o.x = (float*)c->out_buf[OC_X].p; o.y = (float*)c->out_buf[OC_Y].p; o.z = (float*)c->out_buf[OC_Z].p;
    o.vx = (float*)c->out_buf[OC_VX].p; o.vy = (float*)c->out_buf[OC_VY].p; o.vz = (float*)c->out_buf[OC_VZ].p;
    o.redshift = (float*)c->out_buf[OC_RED].p; o.theta = (float*)c->out_buf[OC_TH].p; o.phi = (float*)c->out_buf[OC_PH].p;
    o.id = (long long*)c->out_buf[OC_ID].p; o.rot = (int*)c->out_buf[OC_ROT].p; o.rep = (int*)c->out_buf[OC_REP].p;
    o.theta_u = (float*)c->out_buf[OC_THU].p; o.index = (long long*)c->out_buf[OC_IDX].p;
    return o;
}

int ensure_out_cap(cc_ctx* c, int64_t cap) {
    if (cap <= c->out_cap) return CC_OK;
    cap = std::max<int64_t>(cap, 1024);
    for (int i = 0; i < N_OUT_COLS; ++i) {
        int rc = c->out_buf[i].ensure((size_t)cap * OUT_COL_SIZE[i]);
        if (rc) { c->out_cap = 0; return rc; }
    }
    c->out_cap = cap;
    return CC_OK;
}

int ensure_small(cc_ctx* c, size_t elems) {
    if (elems <= c->h_small_elems) return CC_OK;
    if (c->h_small) cudaFreeHost(c->h_small);
    c->h_small = nullptr;
    c->h_small_elems = 0;
    CU(cudaMallocHost((void**)&c->h_small, elems * sizeof(unsigned long long)));
    memset(c->h_small, 0, elems * sizeof(unsigned long long));
    c->h_small_elems = elems;
    return CC_OK;
}

bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

int use_device(cc_ctx* c) {
    CU(cudaSetDevice(c->device));
    return CC_OK;
}

/* ---------------------------------------------------------------- scan sources
 * A cutout scans x,y,z either from the resident step or, for a streamed step, from a staging slot that holds
 * one chunk; the columns gathered for survivors come from `gcols` (resident columns, or the pinned+mapped
 * host columns of a streamed step). */
struct Chunk {
    const float *x, *y, *z; /* device pointers to the chunk's first particle */
    int64_t n;              /* particles in the chunk */
    int64_t start;          /* index of the chunk's first particle inside the step */
    bool first, last;
};

const size_t K1_HDR = 64;

/* ---- use case 1 (kernels.cuh "K1") ------------------------------------------------------------------------------
 * scratch: [ result: 4 x u64 | ticket u32 (+pad to 64 B) | block_state: nblocks x u64 | seg_info: nseg x u64 |
 *            seg_row: nseg x u32 ]
 * result[0] survivors so far (carried across the chunks of a streamed step), [1] pre-filter candidates,
 * [2] carry seen by the current chunk's gather, [3] max records any warp wanted (record-region overflow check) */
typedef void (*k1s_scan_fn)(const cck::SegScanArgs);
k1s_scan_fn k1s_scan_variant(bool vec) { return vec ? cck::k1s_scan<true> : cck::k1s_scan<false>; }

int launch_const_chunk(cc_ctx* c, const Chunk& ch) {
    if (ch.n > 0xfffffff0ll) return fail(CC_ERR_ARG, "chunk too large (%lld particles)", (long long)ch.n);
    const unsigned int ntiles = (unsigned int)((ch.n + cck::SCAN_TILE - 1) / cck::SCAN_TILE);
    const unsigned int nseg = (ntiles + cck::K1S_TILES - 1) / cck::K1S_TILES;
    const unsigned int nblocks = (nseg + cck::TSCAN_BLOCK - 1) / cck::TSCAN_BLOCK;
    const bool vec = aligned16(ch.x) && aligned16(ch.y) && aligned16(ch.z);
    int& occ = c->k1s_occupancy[vec ? 1 : 0];
    if (occ == 0) {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k1s_scan_variant(vec), cck::SCAN_THREADS, 0));
        if (occ < 1) return fail(CC_ERR_CUDA, "k1s_scan cannot be resident on this device");
    }
    const unsigned int grid_max = (unsigned int)(c->sm_count * occ);
    const unsigned int grid = std::max(1u, std::min((nseg + cck::SCAN_WARPS - 1) / cck::SCAN_WARPS, grid_max));
    const unsigned int warps_max = grid_max * cck::SCAN_WARPS, total_warps = grid * cck::SCAN_WARPS;
    int rc;
    const size_t off_state = K1_HDR;
    const size_t off_info = off_state + ((size_t)nblocks + 8) * 8;
    const size_t off_row = off_info + ((size_t)nseg + 8) * 8;
    const size_t bytes = off_row + ((size_t)nseg + 8) * 4;
    if ((rc = c->scratch.ensure(bytes))) return rc;
    /* records: one private region per warp */
    c->region_cap_w = std::max(c->region_cap_w, (unsigned int)(std::min<int64_t>(0x7fffffff, c->out_cap + c->out_cap / 4) / warps_max + 256));
    const size_t temp_elems = (size_t)warps_max * c->region_cap_w;
    DevBuf* temps[] = {&c->t_idx, &c->t_th, &c->t_ph, &c->t_x, &c->t_y, &c->t_z};
    for (DevBuf* t : temps)
        if ((rc = t->ensure(temp_elems * 4))) return rc;
    char* sp = (char*)c->scratch.p;
    if (ch.first) CU(cudaMemsetAsync(sp, 0, off_info, c->stream));       /* totals + ticket + block states */
    else CU(cudaMemsetAsync(sp + 32, 0, off_info - 32, c->stream));      /* keep the running totals */
    unsigned long long* result = (unsigned long long*)sp;
    if (ch.first && c->timing) CU(cudaEventRecord(c->ev_scan0, c->stream));
    if (nseg > 0) {
        cck::SegScanArgs a;
        a.x = ch.x; a.y = ch.y; a.z = ch.z;
        a.n = ch.n;
        a.th_lo = c->p_th[0]; a.th_hi = c->p_th[1]; a.ph_lo = c->p_ph[0]; a.ph_hi = c->p_ph[1];
        {
            float clo2, tlo;
            ccpf::uc1_bounds(c->p_th, c->p_ph, &a.chi2, &clo2, &tlo, &a.thi);
            a.nclo2 = -clo2;
            a.ntlo = -tlo;
        }
        a.t_idx = (unsigned int*)c->t_idx.p; a.t_th = (float*)c->t_th.p; a.t_ph = (float*)c->t_ph.p;
        a.t_x = (float*)c->t_x.p; a.t_y = (float*)c->t_y.p; a.t_z = (float*)c->t_z.p;
        a.region_cap = c->region_cap_w;
        a.seg_info = (unsigned long long*)(sp + off_info);
        a.result = result;
        a.nseg = nseg;
        a.ntiles = ntiles;
        /* auto: a dense window (survivors of the previous cutout) or a step too short to fill the memory pipeline */
        a.prefetch = !c->streamed && (c->prefetch == 1 || (c->prefetch < 0 && (ch.n <= (64ll << 20) ||
                                                                              (c->last_const_count >= 0 && c->last_const_count * 256 >= ch.n))));
        k1s_scan_variant(vec)<<<grid, cck::SCAN_THREADS, 0, c->stream>>>(a);
        CU(cudaGetLastError());
        if (ch.last && c->timing) CU(cudaEventRecord(c->ev_scan1, c->stream));

        cck::SegPrefixArgs t;
        t.seg_info = a.seg_info;
        t.seg_row = (unsigned int*)(sp + off_row);
        t.block_state = (unsigned long long*)(sp + off_state);
        t.ticket = (unsigned int*)(sp + 32);
        t.result = result;
        t.nseg = nseg;
        t.nblocks = nblocks;
        t.host_result = ch.last ? c->h_small : nullptr; /* the step's totals go straight to pinned host memory */
        cck::k1s_seg_scan<<<nblocks, cck::TSCAN_THREADS, 0, c->stream>>>(t);
        CU(cudaGetLastError());
        if (ch.last) CU(cudaEventRecord(c->ev_counts, c->stream));

        cck::SegGatherArgs g;
        g.in = c->gcols;
        /* survivor columns other than x,y,z are indexed by the chunk-local particle index */
        g.in.vx += ch.start; g.in.vy += ch.start; g.in.vz += ch.start; g.in.a += ch.start; g.in.id += ch.start;
        g.in.rot += ch.start; g.in.rep += ch.start;
        if (c->streamed && c->host_gather) g.in = cck::ColsIn{}; /* the host picks these columns itself (fetch_rows) */
        g.out = cols_out_of(c);
        g.out.theta_u = nullptr;
        g.t_idx = a.t_idx; g.t_th = a.t_th; g.t_ph = a.t_ph; g.t_x = a.t_x; g.t_y = a.t_y; g.t_z = a.t_z;
        g.seg_info = a.seg_info;
        g.seg_row = t.seg_row;
        g.result = result;
        g.payload = (c->streamed || !c->payload_valid) ? nullptr : (const cck::Payload*)c->payload.p + ch.start;
        g.region_cap = c->region_cap_w;
        g.total_warps = total_warps;
        g.nseg = nseg;
        g.cap = c->out_cap;
        g.index_base = c->index_base + ch.start;
        const unsigned int ggrid = (nseg + cck::K1G_SEGS - 1) / cck::K1G_SEGS;
        /* bulk stores pay when a CTA's block of rows fills the staging buffer (measured: 3 % faster gather at 4.5 %
         * survivors, 2x slower at 0.2 %, profiles/r02): "auto" decides on the survivors of the previous cutout */
        const int64_t expect = c->last_const_count >= 0 ? c->last_const_count : 0;
        const bool bulk = c->gather_mode == 2 || (c->gather_mode == 0 && expect / std::max(1u, ggrid) >= cck::K1GB_ROWS);
        if (bulk) {
            if (!c->gather_bulk_ready) {
                CU(cudaFuncSetAttribute((const void*)cck::k1s_gather_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cck::K1GB_SMEM));
                c->gather_bulk_ready = true;
            }
            cck::k1s_gather_bulk<<<ggrid, cck::K1G_THREADS, cck::K1GB_SMEM, c->stream>>>(g);
        } else {
            cck::k1s_gather<<<ggrid, cck::K1G_THREADS, 0, c->stream>>>(g);
        }
        CU(cudaGetLastError());
        c->launches += 3;
    } else if (ch.last) {
        if (c->timing) CU(cudaEventRecord(c->ev_scan1, c->stream));
        CU(cudaEventRecord(c->ev_counts, c->stream));
        CU(cudaMemcpyAsync(c->h_small, result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    }
    if (ch.last) {
        if (c->timing) CU(cudaEventRecord(c->ev_total1, c->stream));
        c->have_times = c->timing;
    }
    return CC_OK;
}


/* pinned read-back area: [0,32+3H) the cutout's own results, then the exchanged counts of all ranks + offsets */
size_t small_elems(int H, int R) { return 32 + 3 * (size_t)H + (size_t)R * 2 * H + H + 24; }

int prepare_const(cc_ctx* c) {
    { int rc0 = ensure_small(c, small_elems(1, c->nranks)); if (rc0) return rc0; }
    /* first guess: 1/16 of the step; cc_cutout_sync grows it to the exact count and re-runs on overflow */
    {
        int rc = ensure_out_cap(c, std::max<int64_t>(c->n / 16, 1 << 16));
        if (rc) return rc;
    }
    return CC_OK;
}

/* bitmap hit rate above which a halo set is walked by the one-kernel scan (profiles/r02: 10 % hits: staged 0.59 ms vs
 * 0.9 ms; 22 % hits: staged 0.94 ms vs 0.71 ms) */
#ifndef CC_CELLS_MONO_ABOVE
#define CC_CELLS_MONO_ABOVE 0.16
#endif

/* ---------------------------------------------------------------- use case 2 launch */
size_t hstate_elems(int H) { return 8 + 2 * (size_t)H + ((size_t)H + 1) + (size_t)cck::K3_CURSOR_STRIDE * (size_t)H; }
unsigned long long* hs_misc(const cc_ctx* c) { return (unsigned long long*)c->hstate.p; }
unsigned long long* hs_counts(const cc_ctx* c) { return hs_misc(c) + 8; }
unsigned long long* hs_seg(const cc_ctx* c, int H) { return hs_counts(c) + 2 * (size_t)H; }
unsigned long long* hs_cursor(const cc_ctx* c, int H) { return hs_seg(c, H) + (size_t)H + 1; }

int64_t halo_scan_limit(const cc_ctx* c) {
    int64_t n_scan = c->n;
    if (c->p_n_limit >= 0) n_scan = std::max<int64_t>(0, std::min<int64_t>(c->n, c->p_n_limit - c->index_base));
    return n_scan;
}

int prepare_halo(cc_ctx* c) {
    const int H = (int)c->p_halos.size();
    if (c->n > 0xffffffffll) return fail(CC_ERR_ARG, "step too large for one shard (%lld particles)", (long long)c->n);
    int rc;
    if (hstate_elems(H) * 8 > c->hstate.bytes) c->hstate_clean = false;
    if ((rc = c->hstate.ensure(hstate_elems(H) * 8))) return rc;
    if (c->rec_cap == 0) c->rec_cap = 1 << 20;
    if ((rc = c->rec_a.ensure((size_t)c->rec_cap * sizeof(cck::Rec)))) return rc;
    if ((rc = c->rec_b.ensure((size_t)c->rec_cap * sizeof(cck::Rec)))) return rc;
    if ((rc = c->rank_buf.ensure((size_t)c->rec_cap * 4))) return rc;
    if ((rc = c->place_buf.ensure((size_t)c->rec_cap * 4))) return rc;
    if ((rc = ensure_out_cap(c, c->rec_cap))) return rc;
    if ((rc = ensure_small(c, small_elems(H, c->nranks)))) return rc;

    /* many halos: (cos theta, sin phi) cell index per launch batch (kernels.cuh, k2c_sift / k2c_match / k2c_exact) */
    c->use_bins = H >= 8 && !getenv("CC_NO_BINS");
    /* repeated cutouts with a known halo set (and the same kernel variant) reuse its device-resident parameters and
     * cell index: nothing is recomputed on the host */
    const size_t key_bytes = c->p_halos.size() * sizeof(cc_halo);
    cc_ctx::HaloSet* hs = nullptr;
    cc_ctx::HaloSet* lru = &c->halo_sets[0];
    for (auto& e : c->halo_sets) {
        if (e.key.size() == c->p_halos.size() && e.bins == c->use_bins && !e.key.empty() &&
            memcmp(e.key.data(), c->p_halos.data(), key_bytes) == 0) { hs = &e; break; }
        if (e.last_use < lru->last_use) lru = &e;
    }
    const bool known = hs != nullptr;
    if (!known) hs = lru;
    hs->last_use = ++c->halo_set_clock;
    c->hs = hs;
    if (!known) {
        CU(cudaStreamSynchronize(c->stream)); /* kernels of an earlier cutout may still read the set being replaced */
        hs->key.clear();
        if ((rc = hs->pre_buf.ensure((size_t)H * sizeof(float4)))) return rc;
        if ((rc = hs->halo_buf.ensure((size_t)H * sizeof(cck::HaloDev)))) return rc;
        /* host: exact + pre-filter parameters */
        std::vector<float4> pre(H);
        std::vector<cck::HaloDev> hd(H);
        for (int h = 0; h < H; ++h) {
            const cc_halo& s = c->p_halos[h];
            float p4[4];
            ccpf::uc2_bounds(s.theta_rough, s.phi_rough, p4);
            pre[h] = make_float4(p4[0], p4[1], p4[2], p4[3]);
            cck::HaloDev& d = hd[h];
            memset(&d, 0, sizeof(d));
            for (int i = 0; i < 9; ++i) d.R[i] = s.R[i];
            d.th_lo = s.theta_cut[0]; d.th_hi = s.theta_cut[1]; d.ph_lo = s.phi_cut[0]; d.ph_hi = s.phi_cut[1];
            d.rth_lo = s.theta_rough[0]; d.rth_hi = s.theta_rough[1]; d.rph_lo = s.phi_rough[0]; d.rph_hi = s.phi_rough[1];
            ccpf::uc2_bounds_cells_inner(s.theta_rough, s.phi_rough, p4); /* second filter of the cell-index kernel */
            d.in_chi = p4[0]; d.in_clo = p4[1]; d.in_slo = p4[2]; d.in_shi = p4[3];
            ccpf::uc2_bounds(s.theta_cut, s.phi_cut, p4);
            d.f_chi = p4[0]; d.f_clo = p4[1]; d.f_tlo = p4[2]; d.f_thi = p4[3];
        }
        if (c->use_bins) {
            const int G = cck::K2_GRID;
            const size_t ncell = (size_t)G * G;
            const int nbatch = (H + cck::K2_MAX_HALOS - 1) / cck::K2_MAX_HALOS;
            for (int h = 0; h < H; ++h) { /* the cell kernel tests sin(phi), not tan(phi) */
                float p4[4];
                ccpf::uc2_bounds_cells(c->p_halos[h].theta_rough, c->p_halos[h].phi_rough, p4);
                pre[h] = make_float4(p4[0], p4[1], p4[2], p4[3]);
            }
            /* rasterise every halo's window, then a counting sort into CSR lists */
            struct Box { int c0, c1, s0, s1; };
            auto cell_of = [G](double v) { return std::floor((v + 1.0) * (0.5 * G)); };
            std::vector<unsigned int> starts(ncell + 1);
            std::vector<uint2> ranges((size_t)nbatch * ncell);
            std::vector<unsigned int> bits((size_t)nbatch * cck::K2_BITWORDS, 0u);
            std::vector<unsigned short> items;
            hs->bin_items_off.assign((size_t)nbatch, 0);
            std::vector<uint2> pitems; /* particle-index walker: one item per (halo, grid line), then the out-of-range bin */
            hs->pidx_items_off.assign((size_t)nbatch, 0);
            hs->pidx_items_cnt.assign((size_t)nbatch, 0);
            for (int bi = 0; bi < nbatch; ++bi) {
                const int h0 = bi * cck::K2_MAX_HALOS, hn = std::min(cck::K2_MAX_HALOS, H - h0);
                unsigned int* st = starts.data();
                std::fill(starts.begin(), starts.end(), 0u);
                std::vector<Box> boxes((size_t)hn);
                for (int h = 0; h < hn; ++h) {
                    const float4 P = pre[h0 + h];
                    Box bx = {0, -1, 0, -1}; /* empty */
                    if (P.y <= P.x && P.z <= P.w) {
                        auto clampi = [G](double v) { return (int)std::max(0.0, std::min((double)G - 1, v)); };
                        bx.c0 = clampi(cell_of(std::max<double>(P.y, -1.0)) - 1.0);
                        bx.c1 = clampi(cell_of(std::min<double>(P.x, 1.0)) + 1.0);
                        bx.s0 = clampi(cell_of(std::max<double>(P.z, -1.0)) - 1.0);
                        bx.s1 = clampi(cell_of(std::min<double>(P.w, 1.0)) + 1.0);
                    }
                    boxes[(size_t)h] = bx;
                    for (int ic = bx.c0; ic <= bx.c1; ++ic)
                        for (int is = bx.s0; is <= bx.s1; ++is) st[(size_t)ic * G + is + 1] += 1;
                }
                hs->pidx_items_off[(size_t)bi] = pitems.size();
                for (int h = 0; h < hn; ++h) {
                    const Box& bx = boxes[(size_t)h];
                    if (bx.s1 < bx.s0) continue;
                    /* runs of at most 4 cells (~500 rows, 15 steps of 32): fine enough for the atomic hand-out to balance */
                    int run = 4; /* (longer runs for a window that covers a large part of the sky: at most 4096 items per halo) */
                    while ((int64_t)(bx.c1 - bx.c0 + 1) * ((bx.s1 - bx.s0 + run) / run) > 4096) run *= 2;
                    for (int ic = bx.c0; ic <= bx.c1; ++ic)
                        for (int s0 = bx.s0; s0 <= bx.s1; s0 += run)
                            pitems.push_back(make_uint2((unsigned int)h | ((unsigned int)ic << 16),
                                                        (unsigned int)s0 | ((unsigned int)std::min(bx.s1, s0 + run - 1) << 16)));
                }
                /* widest runs first: the hand-out ends with the short items (longest-processing-time-first) */
                std::stable_sort(pitems.begin() + (ptrdiff_t)hs->pidx_items_off[(size_t)bi], pitems.end(), [](const uint2& u, const uint2& v) {
                    return ((u.y >> 16) - (u.y & 0xffffu)) > ((v.y >> 16) - (v.y & 0xffffu));
                });
                pitems.push_back(make_uint2(cck::K2P_ALL, 0u));
                hs->pidx_items_cnt[(size_t)bi] = (unsigned int)(pitems.size() - hs->pidx_items_off[(size_t)bi]);
                for (size_t i = 0; i < ncell; ++i) st[i + 1] += st[i]; /* st[i] = first item of cell i */
                const size_t base_item = items.size();
                hs->bin_items_off[(size_t)bi] = base_item;
                items.resize(base_item + st[ncell]);
                std::vector<unsigned int> fill(st, st + ncell);
                for (int h = 0; h < hn; ++h) {
                    const Box& bx = boxes[(size_t)h];
                    for (int ic = bx.c0; ic <= bx.c1; ++ic)
                        for (int is = bx.s0; is <= bx.s1; ++is) items[base_item + fill[(size_t)ic * G + is]++] = (unsigned short)h;
                }
                uint2* rg = ranges.data() + (size_t)bi * ncell;
                unsigned int* bw = bits.data() + (size_t)bi * cck::K2_BITWORDS;
                for (int ic = 0; ic < G; ++ic)
                    for (int is = 0; is < G; ++is) {
                        const size_t cell = (size_t)ic * G + is;
                        rg[cell] = make_uint2(st[cell], st[cell + 1]);
                        if (st[cell + 1] > st[cell]) {
                            const unsigned int bit = (unsigned int)(ic >> cck::K2_CSHIFT) * cck::K2_COARSE + (unsigned int)(is >> cck::K2_CSHIFT);
                            bw[bit >> 5] |= 1u << (bit & 31u);
                        }
                    }
            }
            if ((rc = hs->bin_start_buf.ensure(ranges.size() * sizeof(uint2)))) return rc;
            if ((rc = hs->bin_bits_buf.ensure(bits.size() * 4))) return rc;
            if ((rc = hs->bin_items_buf.ensure(std::max<size_t>(items.size(), 8) * 2))) return rc;
            CU(cudaMemcpyAsync(hs->bin_start_buf.p, ranges.data(), ranges.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(hs->bin_bits_buf.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, c->stream));
            if (!items.empty())
                CU(cudaMemcpyAsync(hs->bin_items_buf.p, items.data(), items.size() * 2, cudaMemcpyHostToDevice, c->stream));
            if ((rc = hs->pidx_items_buf.ensure(pitems.size() * sizeof(uint2)))) return rc;
            CU(cudaMemcpyAsync(hs->pidx_items_buf.p, pitems.data(), pitems.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
        }
        CU(cudaMemcpyAsync(hs->pre_buf.p, pre.data(), (size_t)H * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(hs->halo_buf.p, hd.data(), (size_t)H * sizeof(cck::HaloDev), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream)); /* the host vectors are pageable and die here */
        hs->key = c->p_halos;
        hs->bins = c->use_bins;
        hs->walker = 0;
    }
    if (!(c->hstate_clean && c->hstate_clean_H == H)) CU(cudaMemsetAsync(c->hstate.p, 0, hstate_elems(H) * 8, c->stream));
    c->hstate_clean = false;
    return CC_OK;
}

/* where the exchanged counts land in the pinned read-back area: behind the cutout's own results */
size_t xchg_host_off(int H) { return 32 + 3 * (size_t)H; }

int fill_xchg_args(cc_ctx* c, int H, const void* send, int uc1, cck::XchgArgs* x) {
    const int R = c->nranks;
    int rc;
    if ((rc = c->gather_buf.ensure((size_t)R * 2 * H * 8 + 64))) return rc;
    if ((rc = c->offsets_buf.ensure((size_t)H * 8))) return rc;
    unsigned long long* hg = c->h_small + xchg_host_off(H);
    memset(x, 0, sizeof(*x));
    x->send = (const unsigned long long*)send;
    for (int r = 0; r < R; ++r) x->peer[r] = c->peer_mb[r];
    x->gathered = (unsigned long long*)c->gather_buf.p;
    x->offsets = (unsigned long long*)c->offsets_buf.p;
    x->status = (unsigned long long*)((char*)c->gather_buf.p + (size_t)R * 2 * H * 8);
    x->host_gathered = hg;                        /* results also land in pinned host memory: no copy commands */
    x->host_offsets = hg + (size_t)R * 2 * H + 1;
    x->uc1 = uc1;
    x->epoch = ++c->xchg_epoch;
    const char* to = getenv("CC_PEER_TIMEOUT_MS");
    x->timeout_ns = (unsigned long long)(to ? atoll(to) : 10000) * 1000000ull;
    x->n = 2 * H; x->H = H; x->R = R; x->rank = c->rank;
    return CC_OK;
}

int launch_halo_chunk(cc_ctx* c, const Chunk& ch) {
    const int H = (int)c->p_halos.size();
    unsigned long long* rec_count = hs_misc(c);
    unsigned long long* n_cand = rec_count + 1;
    if (ch.first && c->timing) CU(cudaEventRecord(c->ev_scan0, c->stream));
    /* the n_limit clamp (tail drop of the reference's load balancer) applies to step-local indices */
    const int64_t n_scan = std::max<int64_t>(0, std::min<int64_t>(ch.n, halo_scan_limit(c) - ch.start));
    const bool vec = aligned16(ch.x) && aligned16(ch.y) && aligned16(ch.z);
    for (int h0 = 0; h0 < H && n_scan > 0; h0 += cck::K2_MAX_HALOS) {
        cck::HaloArgs a;
        a.x = ch.x; a.y = ch.y; a.z = ch.z;
        a.n = n_scan;
        a.local_base = ch.start;
        a.nhalos = std::min(cck::K2_MAX_HALOS, H - h0);
        a.halo0 = h0;
        a.pre = (const float4*)c->hs->pre_buf.p + h0;
        a.halos = (const cck::HaloDev*)c->hs->halo_buf.p + h0;
        a.recs = (cck::Rec*)c->rec_a.p;
        a.rec_cap = (unsigned long long)c->rec_cap;
        a.rec_count = rec_count;
        a.counts = hs_counts(c);
        a.n_cand = n_cand;
        a.nan_count = rec_count + 5;
        a.total_halos = H;
        a.ntiles = (unsigned int)((n_scan + cck::SCAN_TILE - 1) / cck::SCAN_TILE);
        a.prefetch = c->prefetch != 0 && !c->streamed;
        const size_t pre_bytes = (size_t)a.nhalos * sizeof(float4);
        int& occ = c->k2_occupancy[vec ? 1 : 0];
        if (occ == 0) {
            const void* fn = vec ? (const void*)cck::k2_cutout_halo<true> : (const void*)cck::k2_cutout_halo<false>;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, cck::SCAN_THREADS, (size_t)cck::K2_MAX_HALOS * sizeof(float4)));
            if (occ < 1) occ = 1;
        }
        const unsigned int grid = std::max(1u, std::min((a.ntiles + cck::SCAN_WARPS - 1) / cck::SCAN_WARPS, (unsigned int)(c->sm_count * occ)));
        const bool indexed = c->use_bins && c->pidx_valid && !c->streamed && c->index_policy != 2;
        const bool mono = !indexed && c->use_bins && (c->cells_mode == 2 || (c->cells_mode == 0 && c->hs->walker == 2));
        if (c->use_bins) { c->used_mono = mono; c->used_index = indexed; }
        if (indexed) { /* particle index: only the cells under the halos' windows are read (kernels.cuh k2p_*) */
            const int bi = h0 / cck::K2_MAX_HALOS;
            cck::PidxArgs pa;
            pa.h = a;
            pa.sorted = (const float4*)c->pidx_sorted.p;
            pa.start = (const unsigned int*)c->pidx_start.p;
            pa.items = (const uint2*)c->hs->pidx_items_buf.p + c->hs->pidx_items_off[(size_t)bi];
            pa.nitems = c->hs->pidx_items_cnt[(size_t)bi];
            /* items are handed out through one zeroed counter per launch batch */
            const int nbatch = (H + cck::K2_MAX_HALOS - 1) / cck::K2_MAX_HALOS;
            int rc2;
            if ((rc2 = c->pidx_work.ensure((size_t)nbatch * 4))) return rc2;
            if (bi == 0) CU(cudaMemsetAsync(c->pidx_work.p, 0, (size_t)nbatch * 4, c->stream));
            pa.work = (unsigned int*)c->pidx_work.p + bi;
            pa.tested = rec_count + 6;
            cck::k2p_query<<<(unsigned int)c->sm_count * 5u, cck::K2P_THREADS, 0, c->stream>>>(pa);
            CU(cudaGetLastError());
            c->launches += 1;
        } else if (mono) { /* cell index, one kernel: one 20-warp CTA per SM, bitmap + windows + stacks + counters in shared memory */
            constexpr int W = 20;
            const size_t csmem = cck::k2c_smem(W) + (size_t)2 * cck::K2_MAX_HALOS * 4;
            cck::HaloBinArgs b;
            b.h = a;
            const int bi = h0 / cck::K2_MAX_HALOS;
            b.cell_range = (const uint2*)c->hs->bin_start_buf.p + (size_t)bi * ((size_t)cck::K2_GRID * cck::K2_GRID);
            b.cell_bits = (const unsigned int*)c->hs->bin_bits_buf.p + (size_t)bi * cck::K2_BITWORDS;
            b.cell_items = (const unsigned short*)c->hs->bin_items_buf.p + c->hs->bin_items_off[(size_t)bi];
            if (!c->cells_mono_ready) {
                CU(cudaFuncSetAttribute((const void*)cck::k2_cutout_halo_cells<true, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
                CU(cudaFuncSetAttribute((const void*)cck::k2_cutout_halo_cells<false, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
                c->cells_mono_ready = true;
            }
            const unsigned int ctiles = (unsigned int)((n_scan + cck::K2C_TILE - 1) / cck::K2C_TILE);
            const unsigned int cgrid = std::max(1u, std::min((ctiles + W - 1) / W, (unsigned int)c->sm_count));
            if (vec) cck::k2_cutout_halo_cells<true, W><<<cgrid, W * 32, csmem, c->stream>>>(b);
            else cck::k2_cutout_halo_cells<false, W><<<cgrid, W * 32, csmem, c->stream>>>(b);
            CU(cudaGetLastError());
            c->launches += 1;
        } else if (c->use_bins) { /* cell index, staged: sift -> match -> exact (kernels.cuh) */
            const int bi = h0 / cck::K2_MAX_HALOS;
            int& socc = c->sift_occupancy[vec ? 1 : 0];
            if (socc == 0) {
                const void* fn = vec ? (const void*)cck::k2c_sift<true> : (const void*)cck::k2c_sift<false>;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&socc, fn, cck::SCAN_THREADS, 0));
                if (socc < 1) socc = 1;
            }
            const unsigned int sgrid_max = (unsigned int)(c->sm_count * socc);
            const unsigned int sgrid = std::max(1u, std::min((a.ntiles + cck::SCAN_WARPS - 1) / cck::SCAN_WARPS, sgrid_max));
            const unsigned int warps_max = sgrid_max * cck::SCAN_WARPS, nregions = sgrid * cck::SCAN_WARPS;
            if (c->cand_region_cap == 0) c->cand_region_cap = (unsigned int)std::min<int64_t>(0x7ffffff0, c->n / 6 / warps_max + 512);
            if (c->pair_cap == 0) c->pair_cap = 4 << 20;
            int rc2;
            if ((rc2 = c->cand.ensure((size_t)warps_max * c->cand_region_cap * sizeof(float4)))) return rc2;
            if ((rc2 = c->cand_count.ensure((size_t)warps_max * 4))) return rc2;
            if ((rc2 = c->pair_p.ensure((size_t)c->pair_cap * sizeof(float4)))) return rc2;
            if ((rc2 = c->pair_h.ensure((size_t)c->pair_cap * 4))) return rc2;
            if ((rc2 = c->pair_count.ensure(64))) return rc2;
            CU(cudaMemsetAsync(c->pair_count.p, 0, 8, c->stream));
            cck::SiftArgs sa;
            sa.x = a.x; sa.y = a.y; sa.z = a.z; sa.n = a.n; sa.local_base = a.local_base;
            sa.cell_bits = (const unsigned int*)c->hs->bin_bits_buf.p + (size_t)bi * cck::K2_BITWORDS;
            sa.cand = (float4*)c->cand.p; sa.region_cap = c->cand_region_cap; sa.cand_count = (unsigned int*)c->cand_count.p;
            sa.misc = hs_misc(c); sa.ntiles = a.ntiles; sa.prefetch = a.prefetch;
            if (vec) cck::k2c_sift<true><<<sgrid, cck::SCAN_THREADS, 0, c->stream>>>(sa);
            else cck::k2c_sift<false><<<sgrid, cck::SCAN_THREADS, 0, c->stream>>>(sa);
            CU(cudaGetLastError());
            cck::MatchArgs ma;
            ma.h = a;
            ma.cell_range = (const uint2*)c->hs->bin_start_buf.p + (size_t)bi * ((size_t)cck::K2_GRID * cck::K2_GRID);
            ma.cell_items = (const unsigned short*)c->hs->bin_items_buf.p + c->hs->bin_items_off[(size_t)bi];
            ma.cand = sa.cand; ma.cand_count = sa.cand_count; ma.region_cap = sa.region_cap; ma.nregions = nregions;
            ma.pair_p = (float4*)c->pair_p.p; ma.pair_h = (unsigned int*)c->pair_h.p;
            ma.pair_cap = (unsigned long long)c->pair_cap; ma.pair_count = (unsigned long long*)c->pair_count.p;
            cck::k2c_match<<<std::min(nregions, (unsigned int)c->sm_count * 8u), cck::K2M_THREADS, 0, c->stream>>>(ma);
            CU(cudaGetLastError());
            cck::ExactArgs ea;
            ea.h = a;
            ea.pair_p = ma.pair_p; ea.pair_h = ma.pair_h; ea.pair_cap = ma.pair_cap; ea.pair_count = ma.pair_count;
            ea.misc = hs_misc(c);
            cck::k2c_exact<<<(unsigned int)c->sm_count * 2u, cck::K2X_THREADS, 0, c->stream>>>(ea);
            CU(cudaGetLastError());
            c->launches += 3;
        } else if (c->k2_tma && vec) { /* bulk-copy fed variant (CC_K2_TMA=1) */
            const size_t smem = cck::K2T_SMEM_BASE + pre_bytes;
            if (!c->k2t_ready) {
                CU(cudaFuncSetAttribute((const void*)cck::k2_cutout_halo_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(cck::K2T_SMEM_BASE + cck::K2_MAX_HALOS * sizeof(float4))));
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->k2t_occ, (const void*)cck::k2_cutout_halo_tma, cck::SCAN_THREADS,
                                                                  cck::K2T_SMEM_BASE + 64 * sizeof(float4)));
                if (c->k2t_occ < 1) c->k2t_occ = 1;
                c->k2t_ready = true;
            }
            const unsigned int gridt = std::max(1u, std::min((a.ntiles + cck::SCAN_WARPS - 1) / cck::SCAN_WARPS, (unsigned int)(c->sm_count * c->k2t_occ)));
            cck::k2_cutout_halo_tma<<<gridt, cck::SCAN_THREADS, smem, c->stream>>>(a);
            CU(cudaGetLastError());
            c->launches += 1;
        } else {
            if (vec) cck::k2_cutout_halo<true><<<grid, cck::SCAN_THREADS, pre_bytes, c->stream>>>(a);
            else cck::k2_cutout_halo<false><<<grid, cck::SCAN_THREADS, pre_bytes, c->stream>>>(a);
            CU(cudaGetLastError());
            c->launches += 1;
        }
    }
    if (ch.last) {
        if (c->timing) CU(cudaEventRecord(c->ev_scan1, c->stream));
        /* ordering + gather, sized on the device: the whole pass is one enqueue, one host synchronisation */
        cck::GatherArgs g;
        g.rank = (const unsigned int*)c->rank_buf.p;
        g.place = nullptr;
        g.seg_begin = hs_seg(c, H);
        g.rec_count = rec_count;
        g.rec_cap = (unsigned long long)c->rec_cap;
        g.out_cap = c->out_cap;
        g.in = c->gcols;
        if (c->streamed && c->host_gather) g.in = cck::ColsIn{}; /* the host picks these columns itself (fetch_rows) */
        g.payload = (c->streamed || !c->payload_valid) ? nullptr : (const cck::Payload*)c->payload.p;
        g.out = cols_out_of(c);
        g.index_base = c->index_base;
        g.pos_only = c->p_pos_only;
        /* few survivors last time: speculate that one single-CTA kernel can order and gather everything */
        const bool small = !c->force_general && c->last_nrec >= 0 && c->last_nrec <= cck::K3_SORT_MAX * 3 / 4 && H <= 65535;
        c->used_small = small;
        if (small) {
            if (!c->k3s_smem_set) {
                CU(cudaFuncSetAttribute((const void*)cck::k3_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cck::K3S_SMEM_BYTES));
                c->k3s_smem_set = true;
            }
            cck::SmallArgs sa;
            sa.g = g;
            sa.g.recs = (const cck::Rec*)c->rec_a.p;
            sa.counts = hs_counts(c);
            sa.seg_begin = hs_seg(c, H);
            sa.misc = hs_misc(c);
            sa.H = H;
            /* the last kernel of the pass hands the results to the host itself (pinned memory, no copy command behind
             * it) and leaves the device state zeroed for the next cutout (no memset in front of the next scan) */
            sa.host_out = c->h_small;
            sa.seq = ++c->small_seq;
            sa.state_elems = (unsigned int)hstate_elems(H);
            { int rc2 = c->xchg_send.ensure((size_t)2 * H * 8); if (rc2) return rc2; }
            sa.xchg_send = (unsigned long long*)c->xchg_send.p;
            /* eager exchange: this kernel posts the counts to the peers first and collects theirs last */
            sa.do_xchg = 0;
            if (c->xchg_mode == 3 && !c->eager_done && c->peers_ready && c->nranks > 1 && 2 * H <= cck::XCHG_PAYLOAD &&
                c->h_small_elems >= small_elems(H, c->nranks)) {
                int rc2 = fill_xchg_args(c, H, hs_counts(c), 0, &sa.x);
                if (rc2) return rc2;
                sa.do_xchg = 1;
                c->eager_done = true;
            }
            cck::k3_small<<<1, cck::K3S_THREADS, cck::K3S_SMEM_BYTES, c->stream>>>(sa);
            CU(cudaGetLastError());
            c->launches += 1;
            c->hstate_clean = true;
            c->hstate_clean_H = H;
            c->xchg_src = c->xchg_send.p;
            CU(cudaEventRecord(c->ev_counts, c->stream)); /* the exchange reads the saved counts: after this kernel */
        } else {
            c->xchg_src = hs_counts(c);
            CU(cudaEventRecord(c->ev_counts, c->stream)); /* counts are final: the exchange runs next to ordering + gather */
            cck::k3_scan_counts<<<1, 256, 0, c->stream>>>(hs_counts(c), hs_seg(c, H), H);
            CU(cudaGetLastError());
            const unsigned int g1 = (unsigned int)c->sm_count * 8; /* 2048 threads per SM: every record is a chain of dependent L2 round trips */
            cck::BucketArgs b;
            b.recs = (const cck::Rec*)c->rec_a.p;
            b.rec_count = rec_count;
            b.rec_cap = (unsigned long long)c->rec_cap;
            b.seg_begin = hs_seg(c, H);
            b.cursor = hs_cursor(c, H);
            b.out = (cck::Rec*)c->rec_b.p;
            b.H = H;
            cck::k3_bucket<<<(unsigned int)c->sm_count * 4u, cck::K3B_THREADS, 0, c->stream>>>(b); /* 4 CTAs x 512 threads x 32 KB per SM */
            CU(cudaGetLastError());
            cck::RankArgs r;
            r.rec_count = rec_count;
            r.recs = (const cck::Rec*)c->rec_b.p;
            r.seg_begin = hs_seg(c, H);
            r.rec_cap = (unsigned long long)c->rec_cap;
            r.rank = (unsigned int*)c->rank_buf.p;
            r.sorted = (unsigned long long*)c->rec_a.p; /* the unbucketed records are dead once k3_bucket ran */
            r.place = (unsigned int*)c->place_buf.p;
            r.H = H;
            /* few halos: many chunk slots per halo; many halos: a few, the y dimension provides the parallelism */
            const unsigned int gy = (unsigned int)std::min(H, 8192);
            /* (a CTA holds 196 KB of shared memory, i.e. an SM: idle chunk slots of a many-halo pass cost as much as busy ones) */
            const unsigned int gx = (unsigned int)std::max(1, std::min(64, (c->sm_count * 2) / (int)gy));
            if (!c->k3_smem_set) {
                CU(cudaFuncSetAttribute((const void*)cck::k3_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cck::K3_SMEM_BYTES));
                c->k3_smem_set = true;
            }
            cck::k3_rank<<<dim3(gx, gy), cck::K3_THREADS, cck::K3_SMEM_BYTES, c->stream>>>(r);
            CU(cudaGetLastError());
            cck::PlaceArgs pl;
            pl.recs = (const cck::Rec*)c->rec_b.p;
            pl.rank = (const unsigned int*)c->rank_buf.p;
            pl.sorted = (const unsigned long long*)c->rec_a.p;
            pl.seg_begin = hs_seg(c, H);
            pl.rec_count = rec_count;
            pl.rec_cap = (unsigned long long)c->rec_cap;
            pl.place = (unsigned int*)c->place_buf.p;
            pl.H = H;
            {
                /* few halos: many slices per segment; many halos: the y dimension provides the parallelism */
                const unsigned int py = (unsigned int)std::min(H, 4096);
                const unsigned int px = (unsigned int)std::max(2, std::min(1024, (c->sm_count * 8) / (int)py));
                cck::k3_place<<<dim3(px, py), 256, 0, c->stream>>>(pl);
            }
            CU(cudaGetLastError());
            g.recs = (const cck::Rec*)c->rec_b.p;
            g.place = (const unsigned int*)c->place_buf.p;
            cck::k3_gather_halo<<<g1, 256, 0, c->stream>>>(g);
            CU(cudaGetLastError());
            c->launches += 5;
        }
        if (c->timing) CU(cudaEventRecord(c->ev_total1, c->stream));
        c->have_times = c->timing;
        /* misc {rec_count, n_cand, ...}, counts, rough counts, segment starts: one read-back (k3_small wrote it itself) */
        if (!small) CU(cudaMemcpyAsync(c->h_small, c->hstate.p, (8 + 3 * (size_t)H + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    return CC_OK;
}

/* ---------------------------------------------------------------- whole-step drivers */
int maybe_build_payload(cc_ctx* c);
int maybe_build_index(cc_ctx* c);

int run_resident(cc_ctx* c) {
    int rc = (c->pending == P_CONST) ? prepare_const(c) : prepare_halo(c);
    if (rc) return rc;
    if ((rc = maybe_build_payload(c))) return rc;
    if ((rc = maybe_build_index(c))) return rc;
    c->gcols = cols_in_of(c);
    Chunk ch;
    ch.x = c->gcols.x; ch.y = c->gcols.y; ch.z = c->gcols.z;
    ch.n = c->n; ch.start = 0; ch.first = true; ch.last = true;
    return (c->pending == P_CONST) ? launch_const_chunk(c, ch) : launch_halo_chunk(c, ch);
}

/* pinned host -> HBM, double-buffered: the copy stream fills slot s while the cutout kernel consumes slot s^1 */
int run_streamed(cc_ctx* c) {
    int rc = (c->pending == P_CONST) ? prepare_const(c) : prepare_halo(c);
    if (rc) return rc;
    const int64_t n = c->n;
    const int64_t chunk = std::max<int64_t>(cck::SCAN_TILE, (c->stream_chunk / cck::SCAN_TILE) * cck::SCAN_TILE);
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int s = 0; s < 2; ++s) {
            CU(cudaEventCreateWithFlags(&c->ev_copied[s], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_consumed[s], cudaEventDisableTiming));
        }
    }
    const int64_t slot_elems = std::min<int64_t>(chunk, std::max<int64_t>(n, 1));
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < 3; ++k)
            if ((rc = c->stage[s][k].ensure((size_t)slot_elems * 4))) return rc;
    /* the copy stream must not overwrite a slot that kernels of an earlier call may still read */
    CU(cudaEventRecord(c->ev_consumed[0], c->stream));
    CU(cudaEventRecord(c->ev_consumed[1], c->stream));
    const int64_t nchunks = std::max<int64_t>(1, (n + chunk - 1) / chunk);
    for (int64_t i = 0; i < nchunks; ++i) {
        const int s = (int)(i & 1);
        const int64_t start = i * chunk, len = std::min<int64_t>(chunk, n - start);
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[s], 0));
        const float* src[3] = {c->host_cols.x, c->host_cols.y, c->host_cols.z};
        for (int k = 0; k < 3 && len > 0; ++k)
            CU(cudaMemcpyAsync(c->stage[s][k].p, src[k] + start, (size_t)len * 4, cudaMemcpyHostToDevice, c->copy_stream));
        c->bytes_uploaded += 12 * len;
        CU(cudaEventRecord(c->ev_copied[s], c->copy_stream));
        CU(cudaStreamWaitEvent(c->stream, c->ev_copied[s], 0));
        Chunk ch;
        ch.x = (const float*)c->stage[s][0].p; ch.y = (const float*)c->stage[s][1].p; ch.z = (const float*)c->stage[s][2].p;
        ch.n = len; ch.start = start; ch.first = (i == 0); ch.last = (i == nchunks - 1);
        rc = (c->pending == P_CONST) ? launch_const_chunk(c, ch) : launch_halo_chunk(c, ch);
        if (rc) return rc;
        CU(cudaEventRecord(c->ev_consumed[s], c->stream));
    }
    return CC_OK;
}

int run_pending(cc_ctx* c) { return c->streamed ? run_streamed(c) : run_resident(c); }

/* Interleave the seven gather-only columns of the resident step (kernels.cuh, "gather payload").
 * Costs one pass over 64 B/particle (1.46 ms per 100 M particles) and 32 B/particle of HBM, so it is built LAZILY, inside
 * the cutout call that first profits from it (its time lands in that cutout): policy "auto" builds it when the previous
 * cutout on this resident step left at least max(65536, n/512) survivors -- a sparse cutout (a cluster-sized halo: a
 * few hundred survivors) never pays for it and gathers from the caller's columns instead; "always"/"never" force it
 * (cc_set_option "payload", or CC_PAYLOAD / CC_NO_PAYLOAD in the environment). */
int maybe_build_payload(cc_ctx* c) {
    if (c->payload_valid || c->streamed || c->n == 0 || c->payload_policy == 2) return CC_OK;
    if (c->payload_policy == 0) {
        const int64_t need = std::max<int64_t>(65536, c->n / 512);
        if (c->step_last_survivors < need) return CC_OK;
    }
    if (c->payload.bytes < (size_t)c->n * sizeof(cck::Payload)) {
        c->payload.release();
        const cudaError_t e = cudaMalloc(&c->payload.p, (size_t)c->n * sizeof(cck::Payload));
        if (e != cudaSuccess) { /* not enough HBM for the extra 32 B/particle: gather from the columns instead */
            cudaGetLastError();
            c->payload.p = nullptr;
            c->payload_policy = 2;
            return CC_OK;
        }
        c->payload.bytes = (size_t)c->n * sizeof(cck::Payload);
    }
    const cck::ColsIn ci = cols_in_of(c);
    const int with_vel = (ci.vx && ci.vy && ci.vz && ci.rot && ci.rep) ? 1 : 0;
    CU(cudaEventRecord(c->ev_pay0, c->stream));
    cck::k_build_payload<<<c->sm_count * 8, 256, 0, c->stream>>>(ci, (cck::Payload*)c->payload.p, c->n, with_vel);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_pay1, c->stream));
    c->launches += 1;
    c->payload_valid = true;
    c->payload_timed = true;
    return CC_OK;
}

/* the particle index of the resident step (kernels.cuh k2p_*): counting sort by (cos theta_u, sin phi_u) cell */
int maybe_build_index(cc_ctx* c) {
    if (c->pending != P_HALO || !c->use_bins || c->streamed || c->pidx_valid || c->n == 0 || c->index_policy == 2) return CC_OK;
    if (c->index_policy == 0 && (c->step_many_halo_cutouts < 1 || c->n < (4ll << 20))) return CC_OK;
    const size_t nbin = cck::K2P_NBIN;
    const unsigned int nblocks = (unsigned int)((nbin + cck::K2P_SCAN_BLOCK - 1) / cck::K2P_SCAN_BLOCK);
    {
        /* not enough HBM for the extra 16 B/particle: keep scanning */
        DevBuf* need[] = {&c->pidx_sorted, &c->pidx_start, &c->pidx_tmp, &c->pidx_partial};
        const size_t bytes[] = {(size_t)c->n * sizeof(float4), (nbin + 1) * 4, (nbin + 1) * 4, (size_t)nblocks * 4};
        for (int i = 0; i < 4; ++i) {
            if (need[i]->bytes >= bytes[i]) continue;
            need[i]->release();
            if (cudaMalloc(&need[i]->p, bytes[i]) != cudaSuccess) {
                cudaGetLastError();
                need[i]->p = nullptr;
                c->index_policy = 2;
                return CC_OK;
            }
            need[i]->bytes = bytes[i];
        }
    }
    const cck::ColsIn ci = cols_in_of(c);
    unsigned int* hist = (unsigned int*)c->pidx_tmp.p;
    unsigned int* start = (unsigned int*)c->pidx_start.p;
    unsigned int* partial = (unsigned int*)c->pidx_partial.p;
    CU(cudaEventRecord(c->ev_idx0, c->stream));
    CU(cudaMemsetAsync(hist, 0, (nbin + 1) * 4, c->stream));
    const unsigned int grid = (unsigned int)c->sm_count * 16u;
    cck::k2p_hist<<<grid, 256, 0, c->stream>>>(ci.x, ci.y, ci.z, c->n, hist);
    CU(cudaGetLastError());
    cck::k2p_scan_a<<<nblocks, 256, 0, c->stream>>>(hist, partial);
    CU(cudaGetLastError());
    cck::k2p_scan_b<<<1, 32, 0, c->stream>>>(partial, nblocks);
    CU(cudaGetLastError());
    cck::k2p_scan_c<<<nblocks, 256, 0, c->stream>>>(hist, partial, start);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hist, start, nbin * 4, cudaMemcpyDeviceToDevice, c->stream)); /* fill cursors */
    cck::k2p_scatter<<<grid, 256, 0, c->stream>>>(ci.x, ci.y, ci.z, c->n, hist, (float4*)c->pidx_sorted.p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_idx1, c->stream));
    c->launches += 5;
    c->pidx_valid = true;
    c->pidx_gen = ++c->pidx_gen_clock;
    c->index_timed = true;
    return CC_OK;
}

/* a new step became resident */
void reset_step_state(cc_ctx* c) {
    c->payload_valid = false;
    c->step_last_survivors = -1;
    c->payload_timed = false;
    c->pidx_valid = false;
    c->step_many_halo_cutouts = 0;
    c->index_timed = false;
}

int check_sm100(int device, int* sm_count) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CC_ERR_CUDA, "device %d (%s, sm_%d%d) is not a Blackwell sm_100 GPU; this library has no other path",
                    device, prop.name, prop.major, prop.minor);
    *sm_count = prop.multiProcessorCount;
    return CC_OK;
}

}  // namespace

extern "C" {

const char* cc_version(void) { return "cosmo-cutout_b200 0.2 (sm_100a)"; }
const char* cc_last_error(void) {
    if (!g_err.empty()) return g_err.c_str();
    std::lock_guard<std::mutex> lock(g_err_any_mutex);
    g_err_copy = g_err_any;
    return g_err_copy.c_str();
}

int cc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cc_create(int device, cc_ctx** out) {
    if (!out) return fail(CC_ERR_ARG, "cc_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CC_ERR_CUDA, "no CUDA device available (%s); cosmo-cutout_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return fail(CC_ERR_ARG, "cc_create: device %d out of range [0,%d)", device, ndev);
    int sms = 0;
    int rc = check_sm100(device, &sms);
    if (rc) return rc;
    CU(cudaSetDevice(device));
    cc_ctx* c = new cc_ctx();
    if (const char* e = getenv("CC_K2_TMA")) c->k2_tma = atoi(e) != 0;
    if (const char* e = getenv("CC_GATHER")) c->gather_mode = !strcmp(e, "bulk") ? 2 : !strcmp(e, "ldst") ? 1 : 0;
    if (const char* e = getenv("CC_TIMING")) c->timing = atoi(e) != 0;
    if (const char* e = getenv("CC_PREFETCH")) c->prefetch = atoi(e);
    if (const char* e = getenv("CC_CELLS")) c->cells_mode = !strcmp(e, "staged") ? 1 : !strcmp(e, "mono") ? 2 : 0;
    if (const char* e = getenv("CC_INDEX")) c->index_policy = !strcmp(e, "always") ? 1 : !strcmp(e, "never") ? 2 : 0;
    if (getenv("CC_NO_PAYLOAD")) c->payload_policy = 2;
    if (const char* e = getenv("CC_PAYLOAD")) c->payload_policy = !strcmp(e, "always") ? 1 : !strcmp(e, "never") ? 2 : 0;
    c->device = device;
    c->sm_count = sms;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev_scan0) != cudaSuccess || cudaEventCreate(&c->ev_scan1) != cudaSuccess ||
        cudaEventCreate(&c->ev_total1) != cudaSuccess ||
        cudaEventCreate(&c->ev_pay0) != cudaSuccess || cudaEventCreate(&c->ev_pay1) != cudaSuccess ||
        cudaEventCreate(&c->ev_idx0) != cudaSuccess || cudaEventCreate(&c->ev_idx1) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_counts, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_xchg, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->xchg_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return fail(CC_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    rc = ensure_small(c, 64);
    if (rc) { cc_destroy(c); return rc; }
    *out = c;
    return CC_OK;
}

void cc_destroy(cc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->xchg_stream) cudaStreamSynchronize(c->xchg_stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (auto& sl : c->slots) sl.release();
    for (int r = 0; r < cck::XCHG_MAX_RANKS; ++r)
        if (c->peer_opened[r]) cudaIpcCloseMemHandle(c->peer_mb[r]);
    c->mailbox.release();
    for (auto& b : c->col_buf) b.release();
    c->payload.release();
    for (auto& b : c->out_buf) b.release();
    DevBuf* rest[] = {&c->pidx_sorted, &c->pidx_start, &c->pidx_tmp, &c->pidx_partial, &c->pidx_work, &c->scratch, &c->rec_a, &c->rec_b, &c->rank_buf, &c->place_buf, &c->counts_buf,
                      &c->hstate, &c->gather_buf, &c->offsets_buf, &c->xchg_send, &c->cand, &c->cand_count, &c->pair_p, &c->pair_h, &c->pair_count, &c->t_idx, &c->t_th, &c->t_ph, &c->t_x, &c->t_y, &c->t_z};
    for (DevBuf* b : rest) b->release();
    for (auto& e : c->halo_sets) e.release();
    for (auto& sl : c->stage) for (auto& b : sl) b.release();
    for (int i = 0; i < 2; ++i) { if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]); }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->h_small) cudaFreeHost(c->h_small);
    if (c->ev_scan0) cudaEventDestroy(c->ev_scan0);
    if (c->ev_scan1) cudaEventDestroy(c->ev_scan1);
    if (c->ev_counts) cudaEventDestroy(c->ev_counts);
    if (c->ev_xchg) cudaEventDestroy(c->ev_xchg);
    if (c->xchg_stream) cudaStreamDestroy(c->xchg_stream);
    if (c->ev_total1) cudaEventDestroy(c->ev_total1);
    if (c->ev_pay0) cudaEventDestroy(c->ev_pay0);
    if (c->ev_pay1) cudaEventDestroy(c->ev_pay1);
    if (c->ev_idx0) cudaEventDestroy(c->ev_idx0);
    if (c->ev_idx1) cudaEventDestroy(c->ev_idx1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* cc_stream(cc_ctx* c) { return c ? (void*)c->stream : nullptr; }

/* ------------------------------------------------------------------ step residency */
static int step_alloc(cc_ctx* c, int64_t n, bool with_vel) {
    for (int i = 0; i < N_IN_COLS; ++i) {
        const bool vel = (i == COL_VX || i == COL_VY || i == COL_VZ || i == COL_ROT || i == COL_REP);
        if (vel && !with_vel) { c->col[i] = nullptr; continue; }
        int rc = c->col_buf[i].ensure(std::max<size_t>((size_t)n * IN_COL_SIZE[i], 16));
        if (rc) return rc;
        c->col[i] = c->col_buf[i].p;
    }
    return CC_OK;
}

int cc_step_release(cc_ctx* c) {
    if (!c) return fail(CC_ERR_ARG, "ctx is NULL");
    int rc = use_device(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    for (auto& b : c->col_buf) b.release();
    c->payload.release();
    c->pidx_sorted.release();
    reset_step_state(c);
    for (auto& p : c->col) p = nullptr;
    c->step_valid = false;
    c->step_owned = false;
    c->n = 0;
    c->pending = P_NONE;
    return CC_OK;
}

int cc_step_select(cc_ctx* c, int slot) {
    if (!c || slot < 0 || slot > 65535) return fail(CC_ERR_ARG, "cc_step_select: bad argument");
    int rc = use_device(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream)); /* a pending cutout may still read the current step */
    c->pending = P_NONE;
    if (slot == c->cur_slot) return CC_OK;
    const size_t need = (size_t)std::max(slot, c->cur_slot) + 1;
    if (c->slots.size() < need) c->slots.resize(need);
    StepSlot& out = c->slots[(size_t)c->cur_slot];
    out.valid = c->step_valid; out.owned = c->step_owned; out.n = c->n; out.index_base = c->index_base;
    out.payload = c->payload; out.payload_valid = c->payload_valid; out.last_survivors = c->step_last_survivors;
    out.pidx_sorted = c->pidx_sorted; out.pidx_start = c->pidx_start; out.pidx_valid = c->pidx_valid; out.pidx_gen = c->pidx_gen;
    out.many_halo_cutouts = c->step_many_halo_cutouts;
    for (int i = 0; i < N_IN_COLS; ++i) { out.col[i] = c->col[i]; out.col_buf[i] = c->col_buf[i]; }
    StepSlot& in = c->slots[(size_t)slot];
    c->step_valid = in.valid; c->step_owned = in.owned; c->n = in.n; c->index_base = in.index_base;
    c->payload = in.payload; c->payload_valid = in.payload_valid; c->step_last_survivors = in.last_survivors;
    c->pidx_sorted = in.pidx_sorted; c->pidx_start = in.pidx_start; c->pidx_valid = in.pidx_valid; c->pidx_gen = in.pidx_gen;
    c->step_many_halo_cutouts = in.many_halo_cutouts;
    c->index_timed = false;
    for (int i = 0; i < N_IN_COLS; ++i) { c->col[i] = in.col[i]; c->col_buf[i] = in.col_buf[i]; }
    in = StepSlot(); /* the buffers now belong to the context's current-step fields */
    c->cur_slot = slot;
    c->streamed = false;
    c->payload_timed = false;
    return CC_OK;
}

int cc_device_mem(cc_ctx* c, int64_t* free_bytes, int64_t* total_bytes) {
    if (!c) return fail(CC_ERR_ARG, "ctx is NULL");
    int rc = use_device(c);
    if (rc) return rc;
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return CC_OK;
}

int64_t cc_bytes_uploaded(cc_ctx* c) { return c ? c->bytes_uploaded : 0; }

int cc_step_upload(cc_ctx* c, int64_t n, const cc_cols_in* host, int64_t index_base) {
    if (!c || !host || n < 0) return fail(CC_ERR_ARG, "cc_step_upload: bad argument");
    if (!host->x || !host->y || !host->z || !host->a || !host->id) return fail(CC_ERR_ARG, "cc_step_upload: x,y,z,a,id are required");
    const bool with_vel = host->vx && host->vy && host->vz && host->rotation && host->replication;
    int rc = use_device(c);
    if (rc) return rc;
    if ((rc = step_alloc(c, n, with_vel))) return rc;
    const void* src[N_IN_COLS] = {host->x, host->y, host->z, host->vx, host->vy, host->vz, host->a, host->id, host->rotation, host->replication};
    for (int i = 0; i < N_IN_COLS; ++i)
        if (c->col[i] && n > 0) {
            CU(cudaMemcpyAsync(const_cast<void*>(c->col[i]), src[i], (size_t)n * IN_COL_SIZE[i], cudaMemcpyHostToDevice, c->stream));
            c->bytes_uploaded += n * (int64_t)IN_COL_SIZE[i];
        }
    CU(cudaStreamSynchronize(c->stream));
    c->n = n;
    c->index_base = index_base;
    c->step_valid = true;
    c->step_owned = true;
    c->pending = P_NONE;
    reset_step_state(c);
    return CC_OK;
}

int cc_step_attach(cc_ctx* c, int64_t n, const cc_cols_in* dev, int64_t index_base) {
    if (!c || !dev || n < 0) return fail(CC_ERR_ARG, "cc_step_attach: bad argument");
    if (!dev->x || !dev->y || !dev->z || !dev->a || !dev->id) return fail(CC_ERR_ARG, "cc_step_attach: x,y,z,a,id are required");
    const void* src[N_IN_COLS] = {dev->x, dev->y, dev->z, dev->vx, dev->vy, dev->vz, dev->a, dev->id, dev->rotation, dev->replication};
    for (int i = 0; i < N_IN_COLS; ++i) c->col[i] = src[i];
    c->n = n;
    c->index_base = index_base;
    c->step_valid = true;
    c->step_owned = false;
    c->pending = P_NONE;
    reset_step_state(c);
    return use_device(c);
}

int cc_step_synth(cc_ctx* c, int64_t n, uint64_t seed, int64_t index_base, int kind, float box) {
    if (!c || n < 0) return fail(CC_ERR_ARG, "cc_step_synth: bad argument");
    int rc = use_device(c);
    if (rc) return rc;
    if ((rc = step_alloc(c, n, true))) return rc;
    cck::SynthCols s;
    s.x = (float*)c->col[COL_X]; s.y = (float*)c->col[COL_Y]; s.z = (float*)c->col[COL_Z];
    s.vx = (float*)c->col[COL_VX]; s.vy = (float*)c->col[COL_VY]; s.vz = (float*)c->col[COL_VZ];
    s.a = (float*)c->col[COL_A]; s.id = (long long*)c->col[COL_ID]; s.rot = (int*)c->col[COL_ROT]; s.rep = (int*)c->col[COL_REP];
    if (n > 0) {
        cck::k_synth<<<c->sm_count * 8, 256, 0, c->stream>>>(s, n, seed, index_base, kind, box);
        CU(cudaGetLastError());
        c->launches += 1;
    }
    CU(cudaStreamSynchronize(c->stream));
    c->n = n;
    c->index_base = index_base;
    c->step_valid = true;
    c->step_owned = true;
    c->pending = P_NONE;
    reset_step_state(c);
    return CC_OK;
}

int cc_synth_host(int64_t n, uint64_t seed, int64_t index_base, int kind, float box, float* x, float* y, float* z,
                  float* vx, float* vy, float* vz, float* a, int64_t* id, int32_t* rotation, int32_t* replication) {
    if (n < 0 || !x || !y || !z) return fail(CC_ERR_ARG, "cc_synth_host: bad argument");
    /* launchers such as torchrun pin OMP_NUM_THREADS to 1: CC_HOST_THREADS overrides */
    int nt = 0;
    if (const char* e = getenv("CC_HOST_THREADS")) nt = std::max(1, atoi(e));
    if (nt <= 0) nt = std::max(1, omp_get_max_threads());
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n; ++i) {
        float fx, fy, fz, fvx, fvy, fvz, fa;
        long long fid;
        int frot, frep;
        cck::synth_one(seed, index_base + i, kind, box, fx, fy, fz, fvx, fvy, fvz, fa, fid, frot, frep);
        x[i] = fx; y[i] = fy; z[i] = fz;
        if (vx) vx[i] = fvx;
        if (vy) vy[i] = fvy;
        if (vz) vz[i] = fvz;
        if (a) a[i] = fa;
        if (id) id[i] = fid;
        if (rotation) rotation[i] = frot;
        if (replication) replication[i] = frep;
    }
    return CC_OK;
}

int cc_step_download(cc_ctx* c, int64_t first, int64_t count, float* x, float* y, float* z, float* vx, float* vy,
                     float* vz, float* a, int64_t* id, int32_t* rotation, int32_t* replication) {
    if (!c || !c->step_valid) return fail(CC_ERR_STATE, "cc_step_download: no resident step");
    if (first < 0 || count < 0 || first + count > c->n) return fail(CC_ERR_ARG, "cc_step_download: range out of bounds");
    int rc = use_device(c);
    if (rc) return rc;
    void* dst[N_IN_COLS] = {x, y, z, vx, vy, vz, a, id, rotation, replication};
    for (int i = 0; i < N_IN_COLS; ++i)
        if (dst[i] && c->col[i] && count > 0)
            CU(cudaMemcpyAsync(dst[i], (const char*)c->col[i] + (size_t)first * IN_COL_SIZE[i], (size_t)count * IN_COL_SIZE[i],
                               cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CC_OK;
}

int64_t cc_step_size(cc_ctx* c) { return (c && c->step_valid) ? c->n : -1; }

/* ------------------------------------------------------------------ cutout */
int cc_cutout_const(cc_ctx* c, const float theta_cut[2], const float phi_cut[2]) {
    if (!c || !theta_cut || !phi_cut) return fail(CC_ERR_ARG, "cc_cutout_const: bad argument");
    if (!c->step_valid) return fail(CC_ERR_STATE, "cc_cutout_const: no resident step");
    if (!c->col[COL_VX] || !c->col[COL_VY] || !c->col[COL_VZ] || !c->col[COL_ROT] || !c->col[COL_REP])
        return fail(CC_ERR_STATE, "cc_cutout_const: use case 1 writes all 12 columns (processLC.cpp:291-307); the step lacks vx/vy/vz/rotation/replication");
    int rc = use_device(c);
    if (rc) return rc;
    c->p_th[0] = theta_cut[0]; c->p_th[1] = theta_cut[1];
    c->p_ph[0] = phi_cut[0]; c->p_ph[1] = phi_cut[1];
    c->nhalos = 1;
    c->pending = P_CONST;
    c->synced = false;
    c->streamed = false;
    c->payload_timed = false;
    c->index_timed = false;
    c->eager_done = false;
    return run_pending(c);
}

int cc_cutout_halo(cc_ctx* c, int nhalos, const cc_halo* halos, int pos_only, int64_t n_limit_global) {
    if (!c || nhalos < 0 || (nhalos > 0 && !halos)) return fail(CC_ERR_ARG, "cc_cutout_halo: bad argument");
    if (!c->step_valid) return fail(CC_ERR_STATE, "cc_cutout_halo: no resident step");
    if (!pos_only && (!c->col[COL_VX] || !c->col[COL_VY] || !c->col[COL_VZ] || !c->col[COL_ROT] || !c->col[COL_REP]))
        return fail(CC_ERR_STATE, "cc_cutout_halo: step lacks vx/vy/vz/rotation/replication; pass pos_only=1");
    int rc = use_device(c);
    if (rc) return rc;
    c->p_halos.assign(halos, halos + nhalos);
    c->p_pos_only = pos_only ? 1 : 0;
    c->p_n_limit = n_limit_global;
    c->nhalos = nhalos;
    c->pending = P_HALO;
    c->synced = false;
    c->streamed = false;
    c->payload_timed = false;
    c->index_timed = false;
    c->eager_done = false;
    if (nhalos == 0) { c->counts.clear(); c->rough.clear(); c->seg_begin.assign(1, 0); c->synced = true; return CC_OK; }
    return run_pending(c);
}

/* wait for the results of a cutout whose last kernel was k3_small: that kernel writes everything the host reads to
 * pinned memory itself and then raises h_small[7]; spinning on the flag returns a few microseconds before the
 * driver reports the stream idle (the kernel is still zeroing device state; whatever is enqueued next is ordered
 * behind it by the stream) */
static int wait_small(cc_ctx* c) {
    volatile unsigned long long* flag = c->h_small + 7;
    for (unsigned int spins = 1;; ++spins) {
        if (*flag == c->small_seq) break;
        if ((spins & 0xfffu) == 0) { /* now and then: did the stream end (or fail) without raising the flag? */
            const cudaError_t e = cudaStreamQuery(c->stream);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) return fail(CC_ERR_CUDA, "CUDA error while waiting for a cutout: %s", cudaGetErrorString(e));
        }
        __builtin_ia32_pause();
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return CC_OK;
}

int cc_cutout_sync(cc_ctx* c, int64_t* counts, int64_t* rough_counts) {
    if (!c) return fail(CC_ERR_ARG, "ctx is NULL");
    if (c->pending == P_NONE) return fail(CC_ERR_STATE, "cc_cutout_sync: no cutout pending");
    int rc = use_device(c);
    if (rc) return rc;
    if (!c->synced) {
        if (c->pending == P_CONST) {
            for (int attempt = 0; attempt < 4; ++attempt) {
                CU(cudaStreamSynchronize(c->stream));
                const int64_t k = (int64_t)c->h_small[0];
                const uint64_t max_used = c->h_small[3];
                unsigned int& region_cap = c->region_cap_w;
                if (k <= c->out_cap && max_used <= region_cap) break;
                if (attempt == 3) return fail(CC_ERR_CUDA, "internal: survivor buffers still too small after regrowth");
                /* survivors exceeded the output rows or a CTA's temp region: grow to what was needed, run again */
                if (k > c->out_cap && (rc = ensure_out_cap(c, k + k / 64))) return rc;
                if (max_used > region_cap) region_cap = (unsigned int)std::min<uint64_t>(0xfffffff0u, max_used + max_used / 8 + 256);
                if ((rc = run_pending(c))) return rc;
            }
            c->counts.assign(1, (int64_t)c->h_small[0]);
            c->last_const_count = c->counts[0];
            if (!c->streamed) c->step_last_survivors = c->counts[0];
            c->rough.assign(1, 0);
            c->seg_begin.assign(2, 0);
            c->seg_begin[1] = c->counts[0];
            c->stats[0] = c->n; c->stats[1] = (int64_t)c->h_small[1]; c->stats[2] = c->counts[0]; c->stats[3] = 0;
            c->nan_theta = 0;
            c->particles_read = c->n;
        } else {
            const int H = c->nhalos;
            for (int attempt = 0; attempt < 6; ++attempt) {
                static const bool no_flag_wait = getenv("CC_NO_FLAG_WAIT") != nullptr; /* A/B switch (profiles/r02) */
                if (c->used_small && !c->timing && !no_flag_wait && !c->sync_whole_stream) { if ((rc = wait_small(c))) return rc; }
                else CU(cudaStreamSynchronize(c->stream));
                const int64_t nrec = (int64_t)c->h_small[0];
                const bool refused = c->used_small && c->h_small[4] != 0; /* k3_small found too many records */
                c->force_general = refused;
                /* staged cell-index kernels: a warp's candidate region or the pair buffer was too small */
                const bool staged = c->use_bins && !c->used_mono && !c->used_index;
                const bool cand_over = staged && c->h_small[2] > c->cand_region_cap;
                const bool pair_over = staged && (int64_t)c->h_small[3] > c->pair_cap;
                if (nrec <= c->rec_cap && !refused && !cand_over && !pair_over) break;
                if (attempt == 5) return fail(CC_ERR_CUDA, "internal: halo survivor buffers still too small after regrowth");
                if (cand_over) c->cand_region_cap = (unsigned int)std::min<uint64_t>(0x7ffffff0u, c->h_small[2] + c->h_small[2] / 8 + 512);
                if (pair_over) c->pair_cap = (int64_t)(c->h_small[3] + c->h_small[3] / 8 + 4096);
                if (nrec > c->rec_cap) c->rec_cap = nrec + nrec / 64;
                if ((rc = run_pending(c))) return rc;
            }
            const int64_t nrec = (int64_t)c->h_small[0];
            if (c->use_bins && !c->used_mono && !c->used_index && c->hs && c->hs->walker == 0 && halo_scan_limit(c) > 0) {
                /* the staged pass measured how many particles hit the occupancy bitmap: with many hits the candidate
                 * buffer's round trip through HBM costs more than the one-kernel scan's lower occupancy (profiles/r02) */
                const int nbatch = (H + cck::K2_MAX_HALOS - 1) / cck::K2_MAX_HALOS;
                const double hit = (double)c->h_small[6] / ((double)halo_scan_limit(c) * nbatch);
                c->hs->walker = hit > CC_CELLS_MONO_ABOVE ? 2 : 1;
            }
            c->last_nrec = nrec;
            if (!c->streamed) c->step_last_survivors = nrec;
            if (!c->streamed && c->use_bins) c->step_many_halo_cutouts += 1;
            c->particles_read = (c->use_bins && c->used_index) ? (int64_t)c->h_small[6] : halo_scan_limit(c);
            c->force_general = false;
            c->counts.resize(H); c->rough.resize(H); c->seg_begin.resize(H + 1);
            int64_t tot = 0, tot_rough = 0;
            for (int h = 0; h < H; ++h) {
                c->counts[h] = (int64_t)c->h_small[8 + h];
                c->rough[h] = (int64_t)c->h_small[8 + H + h];
                tot += c->counts[h]; tot_rough += c->rough[h];
            }
            for (int h = 0; h <= H; ++h) c->seg_begin[h] = (int64_t)c->h_small[8 + 2 * H + h];
            if (tot != nrec) return fail(CC_ERR_CUDA, "internal: record count %lld != sum of halo counts %lld", (long long)nrec, (long long)tot);
            c->stats[0] = halo_scan_limit(c); c->stats[1] = (int64_t)c->h_small[1]; c->stats[2] = tot; c->stats[3] = tot_rough;
            c->nan_theta = (int64_t)c->h_small[5];
        }
        c->synced = true;
    }
    for (int h = 0; h < c->nhalos; ++h) {
        if (counts) counts[h] = c->counts[h];
        if (rough_counts) rough_counts[h] = c->rough[h];
    }
    return CC_OK;
}

namespace {
int fetch_rows(cc_ctx* c, int64_t row0, int64_t count, const cc_cols_out* host, const char* who) {
    int rc = use_device(c);
    if (rc) return rc;
    void* dst[N_OUT_COLS] = {host->x, host->y, host->z, host->vx, host->vy, host->vz, host->redshift, host->theta, host->phi,
                             host->id, host->rotation, host->replication, host->theta_u, host->index};
    const bool pos_only = (c->pending == P_HALO) && c->p_pos_only;
    /* streamed step: the device gathered x,y,z,theta,phi(,theta_u),index; the other columns are picked here from
     * the caller's host columns by survivor index */
    const bool by_host = c->streamed && c->host_gather;
    auto host_picked = [](int i) { return i == OC_VX || i == OC_VY || i == OC_VZ || i == OC_RED || i == OC_ID || i == OC_ROT || i == OC_REP; };
    bool need_index = false;
    for (int i = 0; i < N_OUT_COLS; ++i) {
        if (!dst[i] || count == 0) continue;
        if (pos_only && (i == OC_VX || i == OC_VY || i == OC_VZ || i == OC_ROT || i == OC_REP))
            return fail(CC_ERR_ARG, "%s: velocity/rotation/replication were not gathered (pos_only)", who);
        if (c->pending == P_CONST && i == OC_THU) return fail(CC_ERR_ARG, "%s: theta_u exists only for halo cutouts", who);
        if (by_host && host_picked(i)) { need_index = true; continue; }
        CU(cudaMemcpyAsync(dst[i], (const char*)c->out_buf[i].p + (size_t)row0 * OUT_COL_SIZE[i], (size_t)count * OUT_COL_SIZE[i],
                           cudaMemcpyDeviceToHost, c->stream));
    }
    const int64_t* index = host->index;
    if (need_index && !index) {
        c->fetch_index.resize((size_t)count);
        index = c->fetch_index.data();
        CU(cudaMemcpyAsync(c->fetch_index.data(), (const char*)c->out_buf[OC_IDX].p + (size_t)row0 * 8, (size_t)count * 8,
                           cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    if (need_index) {
        const cc_cols_in& in = c->host_cols;
        const int64_t base = c->index_base;
        /* threads: this rank's share of the host cores (launchers such as torchrun pin OMP_NUM_THREADS to 1, which
         * would leave a million random reads per column to one core); CC_HOST_THREADS overrides */
        int nt = std::max(1, (int)std::thread::hardware_concurrency() / std::max(1, c->nranks));
        nt = std::min(nt, 32);
        if (const char* e = getenv("CC_HOST_THREADS")) nt = std::max(1, atoi(e));
        if (count <= 65536) nt = 1;
        /* every row costs up to seven cache misses at unrelated addresses: ask for the lines of the row PF ahead while
         * this one is copied (the survivor indices are known), so that each thread keeps ~PF x 7 misses in flight */
        const int64_t PF = 12;
#pragma omp parallel for schedule(static) num_threads(nt)
        for (int64_t r = 0; r < count; ++r) {
            if (r + PF < count) {
                const int64_t p = index[r + PF] - base;
                if (host->vx) __builtin_prefetch(in.vx + p);
                if (host->vy) __builtin_prefetch(in.vy + p);
                if (host->vz) __builtin_prefetch(in.vz + p);
                if (host->redshift) __builtin_prefetch(in.a + p);
                if (host->id) __builtin_prefetch(in.id + p);
                if (host->rotation) __builtin_prefetch(in.rotation + p);
                if (host->replication) __builtin_prefetch(in.replication + p);
            }
            const int64_t g = index[r] - base;
            if (host->vx) host->vx[r] = in.vx[g];
            if (host->vy) host->vy[r] = in.vy[g];
            if (host->vz) host->vz[r] = in.vz[g];
            if (host->redshift) host->redshift[r] = ccref::a_to_z(in.a[g]); /* processLC.cpp:305 / :1450 */
            if (host->id) host->id[r] = in.id[g];
            if (host->rotation) host->rotation[r] = in.rotation[g];
            if (host->replication) host->replication[r] = in.replication[g];
        }
    }
    return CC_OK;
}
}  // namespace

int cc_cutout_fetch(cc_ctx* c, int halo, int64_t first, int64_t count, const cc_cols_out* host) {
    if (!c || !host) return fail(CC_ERR_ARG, "cc_cutout_fetch: bad argument");
    if (c->pending == P_NONE || !c->synced) return fail(CC_ERR_STATE, "cc_cutout_fetch: call cc_cutout_sync first");
    if (halo < 0 || halo >= c->nhalos) return fail(CC_ERR_ARG, "cc_cutout_fetch: halo %d out of range", halo);
    if (first < 0 || count < 0 || first + count > c->counts[halo]) return fail(CC_ERR_ARG, "cc_cutout_fetch: rows out of range");
    return fetch_rows(c, c->seg_begin[halo] + first, count, host, "cc_cutout_fetch");
}

int cc_cutout_fetch_all(cc_ctx* c, int64_t total, const cc_cols_out* host) {
    if (!c || !host) return fail(CC_ERR_ARG, "cc_cutout_fetch_all: bad argument");
    if (c->pending == P_NONE || !c->synced) return fail(CC_ERR_STATE, "cc_cutout_fetch_all: call cc_cutout_sync first");
    const int64_t have = c->nhalos > 0 ? c->seg_begin[(size_t)c->nhalos] : 0;
    if (total != have) return fail(CC_ERR_ARG, "cc_cutout_fetch_all: total %lld is not the survivor count %lld", (long long)total, (long long)have);
    return fetch_rows(c, 0, total, host, "cc_cutout_fetch_all");
}

int cc_cutout_device_cols(cc_ctx* c, int halo, cc_cols_out* dev, int64_t* row0, int64_t* count) {
    if (!c || !dev) return fail(CC_ERR_ARG, "cc_cutout_device_cols: bad argument");
    if (c->pending == P_NONE || !c->synced) return fail(CC_ERR_STATE, "cc_cutout_device_cols: call cc_cutout_sync first");
    if (halo < 0 || halo >= c->nhalos) return fail(CC_ERR_ARG, "halo out of range");
    const cck::ColsOut o = cols_out_of(c);
    dev->x = o.x; dev->y = o.y; dev->z = o.z; dev->vx = o.vx; dev->vy = o.vy; dev->vz = o.vz;
    dev->redshift = o.redshift; dev->theta = o.theta; dev->phi = o.phi; dev->id = (int64_t*)o.id;
    dev->rotation = o.rot; dev->replication = o.rep; dev->theta_u = o.theta_u; dev->index = (int64_t*)o.index;
    if (c->streamed && c->host_gather) /* never gathered on the device: see cc_cutout_*_streamed */
        dev->vx = dev->vy = dev->vz = dev->redshift = nullptr, dev->id = nullptr, dev->rotation = dev->replication = nullptr;
    if (row0) *row0 = c->seg_begin[halo];
    if (count) *count = c->counts[halo];
    return CC_OK;
}

/* ------------------------------------------------------------------ streamed step */
namespace {
/* a streamed step needs host columns the GPU can address: pinned (cudaHostAlloc / cudaHostRegister) memory */
int map_host_col(const void* h, const char* name, bool required, const void** dev, bool* on_host) {
    *dev = nullptr;
    *on_host = true;
    if (!h) return required ? fail(CC_ERR_ARG, "streamed cutout: column %s is required", name) : CC_OK;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, h);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CC_ERR_ARG, "streamed cutout: cannot query column %s", name); }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) { *dev = h; *on_host = false; return CC_OK; }
    if (at.type != cudaMemoryTypeHost)
        return fail(CC_ERR_ARG, "streamed cutout: column %s is pageable host memory; allocate it with cc_host_alloc or pin it "
                    "with cc_host_register (or use cc_step_upload for pageable data)", name);
    void* d = nullptr;
    e = cudaHostGetDevicePointer(&d, const_cast<void*>(h), 0);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CC_ERR_CUDA, "cudaHostGetDevicePointer(%s): %s", name, cudaGetErrorString(e)); }
    *dev = d;
    return CC_OK;
}

int begin_streamed(cc_ctx* c, int64_t n, const cc_cols_in* host, int64_t index_base, bool need_vel) {
    int rc = use_device(c);
    if (rc) return rc;
    if (c->step_owned) { /* a streamed step replaces the resident one */
        if ((rc = cc_step_release(c))) return rc;
    }
    const void* h[N_IN_COLS] = {host->x, host->y, host->z, host->vx, host->vy, host->vz, host->a, host->id, host->rotation, host->replication};
    const char* names[N_IN_COLS] = {"x", "y", "z", "vx", "vy", "vz", "a", "id", "rotation", "replication"};
    const void* d[N_IN_COLS];
    bool all_host = true;
    for (int i = 0; i < N_IN_COLS; ++i) {
        const bool vel = (i == COL_VX || i == COL_VY || i == COL_VZ || i == COL_ROT || i == COL_REP);
        bool on_host = true;
        if ((rc = map_host_col(h[i], names[i], vel ? need_vel : true, &d[i], &on_host))) return rc;
        if (i != COL_X && i != COL_Y && i != COL_Z) all_host = all_host && on_host;
    }
    c->host_gather = all_host && !getenv("CC_DEVICE_GATHER");
    c->host_cols = *host;
    c->gcols.x = (const float*)d[COL_X]; c->gcols.y = (const float*)d[COL_Y]; c->gcols.z = (const float*)d[COL_Z];
    c->gcols.vx = (const float*)d[COL_VX]; c->gcols.vy = (const float*)d[COL_VY]; c->gcols.vz = (const float*)d[COL_VZ];
    c->gcols.a = (const float*)d[COL_A]; c->gcols.id = (const long long*)d[COL_ID];
    c->gcols.rot = (const int*)d[COL_ROT]; c->gcols.rep = (const int*)d[COL_REP];
    for (int i = 0; i < N_IN_COLS; ++i) c->col[i] = nullptr;
    c->step_valid = false;
    c->n = n;
    c->index_base = index_base;
    c->streamed = true;
    c->synced = false;
    c->payload_timed = false;
    c->index_timed = false;
    return CC_OK;
}
}  // namespace

int cc_cutout_const_streamed(cc_ctx* c, int64_t n, const cc_cols_in* host, int64_t index_base,
                             const float theta_cut[2], const float phi_cut[2]) {
    if (!c || !host || n < 0 || !theta_cut || !phi_cut) return fail(CC_ERR_ARG, "cc_cutout_const_streamed: bad argument");
    int rc = begin_streamed(c, n, host, index_base, true);
    if (rc) return rc;
    c->p_th[0] = theta_cut[0]; c->p_th[1] = theta_cut[1];
    c->p_ph[0] = phi_cut[0]; c->p_ph[1] = phi_cut[1];
    c->nhalos = 1;
    c->pending = P_CONST;
    c->eager_done = false;
    return run_pending(c);
}

int cc_cutout_halo_streamed(cc_ctx* c, int64_t n, const cc_cols_in* host, int64_t index_base, int nhalos,
                            const cc_halo* halos, int pos_only, int64_t n_limit_global) {
    if (!c || !host || n < 0 || nhalos < 0 || (nhalos > 0 && !halos)) return fail(CC_ERR_ARG, "cc_cutout_halo_streamed: bad argument");
    int rc = begin_streamed(c, n, host, index_base, !pos_only);
    if (rc) return rc;
    c->p_halos.assign(halos, halos + nhalos);
    c->p_pos_only = pos_only ? 1 : 0;
    c->p_n_limit = n_limit_global;
    c->nhalos = nhalos;
    c->pending = P_HALO;
    c->eager_done = false;
    if (nhalos == 0) { c->counts.clear(); c->rough.clear(); c->seg_begin.assign(1, 0); c->synced = true; return CC_OK; }
    return run_pending(c);
}

/* ------------------------------------------------------------------ pinned host memory */
int cc_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(CC_ERR_ARG, "ptr is NULL");
    CU(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return CC_OK;
}
int cc_host_free(void* ptr) {
    CU(cudaFreeHost(ptr));
    return CC_OK;
}
int cc_host_register(void* ptr, size_t bytes) {
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return CC_OK;
}
int cc_host_unregister(void* ptr) {
    CU(cudaHostUnregister(ptr));
    return CC_OK;
}

/* ------------------------------------------------------------------ communicator + count exchange */
int cc_comm_unique_id(void* id128) {
    if (!id128) return fail(CC_ERR_ARG, "id is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == CC_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return CC_OK;
}

namespace {
/* The exchange moves {count, rough count} per halo -- a few bytes to a few KB -- so the communicator is limited to
 * ONE CTA.  This is not only thrift: with NCCL 2.28.9 (P2P over cuMem) and its default CTA count, the cell-index
 * kernel of the next step, which needs every SM free for its one 161 KB-shared-memory CTA per SM, ran 1.65x slower
 * (1.31 ms instead of 0.79 ms at 2 GPUs; 2.27.3, NCCL_P2P_DISABLE=1, NCCL_CUMEM_ENABLE=0 or NCCL_MAX_CTAS=1 all
 * remove the effect; profiles/r01/README.md). */
ncclResult_t comm_init_rank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
    if (g_nccl.CommInitRankConfig && !getenv("CC_NCCL_DEFAULT_CTAS")) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = 1;
        const ncclResult_t r = g_nccl.CommInitRankConfig(comm, nranks, id, rank, &cfg);
        if (r != ncclInvalidArgument && r != ncclInvalidUsage) return r;
        /* a library older than the header this was compiled against may refuse the config: plain init */
    }
    return g_nccl.CommInitRank(comm, nranks, id, rank);
}
}  // namespace

int cc_comm_init(cc_ctx* c, int nranks, int rank, const void* id128) {
    if (!c || nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(CC_ERR_ARG, "cc_comm_init: bad argument");
    int rc = nccl_load();
    if (rc) return rc;
    if ((rc = use_device(c))) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NC(comm_init_rank(&c->comm, nranks, id, rank));
    c->nranks = nranks;
    c->rank = rank;
    return CC_OK;
}

int cc_comm_init_all(cc_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return fail(CC_ERR_ARG, "cc_comm_init_all: bad argument");
    int rc = nccl_load();
    if (rc) return rc;
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    { /* ncclCommInitAll, spelled out so that the one-CTA configuration applies (comm_init_rank) */
        ncclUniqueId id;
        NC(g_nccl.GetUniqueId(&id));
        NC(g_nccl.GroupStart());
        ncclResult_t first_err = ncclSuccess;
        for (int i = 0; i < n; ++i) {
            if (cudaSetDevice(devs[i]) != cudaSuccess) { cudaGetLastError(); first_err = ncclUnhandledCudaError; break; }
            const ncclResult_t r = comm_init_rank(&comms[i], n, id, i);
            if (r != ncclSuccess && first_err == ncclSuccess) first_err = r;
        }
        const ncclResult_t ge = g_nccl.GroupEnd();
        NC(first_err);
        NC(ge);
    }
    for (int i = 0; i < n; ++i) { ctxs[i]->comm = comms[i]; ctxs[i]->nranks = n; ctxs[i]->rank = i; }
    /* same process: mailboxes are reachable through plain peer access, no IPC handles needed */
    if (n <= cck::XCHG_MAX_RANKS) {
        bool ok = true;
        for (int i = 0; i < n && ok; ++i)
            for (int j = 0; j < n && ok; ++j) {
                if (i == j) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can) { ok = false; break; }
                cudaSetDevice(devs[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
        const size_t bytes = (size_t)2 * n * cck::XCHG_SLOT * 8;
        for (int i = 0; i < n && ok; ++i) {
            cudaSetDevice(devs[i]);
            ctxs[i]->mailbox.release();
            if (ctxs[i]->mailbox.ensure(bytes) != CC_OK || cudaMemset(ctxs[i]->mailbox.p, 0, bytes) != cudaSuccess) ok = false;
            cudaDeviceSynchronize();
        }
        for (int i = 0; i < n && ok; ++i) {
            for (int j = 0; j < n; ++j) ctxs[i]->peer_mb[j] = (unsigned long long*)ctxs[j]->mailbox.p;
            ctxs[i]->peers_ready = true;
            ctxs[i]->xchg_epoch = 0;
        }
    }
    return CC_OK;
}

int cc_comm_rank(cc_ctx* c) { return c ? c->rank : -1; }
int cc_comm_size(cc_ctx* c) { return c ? c->nranks : -1; }

int cc_exchange_counts(cc_ctx* c, int64_t* all_counts, int64_t* all_rough, int64_t* offsets) {
    if (!c) return fail(CC_ERR_ARG, "ctx is NULL");
    if (c->pending == P_NONE) return fail(CC_ERR_STATE, "cc_exchange_counts: no cutout pending");
    int rc = use_device(c);
    if (rc) return rc;
    const int H = c->nhalos, R = c->nranks;
    if (H == 0) return cc_cutout_sync(c, nullptr, nullptr);
    if ((rc = c->counts_buf.ensure(16))) return rc;
    if ((rc = c->gather_buf.ensure((size_t)R * 2 * H * 8 + 64))) return rc;
    if ((rc = c->offsets_buf.ensure((size_t)H * 8))) return rc;
    const size_t hg_off = 32 + 3 * (size_t)H; /* behind the cutout's own read-back area */
    if (c->h_small_elems < small_elems(H, R)) { /* communicator created after the cutout was enqueued */
        CU(cudaStreamSynchronize(c->stream));
        std::vector<unsigned long long> keep(c->h_small, c->h_small + c->h_small_elems);
        if ((rc = ensure_small(c, small_elems(H, R)))) return rc;
        memcpy(c->h_small, keep.data(), keep.size() * sizeof(unsigned long long));
    }
    /* The collective is enqueued on the exchange stream behind the event that marks the counts final (after the scan
     * kernel of use case 2, after the tile scan of use case 1): no host round trip between the scan and the
     * exchange, and the exchange runs NEXT TO the ordering / gather kernels instead of after them.  The cutout stream
     * then waits for the exchange, so stream order (and anything timed on the cutout stream) includes it.  Survivor
     * counts are exact even when a survivor buffer overflowed (only the stored rows are clipped), so a local re-run
     * after the synchronisation never needs a second collective -- every rank issues exactly one all-gather per call. */
    unsigned long long* hg = c->h_small + hg_off;
    if (c->pending == P_HALO && c->eager_done) {
        /* the cutout's last kernel already exchanged the counts (k3_small, exchange=eager): nothing to enqueue */
        if (c->synced) CU(cudaStreamSynchronize(c->stream));
        else if ((rc = cc_cutout_sync(c, nullptr, nullptr))) return rc;
        goto read_back;
    }
    {
    cudaStream_t xs = c->xchg_stream;
    CU(cudaStreamWaitEvent(xs, c->ev_counts, 0));
    if (c->xchg_mode >= 2 && R > 1 && !(c->peers_ready && 2 * H <= cck::XCHG_PAYLOAD))
        return fail(CC_ERR_STATE, "cc_exchange_counts: exchange=peer but no mailboxes are attached (or more than %d halos)", cck::XCHG_PAYLOAD / 2);
    const bool peer_path = c->peers_ready && R > 1 && 2 * H <= cck::XCHG_PAYLOAD && c->xchg_mode != 1 && !getenv("CC_NO_PEER_XCHG");
    if (peer_path) { /* NVLink peer stores: one single-CTA kernel, no collective launch, results written to pinned memory */
        cck::XchgArgs x;
        /* use case 1 sends {survivor total, 0} straight from the scan's result words */
        const void* send = (c->pending == P_CONST) ? c->scratch.p : c->xchg_src;
        if ((rc = fill_xchg_args(c, H, send, c->pending == P_CONST ? 1 : 0, &x))) return rc;
        cck::k4_exchange_peer<<<1, 256, 0, xs>>>(x);
        CU(cudaGetLastError());
    } else {
        if (c->pending == P_CONST) { /* device-resident send buffer {count, 0}; use case 2 already has {counts, rough} */
            CU(cudaMemcpyAsync(c->counts_buf.p, c->scratch.p, 8, cudaMemcpyDeviceToDevice, xs));
            CU(cudaMemsetAsync((char*)c->counts_buf.p + 8, 0, 8, xs));
        }
        const void* send = (c->pending == P_CONST) ? c->counts_buf.p : c->xchg_src;
        unsigned long long* d_status = (unsigned long long*)((char*)c->gather_buf.p + (size_t)R * 2 * H * 8);
        CU(cudaMemsetAsync(d_status, 0, 8, xs));
        if (c->comm && R > 1) {
            NC(g_nccl.AllGather(send, c->gather_buf.p, (size_t)2 * H, ncclUint64, c->comm, xs));
        } else {
            if (R > 1) return fail(CC_ERR_STATE, "cc_exchange_counts: %d ranks but neither an NCCL communicator nor peer mailboxes for %d halos", R, H);
            CU(cudaMemcpyAsync(c->gather_buf.p, send, (size_t)2 * H * 8, cudaMemcpyDeviceToDevice, xs));
        }
        cck::k4_offsets<<<(H + 255) / 256, 256, 0, xs>>>((const unsigned long long*)c->gather_buf.p,
                                                              (unsigned long long*)c->offsets_buf.p, H, c->rank);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hg, c->gather_buf.p, (size_t)R * 2 * H * 8 + 8, cudaMemcpyDeviceToHost, xs)); /* + status */
        CU(cudaMemcpyAsync(hg + (size_t)R * 2 * H + 1, c->offsets_buf.p, (size_t)H * 8, cudaMemcpyDeviceToHost, xs));
    }
    c->launches += 1;
    CU(cudaEventRecord(c->ev_xchg, xs));
    CU(cudaStreamWaitEvent(c->stream, c->ev_xchg, 0));
    if (c->synced) CU(cudaStreamSynchronize(c->stream));
    else {
        /* one synchronisation for cutout + exchange: the exchange ran on its own stream, so the flag a small cutout's
         * last kernel raises says nothing about it -- wait for the cutout stream, which waits for the exchange */
        c->sync_whole_stream = true;
        rc = cc_cutout_sync(c, nullptr, nullptr);
        c->sync_whole_stream = false;
        if (rc) return rc;
    }
    }
read_back:
    for (int r = 0; r < R; ++r)
        for (int h = 0; h < H; ++h) {
            if (all_counts) all_counts[(size_t)r * H + h] = (int64_t)hg[(size_t)r * 2 * H + h];
            if (all_rough) all_rough[(size_t)r * H + h] = (int64_t)hg[(size_t)r * 2 * H + H + h];
        }
    if (hg[(size_t)R * 2 * H] != 0)
        return fail(CC_ERR_NCCL, "cc_exchange_counts: peer exchange timed out (a rank did not enter the exchange)");
    if (offsets)
        for (int h = 0; h < H; ++h) offsets[h] = (int64_t)hg[(size_t)R * 2 * H + 1 + h];
    return CC_OK;
}

int cc_comm_peer_handle(cc_ctx* c, int nranks, int rank, void* handle64) {
    if (!c || !handle64 || nranks < 1 || nranks > cck::XCHG_MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(CC_ERR_ARG, "cc_comm_peer_handle: bad argument (at most %d ranks)", cck::XCHG_MAX_RANKS);
    if (c->comm && (c->nranks != nranks || c->rank != rank)) return fail(CC_ERR_ARG, "cc_comm_peer_handle: rank/size differ from the NCCL communicator");
    int rc = use_device(c);
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == CC_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    const size_t bytes = (size_t)2 * nranks * cck::XCHG_SLOT * 8;
    c->mailbox.release();
    if ((rc = c->mailbox.ensure(bytes))) return rc;
    CU(cudaMemset(c->mailbox.p, 0, bytes));
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->mailbox.p));
    memcpy(handle64, &h, sizeof(h));
    c->nranks = nranks;
    c->rank = rank;
    c->peers_ready = false;
    c->xchg_epoch = 0;
    return CC_OK;
}

int cc_comm_peer_attach(cc_ctx* c, const void* handles) {
    if (!c || !handles) return fail(CC_ERR_ARG, "cc_comm_peer_attach: bad argument");
    if (!c->mailbox.p) return fail(CC_ERR_STATE, "cc_comm_peer_attach: call cc_comm_peer_handle first");
    int rc = use_device(c);
    if (rc) return rc;
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) { c->peer_mb[r] = (unsigned long long*)c->mailbox.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * CC_PEER_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(CC_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        c->peer_mb[r] = (unsigned long long*)p;
        c->peer_opened[r] = true;
    }
    c->peers_ready = true;
    return CC_OK;
}

/* ------------------------------------------------------------------ measurement */
double cc_last_scan_ms(cc_ctx* c) {
    if (!c || !c->have_times) return -1.0;
    cudaSetDevice(c->device);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1.0;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev_scan0, c->ev_scan1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    return ms;
}
double cc_last_total_ms(cc_ctx* c) {
    if (!c || !c->have_times || !c->synced) return -1.0;
    cudaSetDevice(c->device);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1.0;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev_scan0, c->ev_total1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    return ms;
}
int64_t cc_launch_count(cc_ctx* c) { return c ? c->launches : 0; }
int cc_last_stats(cc_ctx* c, int64_t out[4]) {
    if (!c || !out) return fail(CC_ERR_ARG, "bad argument");
    if (!c->synced) return fail(CC_ERR_STATE, "no synced cutout");
    for (int i = 0; i < 4; ++i) out[i] = c->stats[i];
    return CC_OK;
}

int cc_last_stats_ex(cc_ctx* c, int64_t* out, int n) {
    if (!c || !out || n < 0) return fail(CC_ERR_ARG, "bad argument");
    if (!c->synced) return fail(CC_ERR_STATE, "no synced cutout");
    int64_t v[CC_STATS_EX_FIELDS] = {};
    for (int i = 0; i < 4; ++i) v[i] = c->stats[i];
    v[4] = c->nan_theta;
    v[5] = c->payload_valid ? 1 : 0;
    if (c->payload_timed) {
        cudaSetDevice(c->device);
        float ms = 0.0f;
        if (cudaStreamSynchronize(c->stream) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev_pay0, c->ev_pay1) == cudaSuccess)
            v[6] = (int64_t)(ms * 1000.0f);
        else cudaGetLastError();
    }
    v[7] = (c->pending == P_HALO) ? c->n - c->stats[0] : 0;
    v[8] = (c->pending != P_HALO) ? 0 : !c->use_bins ? 1 : c->used_index ? 4 : c->used_mono ? 3 : 2;
    v[9] = c->particles_read;
    v[10] = c->pidx_valid ? 1 : 0;
    if (c->index_timed) {
        cudaSetDevice(c->device);
        float ms = 0.0f;
        if (cudaStreamSynchronize(c->stream) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev_idx0, c->ev_idx1) == cudaSuccess)
            v[11] = (int64_t)(ms * 1000.0f);
        else cudaGetLastError();
    }
    for (int i = 0; i < n && i < CC_STATS_EX_FIELDS; ++i) out[i] = v[i];
    return CC_OK;
}

int cc_set_option(cc_ctx* c, const char* key, const char* value) {
    if (!c || !key || !value) return fail(CC_ERR_ARG, "cc_set_option: bad argument");
    const std::string k = key, v = value;
    if (k == "index") {
        if (v == "auto") c->index_policy = 0; else if (v == "always") c->index_policy = 1;
        else if (v == "never") c->index_policy = 2; else return fail(CC_ERR_ARG, "cc_set_option: index = auto|always|never");
        return CC_OK;
    }
    if (k == "payload") {
        if (v == "auto") c->payload_policy = 0; else if (v == "always") c->payload_policy = 1;
        else if (v == "never") c->payload_policy = 2; else return fail(CC_ERR_ARG, "cc_set_option: payload = auto|always|never");
        if (c->payload_policy == 2 && c->pending == P_NONE) c->payload_valid = false;
    } else if (k == "exchange") {
        if (v == "auto") c->xchg_mode = 0; else if (v == "nccl") c->xchg_mode = 1; else if (v == "peer") c->xchg_mode = 2;
        else if (v == "eager") c->xchg_mode = 3;
        else return fail(CC_ERR_ARG, "cc_set_option: exchange = auto|nccl|peer|eager");
    } else if (k == "cells") {
        if (v == "auto") c->cells_mode = 0; else if (v == "staged") c->cells_mode = 1; else if (v == "mono") c->cells_mode = 2;
        else return fail(CC_ERR_ARG, "cc_set_option: cells = auto|staged|mono");
    } else if (k == "timing") {
        if (v == "on") c->timing = true; else if (v == "off") c->timing = false;
        else return fail(CC_ERR_ARG, "cc_set_option: timing = on|off");
    } else if (k == "gather") {
        if (v == "auto") c->gather_mode = 0; else if (v == "ldst") c->gather_mode = 1; else if (v == "bulk") c->gather_mode = 2;
        else return fail(CC_ERR_ARG, "cc_set_option: gather = auto|ldst|bulk");
    } else {
        return fail(CC_ERR_ARG, "cc_set_option: unknown key '%s'", key);
    }
    return CC_OK;
}

/* ------------------------------------------------------------------ test hooks */
int cc_debug_eval(cc_ctx* c, int op, int64_t n, const float* in0, const float* in1, const float* in2, float* out0) {
    if (!c || n < 0 || !in0 || !out0 || op < 0 || op > 6) return fail(CC_ERR_ARG, "cc_debug_eval: bad argument");
    if ((op == 2 && (!in1 || !in2)) || ((op == 3 || op == 4) && !in1)) return fail(CC_ERR_ARG, "cc_debug_eval: missing input");
    int rc = use_device(c);
    if (rc) return rc;
    if (n == 0) return CC_OK;
    DevBuf d0, d1, d2, dout;
    const size_t bytes = (size_t)n * 4;
    if ((rc = d0.ensure(bytes)) || (rc = dout.ensure(bytes))) { d0.release(); dout.release(); return rc; }
    if (in1 && (rc = d1.ensure(bytes))) { d0.release(); dout.release(); return rc; }
    if (in2 && (rc = d2.ensure(bytes))) { d0.release(); d1.release(); dout.release(); return rc; }
    cudaError_t e = cudaMemcpyAsync(d0.p, in0, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && in1) e = cudaMemcpyAsync(d1.p, in1, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && in2) e = cudaMemcpyAsync(d2.p, in2, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        cck::k_debug_eval<<<c->sm_count * 8, 256, 0, c->stream>>>(op, n, (const float*)d0.p, (const float*)d1.p, (const float*)d2.p, (float*)dout.p);
        e = cudaGetLastError();
        c->launches += 1;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out0, dout.p, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    d0.release(); d1.release(); d2.release(); dout.release();
    if (e != cudaSuccess) return fail(CC_ERR_CUDA, "cc_debug_eval: %s", cudaGetErrorString(e));
    return CC_OK;
}

int cc_debug_eval_host(int op, int64_t n, const float* in0, const float* in1, const float* in2, float* out0) {
    if (n < 0 || !in0 || !out0 || op < 0 || op > 6) return fail(CC_ERR_ARG, "cc_debug_eval_host: bad argument");
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float r;
        switch (op) {
            case 0: r = ccref::acosf_ref(in0[i]); break;
            case 1: r = ccref::atanf_ref(in0[i]); break;
            case 2: r = ccref::theta_arcsec(in0[i], in1[i], in2[i]); break;
            case 3: r = ccref::phi_arcsec_plain(in0[i], in1[i]); break;
            case 4: r = ccref::phi_arcsec_guarded(in0[i], in1[i]); break;
            case 5: r = ccref::a_to_z(in0[i]); break;
            default: r = ccref::to_arcsec(in0[i]); break; /* 6: specification form (true fp64 division) */
        }
        out0[i] = r;
    }
    return CC_OK;
}

}  // extern "C"
