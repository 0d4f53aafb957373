/*
 * kernels.cuh -- sm_100a kernels of the lightcone cutout path.
 *
 *   k1s_scan / k1s_seg_scan / k1s_gather (+ _bulk)
 *                     use case 1: octant test -> pre-filter -> exact (theta,phi) -> strict window test ->
 *                     ballot/popc compaction into per-warp record regions; exclusive scan of the segment counts by
 *                     decoupled look-back; order-preserving 12-column gather.   replaces processLC.cpp:276-310
 *   k2_cutout_halo (+ _cells, _tma)
 *                     use case 2: one pass over x,y,z for a batch of halos: pre-filter -> exact un-rotated
 *                     (theta_u,phi_u) -> rough test -> R.v -> exact rotated angles -> fine test -> survivor
 *                     records.                               replaces processLC.cpp:897-909, :1334-1340, :1383-1481
 *   k3_scan_counts / k3_bucket / k3_rank / k3_gather_halo, k3_small
 *                     survivor ordering by (theta_u, input index) per halo and column gather
 *                                                            replaces processLC.cpp:1214-1221 (restricted to
 *                                                            survivors) and :1446-1461
 *   k4_offsets / k4_exchange_peer
 *                     exclusive scan over ranks of the exchanged counts (after an NCCL all-gather, or with the
 *                     exchange itself done by NVLink peer stores)   replaces processLC.cpp:331-337, :1569-1583
 *   k_build_payload, k_synth, k_debug_eval                   residency helper, synthetic data, test hook
 *
 * Exactness: whatever reaches an output column or decides membership is computed by ccref:: functions
 * (ref_math.cuh) -- the reference's operation sequence.  The pre-filter is the only approximate code; it may
 * only REJECT particles the exact test provably rejects (margins derived in prefilter.h), every other
 * particle takes the exact path.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ref_math.cuh"

namespace cck {

struct ColsIn {
    const float *x, *y, *z, *vx, *vy, *vz, *a;
    const long long* id;
    const int *rot, *rep;
};
struct ColsOut {
    float *x, *y, *z, *vx, *vy, *vz, *redshift, *theta, *phi;
    long long* id;
    int *rot, *rep;
    float* theta_u;
    long long* index;
};

/* magnitude range inside which the pre-filter's float products can neither overflow nor lose relative
 * precision to underflow; anything outside goes straight to the exact path */
#define CC_SANE_LO 9.3132257e-10f /* 2^-30 */
#define CC_SANE_HI 1.0995116e12f  /* 2^40  */

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/* x86 SSE quiets and propagates a NaN operand's payload; the GPU returns the canonical NaN.  redshift is
 * the only output column computed from a possibly-NaN input, so mirror the host there. */
__device__ __forceinline__ float a_to_z_dev(float a) {
    if (a != a) return __int_as_float(__float_as_int(a) | 0x00400000);
    return ccref::a_to_z(a);
}

/* ---- gather payload -------------------------------------------------------------------------------------------
 * x,y,z are the only columns the scan streams; the other seven are only ever read for survivors, one 4- or 8-byte
 * element at a time, and every such read costs a 64-128 B DRAM burst (profiles/r01: ~110 B of DRAM traffic per
 * gathered element).  A resident step therefore also keeps those seven columns interleaved, 32 B per particle, so a
 * survivor costs one sector instead of seven bursts.  Built once per step by k_build_payload (one extra pass over
 * 32 B/particle at residency time, amortised over every cutout on the resident step). */
struct __align__(16) Payload {
    float vx, vy, vz, a;
    long long id;
    int rot, rep;
};
__global__ void k_build_payload(ColsIn in, Payload* __restrict__ out, long long n, int with_vel) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Payload p;
        p.a = in.a[i];
        p.id = in.id[i];
        if (with_vel) { p.vx = in.vx[i]; p.vy = in.vy[i]; p.vz = in.vz[i]; p.rot = in.rot[i]; p.rep = in.rep[i]; }
        else { p.vx = p.vy = p.vz = 0.0f; p.rot = p.rep = 0; }
        out[i] = p;
    }
}
__device__ __forceinline__ Payload load_payload(const Payload* p) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p)), hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    Payload r;
    r.vx = __uint_as_float(lo.x); r.vy = __uint_as_float(lo.y); r.vz = __uint_as_float(lo.z); r.a = __uint_as_float(lo.w);
    r.id = (long long)(((unsigned long long)hi.y << 32) | hi.x);
    r.rot = (int)hi.z; r.rep = (int)hi.w;
    return r;
}

/* approximate reciprocal / reciprocal square root for the pre-filter only (arguments are range-checked, so
 * flushing subnormals is harmless and the slow-path fix-up code of the non-ftz forms is avoided) */
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_TILE = 256; /* particles per warp tile: 8 per lane = two coalesced float4 chunks per column */

/* load one warp tile: chunk A = [base+4*lane, +4), chunk B = [base+128+4*lane, +4); returns the valid mask */
template <bool VEC>
__device__ __forceinline__ unsigned int load_tile(const float* __restrict__ x, const float* __restrict__ y,
                                                  const float* __restrict__ z, long long base, long long n, int lane,
                                                  float (&px)[8], float (&py)[8], float (&pz)[8]) {
    if (VEC && base + SCAN_TILE <= n) {
        const float4 xa = ldg_stream4(x + base + 4 * lane), xb = ldg_stream4(x + base + 128 + 4 * lane);
        const float4 ya = ldg_stream4(y + base + 4 * lane), yb = ldg_stream4(y + base + 128 + 4 * lane);
        const float4 za = ldg_stream4(z + base + 4 * lane), zb = ldg_stream4(z + base + 128 + 4 * lane);
        px[0] = xa.x; px[1] = xa.y; px[2] = xa.z; px[3] = xa.w; px[4] = xb.x; px[5] = xb.y; px[6] = xb.z; px[7] = xb.w;
        py[0] = ya.x; py[1] = ya.y; py[2] = ya.z; py[3] = ya.w; py[4] = yb.x; py[5] = yb.y; py[6] = yb.z; py[7] = yb.w;
        pz[0] = za.x; pz[1] = za.y; pz[2] = za.z; pz[3] = za.w; pz[4] = zb.x; pz[5] = zb.y; pz[6] = zb.z; pz[7] = zb.w;
        return 0xffu;
    }
    unsigned int valid = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const long long g = base + (j >> 2) * 128 + 4 * lane + (j & 3);
        if (g < n) {
            px[j] = __ldcs(x + g); py[j] = __ldcs(y + g); pz[j] = __ldcs(z + g);
            valid |= 1u << j;
        } else {
            px[j] = py[j] = pz[j] = 1.0f;
        }
    }
    return valid;
}
/* ask the L2 for the 256-particle tile that starts at `base` (3 columns x 8 lines of 128 B: lanes 0-23, one line each), so
 * that the warp's next load_tile finds it there instead of waiting for DRAM */
__device__ __forceinline__ void prefetch_tile(const float* __restrict__ x, const float* __restrict__ y,
                                              const float* __restrict__ z, long long base, long long n, int lane) {
    if (lane < 24 && base + SCAN_TILE <= n) {
        const float* col = lane < 8 ? x : (lane < 16 ? y : z);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(col + base + (lane & 7) * 32));
    }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
/* index inside the tile of element j of this lane */
__device__ __forceinline__ unsigned int tile_local(int lane, int j) { return (unsigned)((j >> 2) * 128 + 4 * lane + (j & 3)); }

/* Branch-free on purpose: eight of these run back to back per lane and divergent early-outs cost more than the
 * arithmetic they skip.  Each window condition is turned into a difference by ONE fma -- fma(a,b,-c) has exactly the
 * sign of a*b - c, so `d > 0` is the exact comparison a*b > c -- and the four differences are reduced with min, so the
 * whole window is one compare instead of four compare+select pairs (the compiler has 7 predicate registers for 8
 * particles x 7 conditions).  A NaN difference is dropped by min, which can only ADD candidates; the exact path
 * rejects them (the reference's x>0 && y>0 && z>0 is false for a NaN). */
template <class Args>
__device__ __forceinline__ bool k1_prefilter(float x, float y, float z, const Args& a) {
    const float mn = fminf(x, fminf(y, z));
    const float r2 = fmaf(z, z, fmaf(y, y, x * x));
    const float z2 = z * z;
    const float d1 = fmaf(a.chi2, r2, -z2);  /* > 0  <=>  z^2 < chi2 r^2 */
    const float d2 = fmaf(a.nclo2, r2, z2);  /* > 0  <=>  z^2 > clo2 r^2   (nclo2 = -clo2) */
    const float d3 = fmaf(a.ntlo, x, y);     /* > 0  <=>  y > tlo x        (ntlo = -tlo)   */
    const float d4 = fmaf(a.thi, x, -y);     /* > 0  <=>  y < thi x */
    const float score = fminf(fminf(d1, d2), fminf(d3, d4));
    /* inside [2^-30, 2^40] nothing above can overflow or lose relative precision; outside, the exact path decides */
    const bool sane = (mn >= CC_SANE_LO) & (r2 <= 1.2089258e24f);
    return sane ? (score > 0.0f) : (mn > 0.0f); /* mn > 0 is the octant test, processLC.cpp:279 */
}

constexpr int TSCAN_THREADS = 256;
constexpr int TSCAN_ITEMS = 16; /* tiles per thread */
constexpr int TSCAN_BLOCK = TSCAN_THREADS * TSCAN_ITEMS;
constexpr unsigned long long TS_AGG = 1ull << 62;
constexpr unsigned long long TS_PREFIX = 2ull << 62;
constexpr unsigned long long TS_MASK = (1ull << 62) - 1;

/* ================================================================================================
 * K1  use case 1, three phases; the unit of ORDER is a segment of K1S_TILES consecutive 256-particle tiles that one
 *     warp streams front to back (segments are dealt round-robin to the warps)
 *
 *   k1s_scan     the bandwidth-bound pass.  Per tile: coalesced float4 loads of x,y,z -> branch-free pre-filter ->
 *                candidates appended in input order to the warp's shared-memory ring (position + index).  Whenever 32
 *                are queued the OLDEST 32 take the exact reference arithmetic, one per lane -- always a full warp,
 *                whatever the tile boundaries (the exact path costs the same ~210 warp instructions for 1 lane as for
 *                32).  Survivors are appended as (index, theta, phi, x, y, z) records to the WARP's private region of a
 *                temp buffer through a cursor the warp owns (no atomics); the ring is first-in-first-out, so the records
 *                of a segment are consecutive there and in input order.  Segment boundaries are found by comparing the
 *                segment of neighbouring lanes: the lane that holds the first candidate of a segment closes the
 *                previous one: seg_info[segment] = (first record in the region, survivor count).  No block barrier,
 *                no inter-warp wait.
 *   k1s_seg_scan exclusive scan of the segment counts (one 8-byte word per 2048 particles): single pass, decoupled
 *                look-back across blocks.
 *   k1s_gather   a CTA owns K1G_SEGS consecutive segments = a contiguous block of OUTPUT rows; thread i of the block
 *                fetches the record of row i (consecutive inside a segment: coalesced), the survivor's 32-byte payload,
 *                and writes row i of all 12 columns (coalesced; processLC.cpp:291-307).  k1s_gather_bulk is the same
 *                kernel with the rows staged in shared memory and written by the bulk-copy engine
 *                (cp.async.bulk shared -> global).
 *
 * History (profiles/r01, profiles/r02): a first version chained the look-back through the particle tiles themselves and
 * was capped at 14 % of the HBM bandwidth by the chain; round 1 ordered by TILE (per-CTA record regions, a scan over
 * 2 M tile counts, one row index per record): at 4.5 % survivors it spent 525 warp instructions per tile where the
 * arithmetic needs ~200 (a second, almost empty round of the exact path per drain, per-tile bookkeeping).
 * ================================================================================================ */
constexpr int K1S_TILES = 8;                                  /* tiles per segment */
constexpr int K1S_SEG = K1S_TILES * SCAN_TILE;                /* particles per segment (2048) */
constexpr int K1S_SEG_SHIFT = 11;
static_assert((1 << K1S_SEG_SHIFT) == K1S_SEG, "segment of a particle = index >> K1S_SEG_SHIFT");
constexpr int K1S_QCAP = 256;                                 /* ring entries per warp: < 32 waiting + up to 128 per half tile */
static_assert((K1S_QCAP & (K1S_QCAP - 1)) == 0 && K1S_QCAP >= 31 + SCAN_TILE / 2, "ring size");

struct SegScanArgs {
    const float *x, *y, *z;
    long long n;                      /* particles in this chunk */
    float th_lo, th_hi, ph_lo, ph_hi; /* exact, strict (processLC.cpp:287-288) */
    float chi2, nclo2, ntlo, thi;     /* pre-filter (prefilter.h) */
    unsigned int* t_idx;              /* records: particle index inside the chunk, theta, phi, position */
    float *t_th, *t_ph, *t_x, *t_y, *t_z;
    unsigned int region_cap;          /* records per WARP region */
    unsigned long long* seg_info;     /* [nseg] (first record inside the warp's region << 32) | survivors */
    unsigned long long* result;       /* [1] += pre-filter candidates, [3] = max records any warp wanted */
    unsigned int nseg, ntiles;
    int prefetch;                     /* L2-prefetch the warp's next tile while this one is worked on */
};
/* The ring keeps the candidate's position, not only its index: reading x,y,z again in the drain was measured (an
 * index-only ring, 8 KB instead of 32 KB per CTA): the streaming loads are evict-first, the re-reads missed the L2 and
 * cost 4 GB of extra DRAM traffic on 500 M particles at 4.5 % survivors (profiles/r02). */
struct __align__(16) K1SRing {
    float4 e[K1S_QCAP];               /* x, y, z, particle index inside the chunk (bits): one 16-byte store per push */
};
struct K1SState {                     /* warp-uniform, in shared memory so that the drain can be a real call */
    unsigned int head, count;         /* ring */
    unsigned int cursor;              /* records this warp has appended */
    unsigned int open_seg, open_start;/* segment whose records are being appended, and its first record */
    unsigned int n_cand;              /* pre-filter candidates (statistics); < 2^32 per warp */
};
constexpr unsigned int K1S_NONE = 0xffffffffu;

__device__ __forceinline__ bool k1s_exact(const SegScanArgs& a, float x, float y, float z, float& th, float& ph) {
    if (!(x > 0.0f && y > 0.0f && z > 0.0f)) return false; /* :279 */
    th = ccref::theta_arcsec(x, y, z);                       /* :282-283 */
    ph = ccref::phi_arcsec_plain(x, y);                      /* :284 */
    return th > a.th_lo && th < a.th_hi && ph > a.ph_lo && ph < a.ph_hi; /* :287-288 */
}

/* warp-collective: the oldest min(count, 32) candidates of the ring take the exact path, one per lane */
__device__ __noinline__ void k1s_drain(const SegScanArgs* __restrict__ ap, K1SRing* __restrict__ rp, K1SState* __restrict__ sp,
                                       unsigned long long region0) {
    const SegScanArgs& a = *ap;
    const int lane = threadIdx.x & 31;
    const unsigned int lane_lt = (1u << lane) - 1u;
    __syncwarp();
    const unsigned int head = sp->head, cnt = min(sp->count, 32u), cursor = sp->cursor;
    const unsigned int open_seg = sp->open_seg, open_start = sp->open_start;
    const bool live = (unsigned int)lane < cnt;
    const unsigned int e = (head + lane) & (K1S_QCAP - 1);
    float x = 0.0f, y = 0.0f, z = 0.0f, th = 0.0f, ph = 0.0f;
    unsigned int idx = 0;
    bool pass = false;
    if (live) {
        const float4 c4 = rp->e[e];
        x = c4.x; y = c4.y; z = c4.z; idx = __float_as_uint(c4.w);
        pass = k1s_exact(a, x, y, z, th, ph);
    }
    const unsigned int B = __ballot_sync(0xffffffffu, pass);
    const unsigned int my_seg = idx >> K1S_SEG_SHIFT;
    unsigned int prev_seg = __shfl_up_sync(0xffffffffu, my_seg, 1);
    if (lane == 0) prev_seg = open_seg;
    const bool is_first = live && my_seg != prev_seg; /* first candidate of its segment */
    const unsigned int F = __ballot_sync(0xffffffffu, is_first);
    const unsigned int my_start = cursor + __popc(B & lane_lt);
    if (is_first && prev_seg != K1S_NONE) { /* close the previous segment: its records end where mine begin */
        const unsigned int pl = F & lane_lt;
        const unsigned int ps = pl ? cursor + __popc(B & ((1u << (31 - __clz(pl))) - 1u)) : open_start;
        a.seg_info[prev_seg] = ((unsigned long long)ps << 32) | (unsigned long long)(my_start - ps);
    }
    if (pass && my_start < a.region_cap) {
        const unsigned long long s = region0 + my_start;
        a.t_idx[s] = idx; a.t_th[s] = th; a.t_ph[s] = ph;
        a.t_x[s] = x; a.t_y[s] = y; a.t_z[s] = z;
    }
    if (F) { /* warp-uniform */
        const int last = 31 - __clz(F);
        const unsigned int seg_l = __shfl_sync(0xffffffffu, my_seg, last);
        if (lane == 0) { sp->open_seg = seg_l; sp->open_start = cursor + __popc(B & ((1u << last) - 1u)); }
    }
    if (lane == 0) { sp->head = head + cnt; sp->count = sp->count - cnt; sp->cursor = cursor + __popc(B); }
    __syncwarp();
}

template <bool VEC>
__global__ void __launch_bounds__(SCAN_THREADS, 4) k1s_scan(const __grid_constant__ SegScanArgs a) {
    __shared__ K1SRing s_ring[SCAN_WARPS];
    __shared__ K1SState s_state[SCAN_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K1SRing& q = s_ring[warp];
    K1SState& st = s_state[warp];
    const unsigned int total_warps = gridDim.x * SCAN_WARPS, gwarp = blockIdx.x * SCAN_WARPS + warp;
    if (lane == 0) { st.head = 0; st.count = 0; st.cursor = 0; st.open_seg = K1S_NONE; st.open_start = 0; st.n_cand = 0; }
    __syncwarp();
    unsigned int qcount = 0, qtail = 0; /* warp-uniform copies: entries queued, next ring position */
    auto drain_to_31 = [&]() {
        while (qcount >= 32u) { k1s_drain(&a, &q, &st, (unsigned long long)gwarp * a.region_cap); qcount -= 32u; }
    };

    for (unsigned int seg = gwarp; seg < a.nseg; seg += total_warps) {
        /* a segment without survivors keeps this 0; otherwise the lane that finds the segment's end overwrites it
         * (k1s_drain; the __syncwarp()s in between order the two stores) */
        if (lane == 0) a.seg_info[seg] = 0ull;
        const unsigned int tile_end = min(a.ntiles, (seg + 1) * K1S_TILES);
        for (unsigned int tile = seg * K1S_TILES; tile < tile_end; ++tile) {
            drain_to_31();
            const long long base = (long long)tile * SCAN_TILE;
            float px[8], py[8], pz[8];
            const unsigned int valid = load_tile<VEC>(a.x, a.y, a.z, base, a.n, lane, px, py, pz);
            if (a.prefetch) { /* the tile this warp loads next: the following one, or the first of its next segment */
                const unsigned int nxt = tile + 1 < tile_end ? tile + 1 : (seg + total_warps) * K1S_TILES;
                prefetch_tile(a.x, a.y, a.z, (long long)nxt * SCAN_TILE, a.n, lane);
            }
            unsigned int m = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) m |= (unsigned)k1_prefilter(px[j], py[j], pz[j], a) << j;
            m &= valid;
            if (!__any_sync(0xffffffffu, m != 0)) continue;
            /* candidates in input order: chunk A (elements 0-3 of every lane), then chunk B */
            const unsigned int cnt = __popc(m & 0xfu) | (__popc(m >> 4) << 16);
            unsigned int inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const unsigned int total = __shfl_sync(0xffffffffu, inc, 31);
            const unsigned int ex = inc - cnt;
            /* a tile can carry up to 256 candidates, the ring has room for 225: push it half by half */
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const unsigned int add = half ? (total >> 16) : (total & 0xffffu);
                if (half == 1 && qcount + add > (unsigned)K1S_QCAP) drain_to_31();
                unsigned int pos = qtail + (half ? (ex >> 16) : (ex & 0xffffu));
                /* straight-line, predicated: one 16-byte shared-memory store per candidate (a loop over the set bits
                 * with register selects was measured at 195 warp instructions per tile, profiles/r02) */
#pragma unroll
                for (int j = 4 * half; j < 4 * half + 4; ++j)
                    if (m & (1u << j)) {
                        q.e[pos & (K1S_QCAP - 1)] = make_float4(px[j], py[j], pz[j], __uint_as_float((unsigned int)base + tile_local(lane, j)));
                        ++pos;
                    }
                qtail += add;
                qcount += add;
                if (lane == 0) { st.count = qcount; st.n_cand += add; }
                __syncwarp();
            }
        }
    }
    while (qcount) { /* the last, partial round */
        k1s_drain(&a, &q, &st, (unsigned long long)gwarp * a.region_cap);
        qcount -= min(qcount, 32u);
    }
    if (lane == 0) {
        if (st.open_seg != K1S_NONE)
            a.seg_info[st.open_seg] = ((unsigned long long)st.open_start << 32) | (unsigned long long)(st.cursor - st.open_start);
        atomicMax(a.result + 3, (unsigned long long)st.cursor);
        if (st.n_cand) atomicAdd(a.result + 1, (unsigned long long)st.n_cand);
    }
}

/* ---- phase B: exclusive scan of the segment counts (single pass, decoupled look-back across blocks) ---- */
struct SegPrefixArgs {
    const unsigned long long* seg_info; /* [nseg] (position << 32) | count */
    unsigned int* seg_row;              /* [nseg] out: output row of the segment's first record, inside the chunk */
    unsigned long long* block_state;    /* [nblocks] zeroed before launch */
    unsigned int* ticket;               /* zeroed before launch */
    unsigned long long* result;         /* [0] survivors so far (in: carry, out: carry + chunk total), [2] = carry */
    unsigned long long* host_result;    /* pinned host memory (or NULL): result[0..3] of the step's last chunk, so that no
                                           copy command has to follow the gather */
    unsigned int nseg, nblocks;
};
__global__ void __launch_bounds__(TSCAN_THREADS) k1s_seg_scan(const SegPrefixArgs a) {
    __shared__ unsigned int s_warp[TSCAN_THREADS / 32];
    __shared__ unsigned int s_block;
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_block = atomicAdd(a.ticket, 1u); /* blocks enter the chain in ticket order: no deadlock */
    __syncthreads();
    const unsigned int blk = s_block;
    const unsigned int first = blk * TSCAN_BLOCK + tid * TSCAN_ITEMS;
    unsigned int c[TSCAN_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < TSCAN_ITEMS; ++i) {
        c[i] = (first + i < a.nseg) ? (unsigned int)(a.seg_info[first + i] & 0xffffffffull) : 0u;
        sum += c[i];
    }
    unsigned int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned int before = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < TSCAN_THREADS / 32; ++w) {
        const unsigned int t = s_warp[w];
        if (w < warp) before += t;
        block_total += t;
    }
    if (warp == 0) {
        unsigned long long exclusive = 0;
        if (blk == 0) {
            if (lane == 0) st_relaxed_u64(a.block_state, TS_PREFIX | block_total);
        } else {
            if (lane == 0) st_relaxed_u64(a.block_state + blk, TS_AGG | block_total);
            long long look = (long long)blk - 1;
            while (true) {
                const long long idx = look - lane;
                unsigned long long stt = TS_PREFIX;
                if (idx >= 0) {
                    stt = ld_relaxed_u64(a.block_state + idx);
                    while ((stt >> 62) == 0) stt = ld_relaxed_u64(a.block_state + idx);
                }
                const unsigned int pmask = __ballot_sync(0xffffffffu, (stt >> 62) == 2);
                const int firstp = pmask ? (__ffs(pmask) - 1) : 31;
                unsigned long long v = (lane <= firstp) ? (stt & TS_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                exclusive += v;
                if (pmask) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_u64(a.block_state + blk, TS_PREFIX | (exclusive + block_total));
        }
        if (lane == 0) {
            s_excl = exclusive;
            if (blk == a.nblocks - 1) { /* chunk total known: roll the carry */
                const unsigned long long carry = a.result[0];
                a.result[2] = carry;
                a.result[0] = carry + exclusive + block_total;
                if (a.host_result) { /* [1] candidates and [3] the largest record region are final since k1s_scan ended */
                    a.host_result[0] = carry + exclusive + block_total;
                    a.host_result[1] = a.result[1];
                    a.host_result[2] = carry;
                    a.host_result[3] = a.result[3]; /* read by the host after the stream has drained: no fence */
                }
            }
        }
    }
    __syncthreads();
    unsigned int run = (unsigned int)s_excl + before + inc - sum;
#pragma unroll
    for (int i = 0; i < TSCAN_ITEMS; ++i) {
        if (first + i < a.nseg) a.seg_row[first + i] = run;
        run += c[i];
    }
}

/* ---- phase C: a CTA owns K1G_SEGS consecutive segments = a contiguous block of output rows ---- */
constexpr int K1G_SEGS = 32;
constexpr int K1G_THREADS = 256;
struct SegGatherArgs {
    ColsIn in;   /* columns other than x,y,z: pointers already offset to the chunk start */
    ColsOut out;
    const unsigned int* t_idx;
    const float *t_th, *t_ph, *t_x, *t_y, *t_z;
    const unsigned long long* seg_info;
    const unsigned int* seg_row;
    const unsigned long long* result; /* [2] = carry: rows written by earlier chunks of a streamed step */
    const Payload* payload;           /* interleaved gather columns of a resident step, or NULL */
    unsigned int region_cap, total_warps, nseg;
    long long cap;                    /* rows available in `out` */
    long long index_base;             /* global index of the chunk's first particle */
};

struct K1GRow { /* everything one output row holds */
    float th, ph, x, y, z, vx, vy, vz, red;
    long long id, index;
    int rot, rep;
};
/* record + payload of local row i of the CTA's block (s_pref/s_pos: the block's segment table in shared memory) */
__device__ __forceinline__ bool k1g_fetch(const SegGatherArgs& a, unsigned int g0, const unsigned int* s_pref,
                                          const unsigned int* s_pos, unsigned int i, K1GRow& r) {
    unsigned int s = 0; /* largest s with s_pref[s] <= i */
#pragma unroll
    for (int d = K1G_SEGS / 2; d > 0; d >>= 1)
        if (s_pref[s + d] <= i) s += d;
    const unsigned int k = s_pos[s] + (i - s_pref[s]);
    if (k >= a.region_cap) return false; /* the warp's region overflowed: the host re-runs with larger regions */
    const unsigned long long slot = (unsigned long long)((g0 + s) % a.total_warps) * a.region_cap + k;
    const long long g = a.t_idx[slot];
    r.th = a.t_th[slot]; r.ph = a.t_ph[slot];
    r.x = a.t_x[slot]; r.y = a.t_y[slot]; r.z = a.t_z[slot];
    r.index = a.index_base + g;
    if (a.payload) {
        const Payload p = load_payload(a.payload + g);
        r.vx = p.vx; r.vy = p.vy; r.vz = p.vz; r.red = a_to_z_dev(p.a);
        r.id = p.id; r.rot = p.rot; r.rep = p.rep;
    } else if (a.in.a) { /* NULL: a streamed step whose other columns the host gathers itself (capi.cu fetch_rows) */
        r.vx = a.in.vx[g]; r.vy = a.in.vy[g]; r.vz = a.in.vz[g];
        r.red = a_to_z_dev(a.in.a[g]);
        r.id = a.in.id[g]; r.rot = a.in.rot[g]; r.rep = a.in.rep[g];
    }
    return true;
}
/* the CTA's segment table; returns the number of rows of the block, *row0 = its first row inside the chunk */
__device__ __forceinline__ unsigned int k1g_table(const SegGatherArgs& a, unsigned int g0, unsigned int* s_pref,
                                                  unsigned int* s_pos, unsigned int* row0) {
    if (threadIdx.x < 32) {
        const unsigned int seg = g0 + threadIdx.x;
        const unsigned long long info = seg < a.nseg ? a.seg_info[seg] : 0ull;
        const unsigned int c = (unsigned int)(info & 0xffffffffull);
        unsigned int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (threadIdx.x >= (unsigned)d) inc += t;
        }
        s_pref[threadIdx.x] = inc - c;
        s_pos[threadIdx.x] = (unsigned int)(info >> 32);
        if (threadIdx.x == 31) { s_pref[32] = inc; s_pref[33] = a.seg_row[g0]; }
    }
    __syncthreads();
    *row0 = s_pref[33];
    return s_pref[32];
}
static_assert(K1G_SEGS == 32, "k1g_table scans the block's segments with one warp");

__global__ void __launch_bounds__(K1G_THREADS) k1s_gather(const SegGatherArgs a) {
    __shared__ unsigned int s_pref[K1G_SEGS + 2], s_pos[K1G_SEGS];
    const unsigned int g0 = blockIdx.x * K1G_SEGS;
    unsigned int row0;
    const unsigned int total = k1g_table(a, g0, s_pref, s_pos, &row0);
    const long long first = (long long)a.result[2] + row0;
    const bool full = a.payload || a.in.a;
    for (unsigned int i = threadIdx.x; i < total; i += K1G_THREADS) {
        const long long r = first + i;
        K1GRow v;
        if (r >= a.cap || !k1g_fetch(a, g0, s_pref, s_pos, i, v)) continue;
        a.out.theta[r] = v.th; a.out.phi[r] = v.ph;
        a.out.x[r] = v.x; a.out.y[r] = v.y; a.out.z[r] = v.z;
        if (full) {
            a.out.vx[r] = v.vx; a.out.vy[r] = v.vy; a.out.vz[r] = v.vz; a.out.redshift[r] = v.red;
            a.out.id[r] = v.id; a.out.rot[r] = v.rot; a.out.rep[r] = v.rep;
        }
        if (a.out.index) a.out.index[r] = v.index;
    }
}

/* The same gather with the output rows staged in shared memory, column by column, and written by the bulk-copy engine
 * (cp.async.bulk.global.shared::cta; north star: "the gather writes output columns with TMA bulk stores").  A staged
 * block covers K1GB_ROWS consecutive output rows starting at a multiple of 4 rows, so that every column's run starts on
 * a 16-byte boundary both in shared memory and in HBM; the rows in front of / behind the 16-byte aligned interior of
 * the CTA's range (< 4 each) are stored by threads. */
constexpr int K1GB_ROWS = 1024;
constexpr int K1GB_NCOL4 = 11; /* theta phi x y z vx vy vz redshift rot rep */
constexpr size_t K1GB_SMEM = (size_t)K1GB_ROWS * (K1GB_NCOL4 * 4 + 2 * 8);
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"((unsigned int)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(K1G_THREADS) k1s_gather_bulk(const SegGatherArgs a) {
    extern __shared__ __align__(128) unsigned char s_stage[];
    __shared__ unsigned int s_pref[K1G_SEGS + 2], s_pos[K1G_SEGS];
    float* s4 = reinterpret_cast<float*>(s_stage);                                         /* [11][ROWS] */
    long long* s8 = reinterpret_cast<long long*>(s_stage + (size_t)K1GB_ROWS * K1GB_NCOL4 * 4); /* [2][ROWS] */
    const unsigned int g0 = blockIdx.x * K1G_SEGS;
    unsigned int row0;
    const unsigned int total = k1g_table(a, g0, s_pref, s_pos, &row0);
    const long long lo = (long long)a.result[2] + row0, hi = min(lo + (long long)total, a.cap);
    const bool full = a.payload || a.in.a;
    float* const o4[K1GB_NCOL4] = {a.out.theta, a.out.phi, a.out.x, a.out.y, a.out.z, a.out.vx, a.out.vy, a.out.vz,
                                   a.out.redshift, reinterpret_cast<float*>(a.out.rot), reinterpret_cast<float*>(a.out.rep)};
    long long* const o8[2] = {a.out.id, a.out.index};
    for (long long A = lo & ~3ll; A < hi; A += K1GB_ROWS) { /* staged block = rows [A, A + ROWS) of the CTA's range */
        const long long b0 = max(A, lo), b1 = min(A + K1GB_ROWS, hi);
        for (long long r = b0 + threadIdx.x; r < b1; r += K1G_THREADS) {
            K1GRow v = {};
            const unsigned int j = (unsigned int)(r - A);
            if (!k1g_fetch(a, g0, s_pref, s_pos, (unsigned int)(r - lo), v)) continue; /* overflow: the host re-runs */
            s4[0 * K1GB_ROWS + j] = v.th; s4[1 * K1GB_ROWS + j] = v.ph;
            s4[2 * K1GB_ROWS + j] = v.x; s4[3 * K1GB_ROWS + j] = v.y; s4[4 * K1GB_ROWS + j] = v.z;
            s4[5 * K1GB_ROWS + j] = v.vx; s4[6 * K1GB_ROWS + j] = v.vy; s4[7 * K1GB_ROWS + j] = v.vz;
            s4[8 * K1GB_ROWS + j] = v.red;
            s4[9 * K1GB_ROWS + j] = __int_as_float(v.rot); s4[10 * K1GB_ROWS + j] = __int_as_float(v.rep);
            s8[j] = v.id; s8[K1GB_ROWS + j] = v.index;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy writes -> visible to the bulk engine */
        __syncthreads();
        const long long i0 = (b0 + 3) & ~3ll, i1 = b1 & ~3ll; /* 16-byte aligned interior */
        if (threadIdx.x == 0 && i1 > i0) {
            const unsigned int j0 = (unsigned int)(i0 - A), nr = (unsigned int)(i1 - i0);
#pragma unroll
            for (int c = 0; c < K1GB_NCOL4; ++c)
                if (c < 5 || full) bulk_s2g(o4[c] + i0, s4 + (size_t)c * K1GB_ROWS + j0, nr * 4u);
            if (full) bulk_s2g(o8[0] + i0, s8 + j0, nr * 8u);
            if (o8[1]) bulk_s2g(o8[1] + i0, s8 + K1GB_ROWS + j0, nr * 8u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        { /* head and tail rows (at most 3 + 3), by threads */
            const long long h1 = min(i0, b1), t0 = max(i1, h1);
            const long long nh = h1 - b0, nt = b1 - t0;
            if ((long long)threadIdx.x < nh + nt) {
                const long long r = (long long)threadIdx.x < nh ? b0 + threadIdx.x : t0 + (threadIdx.x - nh);
                const unsigned int j = (unsigned int)(r - A);
#pragma unroll
                for (int c = 0; c < K1GB_NCOL4; ++c)
                    if (c < 5 || full) o4[c][r] = s4[(size_t)c * K1GB_ROWS + j];
                if (full) o8[0][r] = s8[j];
                if (o8[1]) o8[1][r] = s8[K1GB_ROWS + j];
            }
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); /* staging may be refilled */
        __syncthreads();
    }
}

/* ================================================================================================
 * K2  use case 2: one pass, a batch of halos, warp-autonomous tiles
 *
 * Per particle (once): r^2, cos(theta_u) ~ z * rsqrt(r^2), tan(phi_u) ~ y * rcp(x), range flag.
 * Per particle and halo: four compares against the halo's widened (cos, tan) window.  Whatever passes is pushed
 * as a (particle, halo) pair into the warp's shared-memory queue; after the halo loop the warp drains the queue
 * with all lanes busy: exact un-rotated angles -> rough test -> R.v -> exact rotated angles -> fine test ->
 * survivor record (order is restored by k3_*).
 * ================================================================================================ */
constexpr int K2_MAX_HALOS = 1024; /* halos per launch (pre-filter parameters live in shared memory) */
constexpr int K2_QCAP = SCAN_TILE + 32; /* queued (particle, halo) pairs per warp: drained once more than 32 are
                                           queued, and one halo iteration can add a whole tile (small on purpose:
                                           shared memory competes with the L1 that buffers the streaming loads) */

struct HaloDev { /* exact parameters of one halo */
    float R[9];
    float th_lo, th_hi, ph_lo, ph_hi;     /* fine, strict (processLC.cpp:1439-1440) */
    float rth_lo, rth_hi, rph_lo, rph_hi; /* rough: lo <= theta_u < hi (:1334-1337), lo < phi_u < hi (:1400) */
    float pad[3];
    /* cell-index kernel, second filter (prefilter.h): */
    float in_chi, in_clo, in_slo, in_shi; /* inside this (cos theta_u, sin phi_u) window the rough test is certain to pass */
    float f_chi, f_clo, f_tlo, f_thi;     /* outside this (cos theta, tan phi) window of the ROTATED position the fine test is certain to fail */
};
struct __align__(16) Rec {  /* one survivor, 32 B */
    unsigned long long key; /* (bits of theta_u) << 32 | particle index in the step: the reference's sort order */
    unsigned int halo;
    float th, ph;           /* rotated, halo-centric angles (processLC.cpp:1446-1447) */
    float x, y, z;          /* the position (in registers when the record is made: saves three gathers later) */
};
struct HaloArgs {
    const float *x, *y, *z;
    long long n;           /* particles to scan (after the n_limit clamp) */
    long long local_base;  /* index of x[0] inside the step (non-zero for chunks of a streamed step) */
    int nhalos;            /* in this launch */
    int halo0;             /* index of the first halo of this launch */
    const float4* pre;     /* [nhalos] {chi, clo, tlo, thi}: cos(theta_u) in [clo,chi], y/x in [tlo,thi] */
    const HaloDev* halos;  /* [nhalos] */
    Rec* recs;
    unsigned long long rec_cap;
    unsigned long long* rec_count;   /* appended records (may exceed rec_cap: overflow is detected on the host) */
    unsigned long long* counts;      /* [2*H_total]: survivors per halo, then rough survivors per halo */
    unsigned long long* n_cand;      /* pre-filter candidates (stat) */
    unsigned long long* nan_count;   /* scanned particles whose un-rotated theta is NaN (stat; capi.cu cc_last_stats_ex) */
    int total_halos;
    unsigned int ntiles;
    int prefetch;                    /* L2-prefetch the warp's next tile while this one is worked on */
};

/* exact path for one (particle, halo) pair */
__device__ __forceinline__ void k2_exact(const HaloArgs& a, int h, float x, float y, float z, unsigned int local_n) {
    const HaloDev* __restrict__ Hp = a.halos + h;
    const unsigned int hg = (unsigned int)(a.halo0 + h);
    const float thu = ccref::theta_arcsec(x, y, z);       /* processLC.cpp:899-900 */
    const float phu = ccref::phi_arcsec_guarded(x, y);    /* :903-908 */
    /* A NaN un-rotated theta (NaN coordinate, z = +-Inf, the origin) breaks the ordering contract of the reference's
     * stable_sort / lower_bound on that key (:1214-1221, :1334-1337); such positions are out of the pre-filter's range,
     * so they are paired with EVERY halo and reach this point: count each once (first halo), cc_last_stats_ex */
    if (thu != thu && hg == 0u) atomicAdd(a.nan_count, 1ull);
    if (!(thu >= Hp->rth_lo && thu < Hp->rth_hi)) return; /* :1334-1340 (two lower_bounds) */
    if (!(phu > Hp->rph_lo && phu < Hp->rph_hi)) return;  /* :1400 */
    atomicAdd(a.counts + a.total_halos + hg, 1ull);       /* rough_cutout_size, :1436 */
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Hp->R[i];
    float rx, ry, rz;
    ccref::mat_vec(R, x, y, z, rx, ry, rz);               /* :1419-1422 */
    const float th = ccref::theta_arcsec(rx, ry, rz);     /* :1425-1426 */
    const float ph = ccref::phi_arcsec_guarded(rx, ry);   /* :1430-1435 */
    if (th > Hp->th_lo && th < Hp->th_hi && ph > Hp->ph_lo && ph < Hp->ph_hi) { /* :1439-1440 */
        atomicAdd(a.counts + hg, 1ull);
        const unsigned long long slot = atomicAdd(a.rec_count, 1ull);
        if (slot < a.rec_cap) {
            Rec r;
            r.key = ((unsigned long long)(unsigned int)__float_as_int(thu) << 32) | local_n;
            r.halo = hg;
            r.th = th;
            r.ph = ph;
            r.x = x; r.y = y; r.z = z;
            a.recs[slot] = r;
        }
    }
}

/* the same exact path for up to 32 pairs at once, one per live lane (the cell-index kernel, where a third of the
 * instructions are spent here); EVERY lane of the warp must call.  The
 * halo's parameters come in as five 16-byte loads issued together (one L2 round trip instead of three dependent
 * ones), and the survivors of the warp claim their record slots with one atomic. */
__device__ __forceinline__ void k2_exact_warp(const HaloArgs& a, bool live, int h, float x, float y, float z,
                                              unsigned int local_n, int lane, unsigned int* s_rough = nullptr,
                                              unsigned int* s_fine = nullptr) {
    static_assert(sizeof(HaloDev) == 112 && offsetof(HaloDev, rph_hi) == 64, "five float4 loads below");
    const float4* __restrict__ Hq = reinterpret_cast<const float4*>(a.halos + (live ? h : 0));
    const float4 q0 = __ldg(Hq), q1 = __ldg(Hq + 1), q2 = __ldg(Hq + 2), q3 = __ldg(Hq + 3), q4 = __ldg(Hq + 4);
    const float th_lo = q2.y, th_hi = q2.z, ph_lo = q2.w, ph_hi = q3.x; /* fine, strict (processLC.cpp:1439-1440) */
    const float rth_lo = q3.y, rth_hi = q3.z, rph_lo = q3.w, rph_hi = q4.x;
    const unsigned int hg = (unsigned int)(a.halo0 + h);
    const float thu = ccref::theta_arcsec(x, y, z);       /* processLC.cpp:899-900 */
    const float phu = ccref::phi_arcsec_guarded(x, y);    /* :903-908 */
    if (live && thu != thu && hg == 0u) atomicAdd(a.nan_count, 1ull); /* see k2_exact */
    const bool rough = live && (thu >= rth_lo && thu < rth_hi) /* :1334-1340 (two lower_bounds) */
                            && (phu > rph_lo && phu < rph_hi); /* :1400 */
    if (!__any_sync(0xffffffffu, rough)) return;
    if (rough) { /* rough_cutout_size, :1436 (per-CTA shared-memory counters when the caller keeps them) */
        if (s_rough) atomicAdd(&s_rough[h], 1u);
        else atomicAdd(a.counts + a.total_halos + hg, 1ull);
    }
    const float R[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
    float rx, ry, rz;
    ccref::mat_vec(R, x, y, z, rx, ry, rz);               /* :1419-1422 */
    const float th = ccref::theta_arcsec(rx, ry, rz);     /* :1425-1426 */
    const float ph = ccref::phi_arcsec_guarded(rx, ry);   /* :1430-1435 */
    const bool fine = rough && th > th_lo && th < th_hi && ph > ph_lo && ph < ph_hi; /* :1439-1440 */
    const unsigned int m = __ballot_sync(0xffffffffu, fine);
    if (m == 0) return;
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.rec_count, (unsigned long long)__popc(m));
    slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane) - 1u));
    if (fine) {
        if (s_fine) atomicAdd(&s_fine[h], 1u);
        else atomicAdd(a.counts + hg, 1ull);
        if (slot < a.rec_cap) {
            Rec r;
            r.key = ((unsigned long long)(unsigned int)__float_as_int(thu) << 32) | local_n;
            r.halo = hg;
            r.th = th;
            r.ph = ph;
            r.x = x; r.y = y; r.z = z;
            a.recs[slot] = r;
        }
    }
}

struct K2WarpQueue {
    unsigned int particle[K2_QCAP]; /* index inside the chunk */
    unsigned short halo[K2_QCAP];
};

/* all lanes of the warp: run the queued (particle, halo) pairs through the exact path */
__device__ __forceinline__ void k2_drain(const HaloArgs& a, const K2WarpQueue& q, unsigned int count, int lane) {
    __syncwarp();
    for (unsigned int j = lane; j < count; j += 32) {
        const unsigned int g = q.particle[j];
        k2_exact(a, q.halo[j], a.x[g], a.y[g], a.z[g], (unsigned int)(a.local_base + g));
    }
    __syncwarp();
}

template <bool VEC>
__global__ void __launch_bounds__(SCAN_THREADS, 4) k2_cutout_halo(const HaloArgs a) {
    extern __shared__ float4 s_pre[]; /* [nhalos] */
    __shared__ K2WarpQueue s_q[SCAN_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K2WarpQueue& q = s_q[warp];
    for (int h = threadIdx.x; h < a.nhalos; h += SCAN_THREADS) s_pre[h] = a.pre[h];
    __syncthreads();
    const unsigned int total_warps = gridDim.x * SCAN_WARPS;
    unsigned long long n_cand = 0;
    unsigned int qcount = 0; /* warp-uniform */

    for (unsigned int tile = blockIdx.x * SCAN_WARPS + warp; tile < a.ntiles; tile += total_warps) {
        const long long base = (long long)tile * SCAN_TILE;
        float px[8], py[8], pz[8];
        const unsigned int valid = load_tile<VEC>(a.x, a.y, a.z, base, a.n, lane, px, py, pz);
        if (a.prefetch) prefetch_tile(a.x, a.y, a.z, base + (long long)total_warps * SCAN_TILE, a.n, lane);
        /* per-particle quantities shared by all halos; out-of-range particles are always candidates */
        float pc[8], pt[8];
        unsigned int insane = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float r2 = fmaf(pz[j], pz[j], fmaf(py[j], py[j], px[j] * px[j]));
            /* |x| in range keeps rcp accurate and r^2 away from underflow; r^2 <= 2^80 keeps it finite */
            if (!((fabsf(px[j]) >= CC_SANE_LO) & (r2 <= 1.2089258e24f))) insane |= 1u << j;
            pc[j] = pz[j] * rsqrt_fast(r2);
            pt[j] = py[j] * rcp_fast(px[j]);
        }
        insane &= valid;
        for (int h = 0; h < a.nhalos; ++h) {
            const float4 P = s_pre[h];
            bool mine = insane != 0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                mine |= (pc[j] <= P.x) & (pc[j] >= P.y) & (pt[j] >= P.z) & (pt[j] <= P.w);
            if (__any_sync(0xffffffffu, mine)) { /* rare, warp-uniform: queue the (particle, halo) pairs */
                unsigned int m = insane;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    m |= (unsigned)((pc[j] <= P.x) & (pc[j] >= P.y) & (pt[j] >= P.z) & (pt[j] <= P.w)) << j;
                m &= valid;
                const unsigned int c = __popc(m);
                unsigned int inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += t;
                }
                unsigned int pos = qcount + inc - c;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (m & (1u << j)) {
                        q.particle[pos] = (unsigned int)(base + tile_local(lane, j));
                        q.halo[pos] = (unsigned short)h;
                        ++pos;
                    }
                qcount += __shfl_sync(0xffffffffu, inc, 31);
                if (qcount > K2_QCAP - SCAN_TILE) { /* the next push may add up to a whole tile */
                    k2_drain(a, q, qcount, lane);
                    n_cand += qcount;
                    qcount = 0;
                }
            }
        }
    }
    if (qcount) {
        k2_drain(a, q, qcount, lane);
        n_cand += qcount;
    }
    if (lane == 0 && n_cand) atomicAdd(a.n_cand, n_cand);
}

/* ---- TMA-fed variant of the halo scan -------------------------------------------------------------------------
 * Same arithmetic and queueing as k2_cutout_halo, but a tile's x,y,z reach the SM through the bulk-copy engine
 * (cp.async.bulk global -> shared, completion on an mbarrier) instead of LDG.128 into registers: every warp owns a ring
 * of K2T_STAGES tile buffers and keeps K2T_STAGES-1 tiles in flight at all times, without spending registers or L1
 * lines on them.  Selected with CC_K2_TMA=1 (profiles/r01 has the A/B numbers). */
constexpr int K2T_STAGES = 3;
constexpr int K2T_TILE_BYTES = SCAN_TILE * 4; /* one column of one tile */
struct __align__(128) K2TWarpRing {
    float buf[K2T_STAGES][3][SCAN_TILE];
    unsigned long long bar[K2T_STAGES];
};

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(SCAN_THREADS, 2) k2_cutout_halo_tma(const HaloArgs a) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    K2TWarpRing* s_ring = reinterpret_cast<K2TWarpRing*>(s_dyn);
    float4* s_pre = reinterpret_cast<float4*>(s_dyn + sizeof(K2TWarpRing) * SCAN_WARPS);
    __shared__ K2WarpQueue s_q[SCAN_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K2WarpQueue& q = s_q[warp];
    K2TWarpRing& ring = s_ring[warp];
    for (int h = threadIdx.x; h < a.nhalos; h += SCAN_THREADS) s_pre[h] = a.pre[h];
    if (lane == 0)
        for (int s = 0; s < K2T_STAGES; ++s) mbar_init(&ring.bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const unsigned int total_warps = gridDim.x * SCAN_WARPS;
    const unsigned int first = blockIdx.x * SCAN_WARPS + warp;
    const unsigned int full_tiles = (unsigned int)(a.n / SCAN_TILE); /* tiles the bulk engine can take whole */
    unsigned long long n_cand = 0;
    unsigned int qcount = 0; /* warp-uniform */

    auto issue = [&](unsigned int tile, int stage) { /* lane 0 only */
        const size_t off = (size_t)tile * SCAN_TILE;
        mbar_expect_tx(&ring.bar[stage], 3 * K2T_TILE_BYTES);
        bulk_g2s(ring.buf[stage][0], a.x + off, K2T_TILE_BYTES, &ring.bar[stage]);
        bulk_g2s(ring.buf[stage][1], a.y + off, K2T_TILE_BYTES, &ring.bar[stage]);
        bulk_g2s(ring.buf[stage][2], a.z + off, K2T_TILE_BYTES, &ring.bar[stage]);
    };
    if (lane == 0)
        for (int s = 0; s < K2T_STAGES; ++s) {
            const unsigned long long t = (unsigned long long)first + (unsigned long long)s * total_warps;
            if (t < full_tiles) issue((unsigned int)t, s);
        }

    unsigned int k = 0;
    for (unsigned int tile = first; tile < a.ntiles; tile += total_warps, ++k) {
        const long long base = (long long)tile * SCAN_TILE;
        float px[8], py[8], pz[8];
        unsigned int valid;
        if (tile < full_tiles) {
            const int stage = (int)(k % K2T_STAGES);
            mbar_wait(&ring.bar[stage], (k / K2T_STAGES) & 1u);
            const float4 xa = *reinterpret_cast<const float4*>(&ring.buf[stage][0][4 * lane]);
            const float4 xb = *reinterpret_cast<const float4*>(&ring.buf[stage][0][128 + 4 * lane]);
            const float4 ya = *reinterpret_cast<const float4*>(&ring.buf[stage][1][4 * lane]);
            const float4 yb = *reinterpret_cast<const float4*>(&ring.buf[stage][1][128 + 4 * lane]);
            const float4 za = *reinterpret_cast<const float4*>(&ring.buf[stage][2][4 * lane]);
            const float4 zb = *reinterpret_cast<const float4*>(&ring.buf[stage][2][128 + 4 * lane]);
            px[0] = xa.x; px[1] = xa.y; px[2] = xa.z; px[3] = xa.w; px[4] = xb.x; px[5] = xb.y; px[6] = xb.z; px[7] = xb.w;
            py[0] = ya.x; py[1] = ya.y; py[2] = ya.z; py[3] = ya.w; py[4] = yb.x; py[5] = yb.y; py[6] = yb.z; py[7] = yb.w;
            pz[0] = za.x; pz[1] = za.y; pz[2] = za.z; pz[3] = za.w; pz[4] = zb.x; pz[5] = zb.y; pz[6] = zb.z; pz[7] = zb.w;
            valid = 0xffu;
            __syncwarp(); /* every lane has read the stage: refill it with the tile K2T_STAGES ahead */
            const unsigned long long nxt = (unsigned long long)tile + (unsigned long long)K2T_STAGES * total_warps;
            if (lane == 0 && nxt < full_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue((unsigned int)nxt, stage);
            }
        } else { /* the ragged last tile of the step */
            valid = load_tile<false>(a.x, a.y, a.z, base, a.n, lane, px, py, pz);
        }
        float pc[8], pt[8];
        unsigned int insane = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float r2 = fmaf(pz[j], pz[j], fmaf(py[j], py[j], px[j] * px[j]));
            if (!((fabsf(px[j]) >= CC_SANE_LO) & (r2 <= 1.2089258e24f))) insane |= 1u << j;
            pc[j] = pz[j] * rsqrt_fast(r2);
            pt[j] = py[j] * rcp_fast(px[j]);
        }
        insane &= valid;
        for (int h = 0; h < a.nhalos; ++h) {
            const float4 P = s_pre[h];
            bool mine = insane != 0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                mine |= (pc[j] <= P.x) & (pc[j] >= P.y) & (pt[j] >= P.z) & (pt[j] <= P.w);
            if (__any_sync(0xffffffffu, mine)) {
                unsigned int m = insane;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    m |= (unsigned)((pc[j] <= P.x) & (pc[j] >= P.y) & (pt[j] >= P.z) & (pt[j] <= P.w)) << j;
                m &= valid;
                const unsigned int c = __popc(m);
                unsigned int inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += t;
                }
                unsigned int pos = qcount + inc - c;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (m & (1u << j)) {
                        q.particle[pos] = (unsigned int)(base + tile_local(lane, j));
                        q.halo[pos] = (unsigned short)h;
                        ++pos;
                    }
                qcount += __shfl_sync(0xffffffffu, inc, 31);
                if (qcount > K2_QCAP - SCAN_TILE) {
                    k2_drain(a, q, qcount, lane);
                    n_cand += qcount;
                    qcount = 0;
                }
            }
        }
    }
    if (qcount) {
        k2_drain(a, q, qcount, lane);
        n_cand += qcount;
    }
    if (lane == 0 && n_cand) atomicAdd(a.n_cand, n_cand);
}
constexpr size_t K2T_SMEM_BASE = sizeof(K2TWarpRing) * SCAN_WARPS;

/* ---- many halos: a (cos theta_u, sin phi_u) cell index over the halos of the launch ----------------------------
 * With hundreds of halos per pass the loop "every particle x every halo" is compute-bound (and a 1-D theta index
 * still leaves ~1 % of the halos per particle: profiles/r01).  The host rasterises each halo's widened rough
 * window [clo,chi] x [slo,shi] into a K2_GRID x K2_GRID (2048^2) grid over (c, s) in [-1,1]^2 (+-1 cell of slack) and builds
 * a CSR list of halos per cell, plus a 512 x 512 occupancy bitmap (one bit per 4 x 4 cells).  A particle computes
 * c = z/r and s = sin(phi_u) = y*sign(x)/sqrt(x^2+y^2), is ruled out by the bitmap most of the time, and otherwise
 * looks up its cell (one 8-byte read from an L2-resident table) and owes one test per halo listed there -- for
 * cluster-sized boxes most cells are empty.  This plays the role of the reference's theta-sorted index
 * (processLC.cpp:1214-1221, :1334-1340) without sorting or re-reading 12 B/particle data.
 * Two walkers: ONE kernel (k2_cutout_halo_cells) and a staged form (k2c_sift / k2c_match / k2c_exact). */
#ifndef CC_K2_GRID
#define CC_K2_GRID 2048 /* 1024: 12 % slower on 1000 cluster-sized halos (more pairs per hit); the 32 MB table stays L2-resident */
#endif
constexpr int K2_GRID = CC_K2_GRID;
constexpr int K2_CSHIFT = K2_GRID == 2048 ? 2 : 1;
constexpr int K2_COARSE = K2_GRID >> K2_CSHIFT;              /* occupancy bitmap: 512 x 512 bits, one per 4 x 4 (2 x 2) cells */
static_assert(K2_COARSE == 512, "the bitmap is 32 KB: shared memory of the one-kernel scan, L1 of the staged one");
constexpr int K2_BITWORDS = K2_COARSE * K2_COARSE / 32;

/* (c, s) of the cell index */
__device__ __forceinline__ void k2c_coords(float x, float y, float z, float& c, float& s, bool& sane) {
    const float rxy2 = fmaf(y, y, x * x);
    const float r2 = fmaf(z, z, rxy2);
    /* |x| in range keeps x^2+y^2 away from underflow; r^2 <= 2^80 keeps everything finite */
    sane = (fabsf(x) >= CC_SANE_LO) & (r2 <= 1.2089258e24f);
    c = z * rsqrt_fast(r2);
    /* sin(phi_u) with phi_u = atan(y/x): y / sqrt(x^2+y^2), negated when x < 0 */
    s = __int_as_float(__float_as_int(y * rsqrt_fast(rxy2)) ^ (__float_as_int(x) & 0x80000000));
}

/* ---- many halos, one kernel -------------------------------------------------------------------------------------
 * The cell index walked by ONE kernel: one CTA of 20 warps per SM keeps the occupancy bitmap (32 KB) and the halo
 * windows in shared memory; hits are pushed -- position included -- onto the warp's shared-memory stack, and whenever
 * 32 are stacked the warp pops them, one per lane: the next halo of the cell's list, four compares against that halo's
 * window; a pair that passes goes onto a second stack of tested pairs, a hit with a longer list is pushed back with the
 * list advanced by one; every 32 tested pairs take the second filter and what it cannot decide the exact path.  Tiles
 * are 128 particles, fetched one trip ahead with cp.async.  Nothing is staged in HBM, which is what wins when a large
 * share of the particles hits the bitmap (1000 cluster-sized halos in one pass: 22 %); the staged kernels below win
 * when few do (profiles/r02/README.md).  capi.cu picks per halo set from the hit rate it measured. */
/* one CTA per SM shares the bitmap: WARPS warps (template parameter; 20 at 96 registers by default) */
constexpr int K2C_TILE = 128; /* particles per warp tile */
constexpr int K2C_HCAP = 64;  /* stacked hits per warp (a tile with more is taken in several passes) */
constexpr int K2C_CCAP = 64;  /* stacked tested pairs per warp: < 32 waiting + <= 32 from one round */
struct __align__(16) K2CStacks {
    float4 hp[K2C_HCAP];         /* hit: x, y, z, particle index inside the chunk (bits) */
    uint2 hl[K2C_HCAP];          /*      {next item of the cell's list, items left} */
    float4 cp[K2C_CCAP];         /* tested pair: x, y, z, particle index (bits) */
    unsigned int ch[K2C_CCAP];   /*              halo inside the launch */
    float4 xp[K2C_CCAP];         /* pair that needs the exact arithmetic */
    unsigned int xh[K2C_CCAP];
    float4 next[3][32];          /* x, y, z of the warp's NEXT tile, filled by cp.async while this one is worked on */
};
constexpr size_t k2c_smem(int warps) {
    return (size_t)K2_BITWORDS * 4 + (size_t)K2_MAX_HALOS * sizeof(float4) + sizeof(K2CStacks) * (size_t)warps;
}

struct HaloBinArgs {
    HaloArgs h;                       /* h.pre holds {chi, clo, slo, shi} here (prefilter.h uc2_bounds_cells) */
    const unsigned int* cell_bits;    /* [K2_BITWORDS]: bit set when any of the 2 x 2 cells lists a halo */
    const uint2* cell_range;          /* [K2_GRID*K2_GRID] {first, end} into cell_items */
    const unsigned short* cell_items; /* halo index inside the launch */
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/* one out-of-range position against every halo of the launch (rare) */
__device__ __noinline__ void k2c_exact_all(const HaloArgs& a, float x, float y, float z, unsigned int local_n, int lane) {
    for (int h0 = 0; h0 < a.nhalos; h0 += 32) {
        const int h = h0 + lane;
        k2_exact_warp(a, h < a.nhalos, h < a.nhalos ? h : 0, x, y, z, local_n, lane);
    }
}

template <bool VEC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k2_cutout_halo_cells(const __grid_constant__ HaloBinArgs b) {
    constexpr int K2C_THREADS = WARPS * 32, K2C_WARPS = WARPS;
    extern __shared__ __align__(16) unsigned char s_cells[];
    unsigned int* s_bits = reinterpret_cast<unsigned int*>(s_cells);
    float4* s_pre = reinterpret_cast<float4*>(s_cells + (size_t)K2_BITWORDS * 4);
    K2CStacks* s_st = reinterpret_cast<K2CStacks*>(s_cells + (size_t)K2_BITWORDS * 4 + (size_t)K2_MAX_HALOS * sizeof(float4));
    /* survivors and rough survivors per halo, counted per CTA (a global atomic per pair on 250 counters serialises in
     * the L2: profiles/r02) */
    unsigned int* s_rough = reinterpret_cast<unsigned int*>(s_cells + (size_t)K2_BITWORDS * 4 + (size_t)K2_MAX_HALOS * sizeof(float4) +
                                                            sizeof(K2CStacks) * (size_t)WARPS);
    unsigned int* s_fine = s_rough + K2_MAX_HALOS;
    for (int i = threadIdx.x; i < K2_MAX_HALOS; i += K2C_THREADS) { s_rough[i] = 0u; s_fine[i] = 0u; }
    const HaloArgs& a = b.h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    K2CStacks& S = s_st[warp];
    for (int i = threadIdx.x; i < K2_BITWORDS / 4; i += K2C_THREADS)
        reinterpret_cast<uint4*>(s_bits)[i] = __ldg(reinterpret_cast<const uint4*>(b.cell_bits) + i);
    for (int h = threadIdx.x; h < a.nhalos; h += K2C_THREADS) s_pre[h] = a.pre[h];
    __syncthreads();
    const unsigned int total_warps = gridDim.x * K2C_WARPS;
    const unsigned int ntiles = (unsigned int)((a.n + K2C_TILE - 1) / K2C_TILE);
    unsigned long long n_cand = 0;
    unsigned int nh = 0, nc = 0, nx = 0; /* warp-uniform stack heights: hits, tested pairs, pairs for the exact path */
    unsigned int w0 = 0;         /* hits of the current tile already pushed (non-zero: the tile is being re-read) */
    bool flush = false;

    /* a whole, 16-byte aligned tile is fetched one trip ahead into shared memory (each lane copies, and later reads
     * back, its own 3 x 16 bytes): the DRAM latency of tile t+1 hides behind the work on tile t without holding
     * registers across that work */
    auto whole = [&](unsigned int t) { return VEC && t < ntiles && (long long)(t + 1) * K2C_TILE <= a.n; };
    auto prefetch = [&](unsigned int t) {
        const long long o = (long long)t * K2C_TILE + 4 * lane;
        cp_async16(&S.next[0][lane], a.x + o);
        cp_async16(&S.next[1][lane], a.y + o);
        cp_async16(&S.next[2][lane], a.z + o);
        cp_async_commit();
    };
    {
        const unsigned int first = blockIdx.x * K2C_WARPS + warp;
        if (whole(first)) prefetch(first);
    }

    for (unsigned int tile = blockIdx.x * K2C_WARPS + warp; !flush;) {
        if (tile < ntiles) {
            /* nothing of the tile stays in registers across the rounds below: a tile with more hits than the
             * stack has room for is simply read again for its next pass (rare) */
            const long long base = (long long)tile * K2C_TILE;
            float px[4], py[4], pz[4];
            unsigned int valid = 0;
            if (w0 == 0 && whole(tile)) {
                cp_async_wait_all();
                const float4 vx = S.next[0][lane], vy = S.next[1][lane], vz = S.next[2][lane];
                if (whole(tile + total_warps)) prefetch(tile + total_warps);
                px[0] = vx.x; px[1] = vx.y; px[2] = vx.z; px[3] = vx.w;
                py[0] = vy.x; py[1] = vy.y; py[2] = vy.z; py[3] = vy.w;
                pz[0] = vz.x; pz[1] = vz.y; pz[2] = vz.z; pz[3] = vz.w;
                valid = 0xfu;
            } else if (VEC && base + K2C_TILE <= a.n) { /* another pass over the same tile */
                const float4 vx = ldg_stream4(a.x + base + 4 * lane), vy = ldg_stream4(a.y + base + 4 * lane),
                             vz = ldg_stream4(a.z + base + 4 * lane);
                px[0] = vx.x; px[1] = vx.y; px[2] = vx.z; px[3] = vx.w;
                py[0] = vy.x; py[1] = vy.y; py[2] = vy.z; py[3] = vy.w;
                pz[0] = vz.x; pz[1] = vz.y; pz[2] = vz.z; pz[3] = vz.w;
                valid = 0xfu;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long g = base + 4 * lane + j;
                    px[j] = py[j] = pz[j] = 1.0f;
                    if (g < a.n) {
                        px[j] = __ldcs(a.x + g); py[j] = __ldcs(a.y + g); pz[j] = __ldcs(a.z + g);
                        valid |= 1u << j;
                    }
                }
            }
            unsigned int cell[4], insane = 0, hit = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { /* cell of every particle; most are ruled out by the shared-memory bitmap */
                float pc, ps;
                bool sane;
                k2c_coords(px[j], py[j], pz[j], pc, ps, sane);
                if (!sane) insane |= 1u << j;
                /* NaN -> 0, anything large -> 2^32-1 -> G-1: always a valid cell */
                const unsigned int ic = min(__float2uint_rz(fmaf(pc, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
                const unsigned int is = min(__float2uint_rz(fmaf(ps, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
                cell[j] = ic * K2_GRID + is;
                const unsigned int bit = (ic >> K2_CSHIFT) * K2_COARSE + (is >> K2_CSHIFT);
                hit |= ((s_bits[bit >> 5] >> (bit & 31u)) & 1u) << j;
            }
            insane &= valid;
            hit &= valid & ~insane;
            uint2 rg[4];
            unsigned int mine = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { /* the look-ups that remain are issued together */
                rg[j] = make_uint2(0u, 0u);
                if ((hit >> j) & 1u) rg[j] = __ldg(b.cell_range + cell[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) mine += rg[j].y != rg[j].x ? 1u : 0u;
            /* out-of-range positions: the exact path decides, against every halo of the launch, warp-cooperatively */
            if (w0 == 0 && __any_sync(0xffffffffu, insane != 0)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned int m = __ballot_sync(0xffffffffu, (insane >> j) & 1u);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const float x = __shfl_sync(0xffffffffu, px[j], src), y = __shfl_sync(0xffffffffu, py[j], src),
                                    z = __shfl_sync(0xffffffffu, pz[j], src);
                        k2c_exact_all(a, x, y, z, (unsigned int)(a.local_base + base + 4 * src + j), lane);
                        n_cand += (unsigned int)a.nhalos;
                    }
                }
            }
            /* position of this lane's hits among the tile's hits */
            unsigned int inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const unsigned int total = __shfl_sync(0xffffffffu, inc, 31);
            const unsigned int w1 = min(total, w0 + (K2C_HCAP - nh));
            unsigned int pos = inc - mine;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (rg[j].y != rg[j].x) {
                    if (pos >= w0 && pos < w1) {
                        const unsigned int e = nh + (pos - w0);
                        S.hp[e] = make_float4(px[j], py[j], pz[j], __uint_as_float((unsigned int)(base + 4 * lane + j)));
                        S.hl[e] = make_uint2(rg[j].x, rg[j].y - rg[j].x);
                    }
                    ++pos;
                }
            nh += w1 - w0;
            if (w1 < total) {
                w0 = w1; /* same tile again after the rounds */
            } else {
                w0 = 0;
                tile += total_warps;
            }
            __syncwarp();
        } else {
            flush = true; /* past this warp's last tile: empty the stacks */
        }
        /* pop 32 hits (whatever is left when the stacks are being emptied), test ONE list item of each, and push
         * the hits with a longer list back: every round tests with all lanes busy */
        while (nh >= 32u || (flush && (nh | nc | nx) != 0u)) {
            const unsigned int m = min(nh, 32u);
            nh -= m;
            float4 q = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            uint2 l = make_uint2(0u, 0u);
            unsigned int h = 0;
            bool pass = false;
            if ((unsigned int)lane < m) {
                q = S.hp[nh + lane];
                l = S.hl[nh + lane];
                h = __ldg(b.cell_items + l.x);
                float pc, ps;
                bool sane;
                k2c_coords(q.x, q.y, q.z, pc, ps, sane);
                const float4 W = s_pre[h];
                pass = (pc <= W.x) & (pc >= W.y) & (ps >= W.z) & (ps <= W.w);
            }
            const unsigned int mm = __ballot_sync(0xffffffffu, l.y > 1u);
            const unsigned int pm = __ballot_sync(0xffffffffu, pass);
            __syncwarp(); /* every lane holds its hit in registers: the pushes below may overwrite it */
            if (l.y > 1u) {
                const unsigned int e = nh + __popc(mm & lt_mask);
                S.hp[e] = q;
                S.hl[e] = make_uint2(l.x + 1u, l.y - 1u);
            }
            nh += __popc(mm);
            if (pass) {
                const unsigned int e = nc + __popc(pm & lt_mask);
                S.cp[e] = q;
                S.ch[e] = h;
            }
            nc += __popc(pm);
            __syncwarp();
            /* tested pairs, 32 at a time (at the very end, whatever is left): the second filter, then the exact
             * path for what it cannot decide */
            const bool last = flush && nh == 0u;
            for (;;) {
                if (nx >= 32u || (last && nc == 0u && nx != 0u)) { /* exact arithmetic, newest 32 */
                    const unsigned int mx = min(nx, 32u);
                    nx -= mx;
                    const bool lv = (unsigned int)lane < mx;
                    const unsigned int e = lv ? nx + lane : 0u;
                    const float4 c4 = S.xp[e];
                    k2_exact_warp(a, lv, (int)S.xh[e], c4.x, c4.y, c4.z, (unsigned int)a.local_base + __float_as_uint(c4.w), lane, s_rough, s_fine);
                    n_cand += mx;
                    __syncwarp();
                    continue;
                }
                if (nc >= 32u || (last && nc != 0u)) {
                    /* A pair that is certainly inside the rough window (inner (c, s) window) AND certainly outside
                     * the fine window (rotated position against the widened fine window) only counts as a rough
                     * survivor: 5 of 6 pairs of a cluster-sized cutout.  The rest takes the exact path. */
                    const unsigned int mc = min(nc, 32u);
                    nc -= mc;
                    const bool lv = (unsigned int)lane < mc;
                    const unsigned int e = lv ? nc + lane : 0u;
                    const float4 c4 = S.cp[e];
                    const unsigned int h = lv ? S.ch[e] : 0u;
                    const float4* __restrict__ Hq = reinterpret_cast<const float4*>(a.halos + h);
                    const float4 q0 = __ldg(Hq), q1 = __ldg(Hq + 1), q2 = __ldg(Hq + 2), q5 = __ldg(Hq + 5), q6 = __ldg(Hq + 6);
                    float pc, ps;
                    bool sane;
                    k2c_coords(c4.x, c4.y, c4.z, pc, ps, sane);
                    const bool sure_rough = (pc <= q5.x) & (pc >= q5.y) & (ps >= q5.z) & (ps <= q5.w);
                    const float R[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
                    float rx, ry, rz;
                    ccref::mat_vec(R, c4.x, c4.y, c4.z, rx, ry, rz); /* the exact path's own rotated position */
                    const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                    const bool sane_r = (fabsf(rx) >= CC_SANE_LO) & (r2 <= 1.2089258e24f);
                    const float cf = rz * rsqrt_fast(r2), tf = ry * rcp_fast(rx);
                    const bool maybe_fine = !sane_r | ((cf <= q6.x) & (cf >= q6.y) & (tf >= q6.z) & (tf <= q6.w));
                    const bool counted = lv && sure_rough && !maybe_fine;
                    if (counted) atomicAdd(&s_rough[h], 1u); /* rough_cutout_size, :1436 */
                    const bool need = lv && !counted;
                    const unsigned int xm = __ballot_sync(0xffffffffu, need);
                    if (need) {
                        const unsigned int x = nx + __popc(xm & lt_mask);
                        S.xp[x] = c4;
                        S.xh[x] = h;
                    }
                    nx += __popc(xm);
                    __syncwarp();
                    continue;
                }
                break;
            }
        }
    }
    if (lane == 0 && n_cand) atomicAdd(a.n_cand, n_cand);
    __syncthreads();
    for (int i = threadIdx.x; i < a.nhalos; i += K2C_THREADS) {
        if (s_rough[i]) atomicAdd(a.counts + a.total_halos + (unsigned int)a.halo0 + i, (unsigned long long)s_rough[i]);
        if (s_fine[i]) atomicAdd(a.counts + (unsigned int)a.halo0 + i, (unsigned long long)s_fine[i]);
    }
}

/* ---- many halos on an INDEXED resident step -----------------------------------------------------------------------
 * The reference sorts every step by theta once and then finds each halo's candidates with two binary searches
 * (processLC.cpp:1214-1221, :1334-1340): per halo it touches the annulus, not the step.  The walkers above and below
 * do the opposite -- they stream all 12 B/particle for every pass and look the HALOS up -- which is the right trade
 * for a step that is cut once.  For a step that stays resident and is cut again and again (BASELINE config 5: 1000
 * halos, several box sizes) the particles themselves are indexed once: a counting sort by the particle's cell of the
 * same 2048 x 2048 (cos theta_u, sin phi_u) grid writes {x, y, z, index} (16 B) in cell order, with the first row of
 * every cell in `start`; out-of-range positions (k2c_coords !sane: they must meet every halo) sit in one extra bin
 * behind the last cell.  A halo's window is a rectangle of cells, i.e. one CONTIGUOUS run of rows per grid line:
 * a pass reads 16 B for the particles of those runs and nothing else -- 15-20 M of 125 M for the 1000 cutouts of
 * config 5 -- and hands (particle, halo) pairs to the same second filter and exact path as the cell walkers, so the
 * records, counts and therefore the cutouts are the same, bit for bit.
 *
 *   k2p_hist / k2p_scan_{a,b,c} / k2p_scatter   build (once per resident step, capi.cu maybe_build_index)
 *   k2p_query   per pass: (halo, grid line) items handed out to the warps by an atomic counter, largest first */
constexpr unsigned int K2P_NCELL = (unsigned int)K2_GRID * (unsigned int)K2_GRID; /* + 1 bin: out-of-range positions */
constexpr unsigned int K2P_NBIN = K2P_NCELL + 1u;
constexpr unsigned int K2P_ALL = 0xffffu; /* item halo: every halo of the launch (the out-of-range bin) */
constexpr int K2P_SCAN_BLOCK = 4096;

__device__ __forceinline__ unsigned int k2p_cell(float x, float y, float z) {
    float pc, ps;
    bool sane;
    k2c_coords(x, y, z, pc, ps, sane);
    /* the SAME expressions as the cell walkers (k2_cutout_halo_cells, k2c_sift): the halo lists were rasterised for them */
    const unsigned int ic = min(__float2uint_rz(fmaf(pc, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
    const unsigned int is = min(__float2uint_rz(fmaf(ps, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
    return sane ? ic * (unsigned int)K2_GRID + is : K2P_NCELL;
}
__global__ void k2p_hist(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, long long n,
                         unsigned int* __restrict__ hist) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(hist + k2p_cell(__ldcs(x + i), __ldcs(y + i), __ldcs(z + i)), 1u);
}
/* exclusive scan of the K2P_NBIN bin counts in three small kernels: block sums, scan of the sums, scan + offset */
__global__ void __launch_bounds__(256) k2p_scan_a(const unsigned int* __restrict__ hist, unsigned int* __restrict__ partial) {
    __shared__ unsigned int s_w[8];
    const unsigned int b0 = blockIdx.x * K2P_SCAN_BLOCK;
    unsigned int sum = 0;
    for (unsigned int i = threadIdx.x; i < (unsigned int)K2P_SCAN_BLOCK; i += 256)
        if (b0 + i < K2P_NBIN) sum += hist[b0 + i];
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(32) k2p_scan_b(unsigned int* partial, unsigned int nblocks) {
    const unsigned int lane = threadIdx.x, per = (nblocks + 31u) / 32u;
    unsigned int sum = 0;
    for (unsigned int i = lane * per; i < min(nblocks, (lane + 1u) * per); ++i) sum += partial[i];
    unsigned int inc = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned int)d) inc += t;
    }
    unsigned int run = inc - sum;
    for (unsigned int i = lane * per; i < min(nblocks, (lane + 1u) * per); ++i) { const unsigned int v = partial[i]; partial[i] = run; run += v; }
}
__global__ void __launch_bounds__(256) k2p_scan_c(const unsigned int* __restrict__ hist, const unsigned int* __restrict__ partial,
                                                  unsigned int* __restrict__ start) {
    __shared__ unsigned int s_w[8];
    constexpr int PER = K2P_SCAN_BLOCK / 256;
    const unsigned int b0 = blockIdx.x * K2P_SCAN_BLOCK + threadIdx.x * PER;
    unsigned int v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = (b0 + k < K2P_NBIN) ? hist[b0 + k] : 0u; sum += v[k]; }
    const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned int inc = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned int)d) inc += t;
    }
    if (lane == 31u) s_w[warp] = inc;
    __syncthreads();
    unsigned int run = partial[blockIdx.x] + inc - sum;
    for (unsigned int w = 0; w < warp; ++w) run += s_w[w];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if (b0 + k < K2P_NBIN) start[b0 + k] = run;
        run += v[k];
        if (b0 + k == K2P_NBIN - 1u) start[K2P_NBIN] = run; /* = n */
    }
}
/* fill: a copy of start; every particle takes the next row of its bin (the order inside a bin is whatever the atomics
 * make it: survivors are ordered by (theta_u, index) afterwards, k3_*) */
__global__ void k2p_scatter(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, long long n,
                            unsigned int* __restrict__ fill, float4* __restrict__ sorted) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float px = __ldcs(x + i), py = __ldcs(y + i), pz = __ldcs(z + i);
        const unsigned int row = atomicAdd(fill + k2p_cell(px, py, pz), 1u);
        sorted[row] = make_float4(px, py, pz, __uint_as_float((unsigned int)i));
    }
}

struct PidxArgs {
    HaloArgs h;                      /* x, y, z unused; n = particles in range of the n_limit clamp (index < n) */
    const float4* sorted;            /* {x, y, z, index bits} in bin order */
    const unsigned int* start;       /* [K2P_NBIN + 1] first row of every bin */
    const uint2* items;              /* {halo | grid line << 16, first column | last column << 16}; halo K2P_ALL: the
                                        out-of-range bin against every halo of the launch */
    unsigned int nitems;
    unsigned int* work;              /* zeroed: next item to hand out */
    unsigned long long* tested;      /* stat: particles read */
};
__device__ __forceinline__ void k2p_rows(const PidxArgs& p, uint2 it, unsigned int& begin, unsigned int& end) {
    if ((it.x & 0xffffu) == K2P_ALL) { begin = __ldg(p.start + K2P_NCELL); end = __ldg(p.start + K2P_NBIN); return; }
    const unsigned int line = (it.x >> 16) * (unsigned int)K2_GRID;
    begin = __ldg(p.start + line + (it.y & 0xffffu));
    end = __ldg(p.start + line + (it.y >> 16) + 1u);
}
constexpr int K2P_THREADS = 128; /* five CTAs per SM at 96 registers: 20 warps, each with its own slice of the steps */
struct __align__(16) K2PStacks {
    float4 cp[K2C_CCAP];         /* tested pair: x, y, z, particle index (bits) */
    unsigned int ch[K2C_CCAP];   /*              halo inside the launch */
    float4 xp[K2C_CCAP];         /* pair that needs the exact arithmetic */
    unsigned int xh[K2C_CCAP];
};
__global__ void __launch_bounds__(K2P_THREADS, 5) k2p_query(const __grid_constant__ PidxArgs p) {
    __shared__ K2PStacks s_st[K2P_THREADS / 32];
    __shared__ unsigned int s_rough[K2_MAX_HALOS], s_fine[K2_MAX_HALOS];
    for (int i = threadIdx.x; i < K2_MAX_HALOS; i += K2P_THREADS) { s_rough[i] = 0u; s_fine[i] = 0u; }
    __syncthreads();
    const HaloArgs& a = p.h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    K2PStacks& S = s_st[warp];
    unsigned int nc = 0, nx = 0; /* warp-uniform stack heights: tested pairs, pairs for the exact path */
    unsigned long long n_cand = 0, n_tested = 0;

    /* tested pairs, 32 at a time (when `last`, whatever is left): the second filter, then the exact path for what it
     * cannot decide -- the tail of k2_cutout_halo_cells */
    auto drain = [&](bool last) {
        for (;;) {
            if (nx >= 32u || (last && nc == 0u && nx != 0u)) {
                const unsigned int mx = min(nx, 32u);
                nx -= mx;
                const bool lv = (unsigned int)lane < mx;
                const unsigned int e = lv ? nx + lane : 0u;
                const float4 c4 = S.xp[e];
                k2_exact_warp(a, lv, (int)S.xh[e], c4.x, c4.y, c4.z, (unsigned int)a.local_base + __float_as_uint(c4.w), lane, s_rough, s_fine);
                n_cand += mx;
                __syncwarp();
                continue;
            }
            if (nc >= 32u || (last && nc != 0u)) {
                const unsigned int mc = min(nc, 32u);
                nc -= mc;
                const bool lv = (unsigned int)lane < mc;
                const unsigned int e = lv ? nc + lane : 0u;
                const float4 c4 = S.cp[e];
                const unsigned int h = lv ? S.ch[e] : 0u;
                const float4* __restrict__ Hq = reinterpret_cast<const float4*>(a.halos + h);
                const float4 q0 = __ldg(Hq), q1 = __ldg(Hq + 1), q2 = __ldg(Hq + 2), q5 = __ldg(Hq + 5), q6 = __ldg(Hq + 6);
                float pc, ps;
                bool sane;
                k2c_coords(c4.x, c4.y, c4.z, pc, ps, sane);
                const bool sure_rough = (pc <= q5.x) & (pc >= q5.y) & (ps >= q5.z) & (ps <= q5.w);
                const float R[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
                float rx, ry, rz;
                ccref::mat_vec(R, c4.x, c4.y, c4.z, rx, ry, rz); /* the exact path's own rotated position */
                const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                const bool sane_r = (fabsf(rx) >= CC_SANE_LO) & (r2 <= 1.2089258e24f);
                const float cf = rz * rsqrt_fast(r2), tf = ry * rcp_fast(rx);
                const bool maybe_fine = !sane_r | ((cf <= q6.x) & (cf >= q6.y) & (tf >= q6.z) & (tf <= q6.w));
                const bool counted = lv && sure_rough && !maybe_fine;
                if (counted) atomicAdd(&s_rough[h], 1u); /* rough_cutout_size, :1436 */
                const bool need = lv && !counted;
                const unsigned int xm = __ballot_sync(0xffffffffu, need);
                if (need) {
                    const unsigned int x = nx + __popc(xm & lt_mask);
                    S.xp[x] = c4;
                    S.xh[x] = h;
                }
                nx += __popc(xm);
                __syncwarp();
                continue;
            }
            break;
        }
    };

    /* Items (one halo x one grid line: a contiguous run of rows) are handed out by an atomic counter, largest first
     * (the host sorts them), one per warp at a time: how much exact arithmetic a row costs depends on where it lies in
     * its halo's window, so equal static slices left most warps waiting at the final barrier for a few (ncu: 31 % of
     * the samples).  The hand-out is pipelined two deep -- while item k is worked on, the rows and window of item k+1
     * are on their way and the ticket of item k+2 has been drawn -- and inside an item the 512 bytes of the next step
     * are requested before the current step is worked on (+ an L2 prefetch four steps ahead). */
    auto ticket = [&]() {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(p.work, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    struct Meta { unsigned int h, row_begin, row_end; float4 W; };
    auto meta_of = [&](unsigned int it) {
        Meta m;
        m.h = 0u; m.row_begin = 0u; m.row_end = 0u; m.W = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < p.nitems) {
            const uint2 v = __ldg(p.items + it);
            m.h = v.x & 0xffffu;
            k2p_rows(p, v, m.row_begin, m.row_end);
            if (m.h != K2P_ALL) m.W = __ldg(a.pre + m.h);
        }
        return m;
    };
    unsigned int it_cur = ticket(), it_next = ticket();
    Meta cur = meta_of(it_cur);
    float4 q_next = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    bool have_first = false; /* q_next holds the first 32 rows of `cur` */
    while (it_cur < p.nitems) {
        const unsigned int it_after = ticket();
        const Meta nxt = meta_of(it_next);
        const unsigned int h = cur.h;
        const float4 W = cur.W;
        if (!have_first) {
            const unsigned int row = cur.row_begin + (unsigned int)lane;
            q_next = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            if (row < cur.row_end) q_next = __ldcs(p.sorted + row);
        }
        have_first = false;
        for (unsigned int r0 = cur.row_begin; r0 < cur.row_end; r0 += 32u) {
            const float4 q = q_next;
            const unsigned int row = r0 + (unsigned int)lane;
            bool live = row < cur.row_end;
            q_next = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            if (r0 + 32u < cur.row_end) {
                if (row + 32u < cur.row_end) q_next = __ldcs(p.sorted + row + 32u);
            } else { /* last step of the item: the first rows of the next one */
                const unsigned int rn = nxt.row_begin + (unsigned int)lane;
                if (rn < nxt.row_end) q_next = __ldcs(p.sorted + rn);
                have_first = true;
            }
            if ((lane & 7) == 0 && row + 160u < cur.row_end) prefetch_l2(p.sorted + row + 160u);
            live = live && (long long)__float_as_uint(q.w) < a.n; /* the n_limit clamp (reference tail drop) */
            n_tested += __popc(__ballot_sync(0xffffffffu, live));
            if (h == K2P_ALL) { /* out-of-range positions: the exact path decides, against every halo of the launch */
                unsigned int m = __ballot_sync(0xffffffffu, live);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float x = __shfl_sync(0xffffffffu, q.x, src), y = __shfl_sync(0xffffffffu, q.y, src),
                                z = __shfl_sync(0xffffffffu, q.z, src);
                    const unsigned int idx = __shfl_sync(0xffffffffu, __float_as_uint(q.w), src);
                    k2c_exact_all(a, x, y, z, (unsigned int)a.local_base + idx, lane);
                    n_cand += (unsigned int)a.nhalos;
                }
                continue;
            }
            float pc, ps;
            bool sane;
            k2c_coords(q.x, q.y, q.z, pc, ps, sane);
            const bool pass = live & (pc <= W.x) & (pc >= W.y) & (ps >= W.z) & (ps <= W.w);
            const unsigned int pm = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const unsigned int e = nc + __popc(pm & lt_mask);
                S.cp[e] = q;
                S.ch[e] = h;
            }
            nc += __popc(pm);
            __syncwarp();
            drain(false);
        }
        it_cur = it_next;
        it_next = it_after;
        cur = nxt;
    }
    drain(true);
    if (lane == 0 && n_cand) atomicAdd(a.n_cand, n_cand);
    if (lane == 0 && n_tested) atomicAdd(p.tested, n_tested);
    __syncthreads();
    for (int i = threadIdx.x; i < a.nhalos; i += K2P_THREADS) {
        if (s_rough[i]) atomicAdd(a.counts + a.total_halos + (unsigned int)a.halo0 + i, (unsigned long long)s_rough[i]);
        if (s_fine[i]) atomicAdd(a.counts + (unsigned int)a.halo0 + i, (unsigned long long)s_fine[i]);
    }
}

/* ---- many halos, staged: sift -> match -> exact ----------------------------------------------------------------
 * The same cell index, but every stage is its own full-occupancy kernel and hands the next one a COMPACT work list, so
 * that each kernel runs with all lanes busy and 32 warps per SM hiding its latencies (round 1's single kernel kept the
 * lanes busy with per-warp stacks instead, at 20 warps per SM because of its 161 KB of shared memory):
 *
 *   k2c_sift    streams x,y,z (12 B/particle, the only pass over the step), computes (c, s), looks the particle's
 *               coarse cell up in the 32 KB occupancy bitmap (L1-resident) and appends the 10-25 % that hit -- and the
 *               out-of-range positions -- as (x, y, z, index) to the warp's private region of a candidate buffer
 *               (register cursor, ballot positions, no atomics).
 *   k2c_match   one thread per candidate: cell -> {first, end} of the halo list -> per list item the (c, s) window of
 *               that halo, then the second filter (certainly inside the rough window AND certainly outside the fine
 *               one: only the rough counter, kept per CTA in shared memory); what it cannot decide is a
 *               (particle, halo) PAIR, staged in shared memory and appended to the pair buffer with one atomic per
 *               1024 pairs.  The CTA walks the lists in rounds (item k of every candidate's list), so a round emits at
 *               most one pair per thread.  An out-of-range position is paired with every halo.
 *   k2c_exact   one thread per pair: the exact path (k2_exact_warp).
 * ---------------------------------------------------------------------------------------------------------------- */
struct SiftArgs {
    const float *x, *y, *z;
    long long n;                  /* particles to scan (after the n_limit clamp) */
    long long local_base;         /* index of x[0] inside the step */
    const unsigned int* cell_bits;
    float4* cand;                 /* [warps][region_cap] x, y, z, index inside the step (bits) */
    unsigned int region_cap;
    unsigned int* cand_count;     /* [warps] */
    unsigned long long* misc;     /* [2] = max candidates any warp wanted (region overflow check) */
    unsigned int ntiles;
    int prefetch;                 /* L2-prefetch the warp's next tile while this one is worked on */
};
template <bool VEC>
__global__ void __launch_bounds__(SCAN_THREADS, 4) k2c_sift(const __grid_constant__ SiftArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int total_warps = gridDim.x * SCAN_WARPS, gwarp = blockIdx.x * SCAN_WARPS + warp;
    float4* __restrict__ region = a.cand + (size_t)gwarp * a.region_cap;
    unsigned int cursor = 0; /* warp-uniform */
    for (unsigned int tile = gwarp; tile < a.ntiles; tile += total_warps) {
        const long long base = (long long)tile * SCAN_TILE;
        float px[8], py[8], pz[8];
        const unsigned int valid = load_tile<VEC>(a.x, a.y, a.z, base, a.n, lane, px, py, pz);
        if (a.prefetch) prefetch_tile(a.x, a.y, a.z, base + (long long)total_warps * SCAN_TILE, a.n, lane);
        unsigned int m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float pc, ps;
            bool sane;
            k2c_coords(px[j], py[j], pz[j], pc, ps, sane);
            /* NaN -> 0, anything large -> 2^32-1 -> G-1: always a valid cell */
            const unsigned int ic = min(__float2uint_rz(fmaf(pc, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
            const unsigned int is = min(__float2uint_rz(fmaf(ps, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
            const unsigned int bit = (ic >> K2_CSHIFT) * K2_COARSE + (is >> K2_CSHIFT);
            const unsigned int hit = (__ldg(a.cell_bits + (bit >> 5)) >> (bit & 31u)) & 1u;
            m |= (hit | (sane ? 0u : 1u)) << j;
        }
        m &= valid;
        /* every lane appends its hits at consecutive slots behind those of the lower lanes (the order inside the region
         * does not matter): one warp scan + up to 8 predicated 16-byte stores (a ballot per element was measured at
         * 160 warp instructions per tile, profiles/r02) */
        const unsigned int cnt = __popc(m);
        unsigned int inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        unsigned int pos = cursor + inc - cnt;
        cursor += __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if ((m >> j) & 1u) {
                if (pos < a.region_cap)
                    region[pos] = make_float4(px[j], py[j], pz[j], __uint_as_float((unsigned int)(a.local_base + base) + tile_local(lane, j)));
                ++pos;
            }
    }
    if (lane == 0) {
        a.cand_count[gwarp] = cursor;
        atomicAdd(a.misc + 6, (unsigned long long)cursor); /* candidates of the pass: the hit rate capi.cu decides on */
        atomicMax(a.misc + 2, (unsigned long long)cursor);
    }
}

constexpr int K2M_THREADS = 256;
constexpr int K2M_WARPS = K2M_THREADS / 32;
constexpr int K2M_STAGE = 128; /* pairs staged per WARP before they go to the pair buffer with one atomic */
struct MatchArgs {
    HaloArgs h;                       /* h.pre holds {chi, clo, slo, shi} (prefilter.h uc2_bounds_cells) */
    const uint2* cell_range;
    const unsigned short* cell_items;
    const float4* cand;
    const unsigned int* cand_count;
    unsigned int region_cap, nregions;
    float4* pair_p;                   /* x, y, z, index inside the step (bits) */
    unsigned int* pair_h;             /* halo inside the launch */
    unsigned long long pair_cap;
    unsigned long long* pair_count;   /* appended pairs (may exceed pair_cap: overflow is detected on the host) */
};
__global__ void __launch_bounds__(K2M_THREADS) k2c_match(const __grid_constant__ MatchArgs a) {
    __shared__ unsigned int s_rough[K2_MAX_HALOS];
    __shared__ float4 s_pp[K2M_WARPS][K2M_STAGE];
    __shared__ unsigned int s_ph[K2M_WARPS][K2M_STAGE];
    const HaloArgs& H = a.h;
    const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    for (unsigned int i = tid; i < (unsigned int)H.nhalos; i += K2M_THREADS) s_rough[i] = 0u;
    __syncthreads();
    unsigned int wn = 0; /* warp-uniform: pairs staged by this warp */
    auto flush = [&]() { /* all lanes of the warp */
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.pair_count, (unsigned long long)wn);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (unsigned int i = lane; i < wn; i += 32u)
            if (base + i < a.pair_cap) { a.pair_p[base + i] = s_pp[warp][i]; a.pair_h[base + i] = s_ph[warp][i]; }
        __syncwarp();
        wn = 0;
    };
    /* warps take (region, 32-candidate slice) items independently: no block barrier inside the loop */
    for (unsigned int r = blockIdx.x; r < a.nregions; r += gridDim.x) {
        const unsigned int cnt = min(a.cand_count[r], a.region_cap);
        const float4* __restrict__ region = a.cand + (size_t)r * a.region_cap;
        for (unsigned int e0 = warp * 32u; e0 < cnt; e0 += K2M_THREADS) {
            const unsigned int e = e0 + lane;
            float4 c4 = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            unsigned int it = 0, left = 0;
            bool all = false; /* out-of-range position: paired with every halo of the launch */
            float pc = 0.0f, ps = 0.0f;
            if (e < cnt) {
                c4 = region[e];
                bool sane;
                k2c_coords(c4.x, c4.y, c4.z, pc, ps, sane);
                if (!sane) {
                    all = true;
                    left = (unsigned int)H.nhalos;
                } else {
                    const unsigned int ic = min(__float2uint_rz(fmaf(pc, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
                    const unsigned int is = min(__float2uint_rz(fmaf(ps, 0.5f * K2_GRID, 0.5f * K2_GRID)), (unsigned int)(K2_GRID - 1));
                    const uint2 rg = __ldg(a.cell_range + ic * K2_GRID + is);
                    it = rg.x;
                    left = rg.y - rg.x;
                }
            }
            while (__any_sync(0xffffffffu, left > 0u)) { /* one list item per lane and round */
                bool emit = false;
                unsigned int h = 0;
                if (left > 0u) {
                    h = all ? it : (unsigned int)__ldg(a.cell_items + it);
                    ++it;
                    --left;
                    if (all) {
                        emit = true;
                    } else {
                        const float4 W = __ldg(H.pre + h);
                        if ((pc <= W.x) & (pc >= W.y) & (ps >= W.z) & (ps <= W.w)) {
                            /* certainly inside the rough window AND certainly outside the fine one: only the rough
                             * counter (processLC.cpp:1436); everything else takes the exact path */
                            const float4* __restrict__ Hq = reinterpret_cast<const float4*>(H.halos + h);
                            const float4 q0 = __ldg(Hq), q1 = __ldg(Hq + 1), q2 = __ldg(Hq + 2), q5 = __ldg(Hq + 5), q6 = __ldg(Hq + 6);
                            const bool sure_rough = (pc <= q5.x) & (pc >= q5.y) & (ps >= q5.z) & (ps <= q5.w);
                            const float R[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
                            float rx, ry, rz;
                            ccref::mat_vec(R, c4.x, c4.y, c4.z, rx, ry, rz); /* the exact path's own rotated position */
                            const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                            const bool sane_r = (fabsf(rx) >= CC_SANE_LO) & (r2 <= 1.2089258e24f);
                            const float cf = rz * rsqrt_fast(r2), tf = ry * rcp_fast(rx);
                            const bool maybe_fine = !sane_r | ((cf <= q6.x) & (cf >= q6.y) & (tf >= q6.z) & (tf <= q6.w));
                            if (sure_rough && !maybe_fine) atomicAdd(&s_rough[h], 1u);
                            else emit = true;
                        }
                    }
                }
                const unsigned int em = __ballot_sync(0xffffffffu, emit);
                if (em) { /* warp-uniform */
                    if (emit) {
                        const unsigned int slot = wn + __popc(em & lt_mask);
                        s_pp[warp][slot] = c4;
                        s_ph[warp][slot] = h;
                    }
                    wn += __popc(em);
                    if (wn > (unsigned int)(K2M_STAGE - 32)) flush();
                }
            }
        }
    }
    if (wn) flush();
    __syncthreads();
    for (unsigned int i = tid; i < (unsigned int)H.nhalos; i += K2M_THREADS)
        if (s_rough[i]) atomicAdd(H.counts + H.total_halos + (unsigned int)H.halo0 + i, (unsigned long long)s_rough[i]);
}

struct ExactArgs {
    HaloArgs h;
    const float4* pair_p;
    const unsigned int* pair_h;
    unsigned long long pair_cap;
    const unsigned long long* pair_count;
    unsigned long long* misc; /* [3] = max pairs any round wanted (pair buffer overflow check) */
};
constexpr int K2X_THREADS = 512;
__global__ void __launch_bounds__(K2X_THREADS) k2c_exact(const __grid_constant__ ExactArgs a) {
    /* survivors and rough survivors per halo: counted per CTA in shared memory, added to the global counters once
     * (a global atomic per pair was measured: 4 M atomics on the 32 cache lines of 250 counters = 0.3 ms) */
    __shared__ unsigned int s_rough[K2_MAX_HALOS], s_fine[K2_MAX_HALOS];
    const unsigned long long want = *a.pair_count, npairs = min(want, a.pair_cap);
    const int lane = threadIdx.x & 31;
    for (unsigned int i = threadIdx.x; i < (unsigned int)a.h.nhalos; i += K2X_THREADS) { s_rough[i] = 0u; s_fine[i] = 0u; }
    __syncthreads();
    for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; i0 < npairs;
         i0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = i0 + lane;
        const bool live = i < npairs;
        const float4 c4 = live ? a.pair_p[i] : make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        const unsigned int h = live ? a.pair_h[i] : 0u;
        k2_exact_warp(a.h, live, (int)h, c4.x, c4.y, c4.z, __float_as_uint(c4.w), lane, s_rough, s_fine);
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < (unsigned int)a.h.nhalos; i += K2X_THREADS) {
        if (s_rough[i]) atomicAdd(a.h.counts + a.h.total_halos + (unsigned int)a.h.halo0 + i, (unsigned long long)s_rough[i]);
        if (s_fine[i]) atomicAdd(a.h.counts + (unsigned int)a.h.halo0 + i, (unsigned long long)s_fine[i]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(a.h.n_cand, want);
        atomicMax(a.misc + 3, want);
    }
}

/* ================================================================================================
 * K3  survivor ordering + gather (use case 2)
 * ================================================================================================ */
/* All K3 kernels run with fixed grids and read the record count on the device (min(*rec_count, rec_cap)), so the
 * whole use-case-2 pass is one enqueue with a single host synchronisation at the end. */
/* scatter records into per-halo segments (unordered inside a segment) */
constexpr int K3_CURSOR_STRIDE = 16;
struct BucketArgs {
    const Rec* recs;
    const unsigned long long* rec_count;
    unsigned long long rec_cap;
    const unsigned long long* seg_begin; /* [H] exclusive scan of counts */
    unsigned long long* cursor;          /* [H * K3_CURSOR_STRIDE] zeroed: one 128-byte line per halo (atomics on
                                            one line queue up in the L2) */
    Rec* out;
    int H;
};
/* More survivors than record slots (*rec_count > rec_cap): the host grows the buffers and runs the cutout again, so the
 * ordering kernels do nothing at all -- the per-halo segments would be partly filled with whatever an earlier cutout
 * left in the buffers, and following such records' halo / rank fields reads out of bounds.
 *
 * Records arrive in particle order, i.e. in random halo order: one global atomic per record on the 250 cursors of a
 * 250-halo pass serialises in the L2 (1.9 M records: 79 us, profiles/r02).  Every CTA therefore takes a contiguous
 * slice of the records, counts it per halo in shared memory, claims a range per (CTA, halo) with one global atomic,
 * and deals its records into those ranges (second read of the slice: L2 hits).  Halos beyond the shared-memory table
 * (H > K3B_HMAX) take the per-record atomic. */
constexpr int K3B_THREADS = 512;
constexpr int K3B_HMAX = 8192;
__global__ void __launch_bounds__(K3B_THREADS) k3_bucket(const BucketArgs a) {
    __shared__ unsigned int s_cnt[K3B_HMAX];   /* pass 1: records of the slice per halo; pass 2: next free offset */
    if (*a.rec_count > a.rec_cap) return;
    const unsigned long long nrec = *a.rec_count;
    const unsigned int tid = threadIdx.x;
    const unsigned long long per = (nrec + gridDim.x - 1) / gridDim.x;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * per, i1 = min(nrec, i0 + per);
    if (i0 >= i1) return;
    const bool table = a.H <= K3B_HMAX;
    const unsigned int lane = tid & 31u;
    /* the scans emit a halo's records in runs (a warp works on one halo at a time): when all 32 records of a warp's
     * load name the same halo, one shared-memory atomic stands for 32 conflicting ones */
    if (table) {
        for (unsigned int h = tid; h < (unsigned int)a.H; h += K3B_THREADS) s_cnt[h] = 0u;
        __syncthreads();
        for (unsigned long long i = i0 + tid; i - tid < i1; i += K3B_THREADS) {
            const bool live = i < i1;
            const unsigned int halo = live ? a.recs[i].halo : 0xffffffffu;
            const unsigned int first = __shfl_sync(0xffffffffu, halo, 0);
            if (__all_sync(0xffffffffu, halo == first)) {
                if (lane == 0 && live) atomicAdd(&s_cnt[halo], 32u);
            } else if (live) {
                atomicAdd(&s_cnt[halo], 1u);
            }
        }
        __syncthreads();
        for (unsigned int h = tid; h < (unsigned int)a.H; h += K3B_THREADS) {
            const unsigned int c = s_cnt[h];
            /* the CTA's range inside the halo's segment, as a 32-bit position in the output (record counts fit 32 bits:
             * the row -> record map is 32-bit) */
            s_cnt[h] = c ? (unsigned int)(a.seg_begin[h] + atomicAdd(a.cursor + (size_t)h * K3_CURSOR_STRIDE, (unsigned long long)c)) : 0u;
        }
        __syncthreads();
        for (unsigned long long i = i0 + tid; i - tid < i1; i += K3B_THREADS) {
            const bool live = i < i1;
            Rec r;
            r.halo = 0xffffffffu;
            if (live) r = a.recs[i];
            const unsigned int first = __shfl_sync(0xffffffffu, r.halo, 0);
            unsigned int p = 0;
            if (__all_sync(0xffffffffu, r.halo == first)) {
                if (lane == 0 && live) p = atomicAdd(&s_cnt[r.halo], 32u);
                p = __shfl_sync(0xffffffffu, p, 0) + lane;
            } else if (live) {
                p = atomicAdd(&s_cnt[r.halo], 1u);
            }
            if (live && p < a.rec_cap) a.out[p] = r;
        }
        return;
    }
    for (unsigned long long i = i0 + tid; i < i1; i += K3B_THREADS) {
        const Rec r = a.recs[i];
        const unsigned long long off = atomicAdd(a.cursor + (size_t)r.halo * K3_CURSOR_STRIDE, 1ull);
        const unsigned long long p = a.seg_begin[r.halo] + off;
        if (p < a.rec_cap) a.out[p] = r;
    }
}

/* exclusive scan of H counts (H is small: one CTA, serial chunks) */
__global__ void k3_scan_counts(const unsigned long long* counts, unsigned long long* seg_begin, int H) {
    __shared__ unsigned long long s_part[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5; /* 256 threads */
    const int per = (H + 255) / 256;
    unsigned long long sum = 0;
    for (int i = tid * per; i < min(H, (tid + 1) * per); ++i) sum += counts[i];
    unsigned long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    unsigned long long run = inc - sum;
    for (int w = 0; w < warp; ++w) run += s_part[w];
    for (int i = tid * per; i < min(H, (tid + 1) * per); ++i) { seg_begin[i] = run; run += counts[i]; }
    if (tid == 255) seg_begin[H] = run;
}

/* rank of every record inside its halo segment = its position in (theta_u, index) order, the order of the
 * reference's stable arg-sort (processLC.cpp:1214-1221).  Keys are unique (the low 32 bits are the particle index).
 * A segment is cut into chunks of K3_CHUNK records; ONE CTA orders a chunk in shared memory: theta_u of a cutout is
 * spread evenly over a narrow window, so the keys are first dealt into up to K3_NB buckets by the leading bits of
 * (theta_u - min) -- a monotonic map, ~8 keys per bucket -- with a counting sort, and each key then counts the
 * smaller keys of its own bucket.  (A bitonic network over the chunk did 10x the shared-memory traffic.)
 * Cluster-sized cutouts are a single chunk: the CTA stores, per OUTPUT row, which record goes there, so that
 * k3_gather_halo writes its rows coalesced.  For a longer segment it stores the record's rank inside its chunk and
 * the chunk's sorted keys, and k3_gather_halo adds, per record, the number of smaller keys in the OTHER chunks of
 * the segment (one binary search per chunk) and scatters.  grid = (chunk slots, halos). */
constexpr int K3_THREADS = 1024;
constexpr int K3_SORT_MAX = 8192; /* k3_small: records one CTA sorts */
constexpr int K3_CHUNK = 16384;   /* segments up to here are ordered by ONE CTA and gathered row-major; a thread keeps its
                                     16 keys in registers (all loads in flight at once), shared memory holds the grouped
                                     copy: 196 KB per CTA */
constexpr int K3_PER = K3_CHUNK / K3_THREADS;
constexpr int K3_NB = 1024;       /* buckets per chunk */
static_assert(K3_NB == K3_THREADS, "k3_rank scans one bucket count per thread");
constexpr size_t K3_SMEM_BYTES = (size_t)K3_CHUNK * (sizeof(unsigned long long) + 2 * sizeof(unsigned short)) +
                                 (size_t)(K3_NB + 1) * sizeof(unsigned int);
struct RankArgs {
    const unsigned long long* rec_count; /* > rec_cap: overflow, nothing to do */
    const Rec* recs;                     /* bucketed */
    const unsigned long long* seg_begin; /* [H+1] */
    unsigned long long rec_cap;
    unsigned int* rank;                  /* [nrec] multi-chunk segment: rank of the record inside its chunk */
    unsigned long long* sorted;          /* [nrec] sorted keys of every chunk of a multi-chunk segment */
    unsigned int* place;                 /* [nrec] single-chunk segment: the record (index into recs) of every output row */
    int H;
};
__global__ void __launch_bounds__(K3_THREADS) k3_rank(const RankArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long* s_grp = reinterpret_cast<unsigned long long*>(s_raw);           /* grouped by bucket */
    unsigned short* s_src = reinterpret_cast<unsigned short*>(s_grp + K3_CHUNK);        /* grouped: position as loaded */
    unsigned short* s_slot = s_src + K3_CHUNK;                                          /* as loaded: slot inside the bucket */
    unsigned int* s_start = reinterpret_cast<unsigned int*>(s_slot + K3_CHUNK);         /* [K3_NB + 1] */
    __shared__ unsigned int s_min, s_max, s_warp[K3_THREADS / 32];
    const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (*a.rec_count > a.rec_cap) return;
    for (int h = blockIdx.y; h < a.H; h += gridDim.y) {
        const unsigned long long b = a.seg_begin[h], e = min(a.seg_begin[h + 1], a.rec_cap);
        if (e <= b) continue;
        const unsigned long long len = e - b;
        for (unsigned long long c0 = (unsigned long long)blockIdx.x * K3_CHUNK; c0 < len; c0 += (unsigned long long)gridDim.x * K3_CHUNK) {
            const unsigned int m = (unsigned int)min((unsigned long long)K3_CHUNK, len - c0);
            if (tid == 0) { s_min = 0xffffffffu; s_max = 0u; }
            for (unsigned int i = tid; i <= K3_NB; i += K3_THREADS) s_start[i] = 0u;
            unsigned long long key[K3_PER];
            unsigned int lo = 0xffffffffu, hi = 0u;
#pragma unroll
            for (int k = 0; k < K3_PER; ++k) {
                const unsigned int i = tid + k * K3_THREADS;
                key[k] = ~0ull;
                if (i < m) key[k] = a.recs[b + c0 + i].key;
            }
#pragma unroll
            for (int k = 0; k < K3_PER; ++k) {
                const unsigned int i = tid + k * K3_THREADS;
                if (i < m) { lo = min(lo, (unsigned int)(key[k] >> 32)); hi = max(hi, (unsigned int)(key[k] >> 32)); }
            }
            __syncthreads();
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) { atomicMin(&s_min, lo); atomicMax(&s_max, hi); }
            __syncthreads();
            const unsigned int kmin = s_min;
            unsigned int shift = 0; /* bucket = (theta bits - min) >> shift, < K3_NB */
            while (((s_max - kmin) >> shift) >= (unsigned int)K3_NB) ++shift;
#pragma unroll
            for (int k = 0; k < K3_PER; ++k) {
                const unsigned int i = tid + k * K3_THREADS;
                if (i < m) {
                    const unsigned int bk = ((unsigned int)(key[k] >> 32) - kmin) >> shift;
                    s_slot[i] = (unsigned short)atomicAdd(&s_start[bk + 1], 1u); /* counts land one to the right */
                }
            }
            __syncthreads();
            /* inclusive scan of the counts: s_start[bk] = first position of bucket bk (one entry per thread,
             * K3_NB == K3_THREADS) */
            {
                unsigned int v = s_start[tid + 1];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, v, d);
                    if (lane >= (unsigned int)d) v += t;
                }
                if (lane == 31) s_warp[warp] = v;
                __syncthreads();
                if (warp == 0) {
                    unsigned int w = lane < K3_THREADS / 32 ? s_warp[lane] : 0u;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned int t = __shfl_up_sync(0xffffffffu, w, d);
                        if (lane >= (unsigned int)d) w += t;
                    }
                    if (lane < K3_THREADS / 32) s_warp[lane] = w;
                }
                __syncthreads();
                s_start[tid + 1] = v + (warp ? s_warp[warp - 1] : 0u);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K3_PER; ++k) {
                const unsigned int i = tid + k * K3_THREADS;
                if (i < m) {
                    const unsigned int p = s_start[((unsigned int)(key[k] >> 32) - kmin) >> shift] + s_slot[i];
                    s_grp[p] = key[k];
                    s_src[p] = (unsigned short)i;
                }
            }
            __syncthreads();
            for (unsigned int p = tid; p < m; p += K3_THREADS) {
                const unsigned long long k = s_grp[p];
                const unsigned int bk = ((unsigned int)(k >> 32) - kmin) >> shift;
                const unsigned int q0 = s_start[bk], q1 = s_start[bk + 1];
                unsigned int r = q0;
                for (unsigned int q = q0; q < q1; ++q) r += s_grp[q] < k;
                if (len <= K3_CHUNK) {
                    a.place[b + r] = (unsigned int)b + s_src[p];
                } else {
                    a.rank[b + c0 + s_src[p]] = r;
                    a.sorted[b + c0 + r] = k;
                }
            }
            __syncthreads();
        }
    }
}

struct GatherArgs {
    const Rec* recs;
    const unsigned int* place;        /* multi-chunk segments: record (inside the segment) of every row, from k3_place */
    const unsigned int* rank;
    const unsigned long long* seg_begin;
    const unsigned long long* rec_count;
    unsigned long long rec_cap;
    long long out_cap;
    ColsIn in;
    const Payload* payload; /* interleaved gather columns of a resident step, or NULL */
    ColsOut out;
    long long index_base;
    int pos_only;
};
/* multi-chunk segments: final row of every record = rank inside its chunk + smaller keys in the other chunks of the
 * segment (one binary search per other chunk); stored the other way round -- which record belongs to row o -- so that
 * the gather can run row-major (one scattered 4-byte store here instead of fourteen there) */
struct PlaceArgs {
    const Rec* recs;
    const unsigned int* rank;
    const unsigned long long* sorted;
    const unsigned long long* seg_begin;
    const unsigned long long* rec_count;
    unsigned long long rec_cap;
    unsigned int* place;                 /* the record (index into recs) of every output row */
    int H;
};
/* grid = (slices, segments): a single-chunk segment (nearly all of them in a job of cluster-sized cutouts) costs its
 * CTAs one read of two segment starts -- the kernel used to read every record to learn its segment -- and the records
 * of a multi-chunk segment are spread over gridDim.x CTAs (each is a chain of binary searches through the L2) */
__global__ void __launch_bounds__(256) k3_place(const PlaceArgs a) {
    if (*a.rec_count > a.rec_cap) return; /* overflow: see k3_bucket */
    for (int h = blockIdx.y; h < a.H; h += gridDim.y) {
        const unsigned long long sb = a.seg_begin[h];
        const unsigned long long len = a.seg_begin[h + 1] - sb;
        if (len <= K3_CHUNK) continue; /* single chunk: k3_rank already stored row -> record */
        for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < len;
             j += (unsigned long long)gridDim.x * blockDim.x) {
            const unsigned long long key = a.recs[sb + j].key;
            unsigned long long o = a.rank[sb + j];
            const unsigned long long mine = j / K3_CHUNK;
            for (unsigned long long c = 0; c * K3_CHUNK < len; ++c) {
                if (c == mine) continue;
                const unsigned long long* ck = a.sorted + sb + c * K3_CHUNK;
                unsigned int lo = 0, hi = (unsigned int)min((unsigned long long)K3_CHUNK, len - c * K3_CHUNK);
                while (lo < hi) {
                    const unsigned int mid = (lo + hi) >> 1;
                    if (ck[mid] < key) lo = mid + 1; else hi = mid;
                }
                o += lo;
            }
            a.place[sb + o] = (unsigned int)(sb + j);
        }
    }
}

/* thread = output row: rows are written coalesced */
__global__ void k3_gather_halo(const GatherArgs a) {
    if (*a.rec_count > a.rec_cap) return; /* overflow: see k3_bucket */
    const unsigned long long nrec = *a.rec_count;
    for (unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; o < nrec;
         o += (unsigned long long)gridDim.x * blockDim.x) {
        const Rec r = a.recs[a.place[o]]; /* row -> record: k3_rank (single-chunk segments) or k3_place */
        if ((long long)o >= a.out_cap) continue;
        const long long g = (long long)(r.key & 0xffffffffull);
        a.out.theta[o] = r.th;                        /* processLC.cpp:1446-1447 */
        a.out.phi[o] = r.ph;
        a.out.x[o] = r.x;
        a.out.y[o] = r.y;
        a.out.z[o] = r.z;
        if (a.payload) {
            const Payload p = load_payload(a.payload + g);
            a.out.redshift[o] = a_to_z_dev(p.a);      /* :1450 */
            a.out.id[o] = p.id;
            if (!a.pos_only) { a.out.vx[o] = p.vx; a.out.vy[o] = p.vy; a.out.vz[o] = p.vz; a.out.rot[o] = p.rot; a.out.rep[o] = p.rep; }
        } else if (a.in.a) {
            a.out.redshift[o] = a_to_z_dev(a.in.a[g]);
            a.out.id[o] = a.in.id[g];
            if (!a.pos_only) {                        /* :1455-1461 */
                a.out.vx[o] = a.in.vx[g];
                a.out.vy[o] = a.in.vy[g];
                a.out.vz[o] = a.in.vz[g];
                a.out.rot[o] = a.in.rot[g];
                a.out.rep[o] = a.in.rep[g];
            }
        }
        a.out.theta_u[o] = __int_as_float((int)(r.key >> 32));
        a.out.index[o] = a.index_base + g;
    }
}

/* ---- count exchange by peer stores (NVLink / NVSwitch), no library launch ---------------------------------------
 * Every GPU owns a mailbox in its own HBM: [2 parities][R ranks][1 epoch word + XCHG_PAYLOAD counters].  To exchange,
 * a GPU STORES its {counts, rough counts} into slot [parity][its rank] of every peer's mailbox (the peers' mailboxes
 * are mapped through CUDA IPC or peer access, so these are plain st.global over NVLink), fences, publishes the epoch
 * with a release store, then waits until all R slots of its OWN mailbox carry the epoch, and computes its row offsets
 * = sum of the counts of lower ranks.  One single-CTA kernel enqueued behind the cutout kernels: the exchange costs
 * one NVLink store round instead of a collective launch (~3 us instead of ~40 us of a 250 us step).
 * Parity double-buffering is enough: a rank can only start exchange e+1 after every peer has posted e, and a peer
 * posts e+1 only after it finished reading e. */
constexpr int XCHG_PAYLOAD = 4096;              /* u64 counters per slot: 2 x halos */
constexpr int XCHG_SLOT = XCHG_PAYLOAD + 8;     /* epoch word (+ padding to keep the payload 64-byte aligned) */
constexpr int XCHG_MAX_RANKS = 32;

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct XchgArgs {
    const unsigned long long* send;      /* [n] this GPU's {counts[H], rough[H]} */
    unsigned long long* peer[XCHG_MAX_RANKS]; /* every rank's mailbox as seen from this GPU (peer[rank] = own) */
    unsigned long long* gathered;        /* [R][n] out */
    unsigned long long* offsets;         /* [H] out: sum of counts of lower ranks */
    unsigned long long* status;          /* [0] set to 1 on time-out */
    unsigned long long* host_gathered;   /* pinned host copies of gathered (+ the status word behind it) and offsets */
    unsigned long long* host_offsets;
    int uc1;                             /* use case 1: send {send[0], 0} (send points at the running survivor total) */
    unsigned long long epoch;
    unsigned long long timeout_ns;
    int n, H, R, rank;
};

/* post: my counters into my slot of every mailbox, then the epoch with a release store (all threads of the CTA) */
__device__ __forceinline__ void xchg_post(const XchgArgs& a) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long parity = a.epoch & 1ull;
    const size_t my_slot = ((size_t)parity * a.R + a.rank) * XCHG_SLOT;
    for (int p = 0; p < a.R; ++p) {
        unsigned long long* dst = a.peer[p] + my_slot + 8;
        for (int i = tid; i < a.n; i += nt) st_relaxed_sys_u64(dst + i, (a.uc1 && i > 0) ? 0ull : a.send[i]);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.R) st_release_sys_u64(a.peer[tid] + my_slot, a.epoch);
}
/* wait for every rank's post in MY mailbox, hand everything to the host (pinned memory) and compute my row offsets
 * (all threads of the CTA; s_timeout: one shared int) */
__device__ __forceinline__ void xchg_finish(const XchgArgs& a, int* s_timeout) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long parity = a.epoch & 1ull;
    if (tid == 0) *s_timeout = 0;
    __syncthreads();
    unsigned long long* mine = a.peer[a.rank];
    if (tid < a.R) {
        const unsigned long long* flag = mine + ((size_t)parity * a.R + tid) * XCHG_SLOT;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys_u64(flag) != a.epoch) {
            if (global_timer_ns() - t0 > a.timeout_ns) { *s_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (*s_timeout) {
        if (tid == 0) { *a.status = 1ull; a.host_gathered[(size_t)a.R * a.n] = 1ull; __threadfence_system(); }
        return;
    }
    for (int r = 0; r < a.R; ++r) {
        const unsigned long long* src = mine + ((size_t)parity * a.R + r) * XCHG_SLOT + 8;
        for (int i = tid; i < a.n; i += nt) {
            const unsigned long long v = ld_relaxed_sys_u64(src + i);
            a.gathered[(size_t)r * a.n + i] = v;
            a.host_gathered[(size_t)r * a.n + i] = v;
        }
    }
    if (tid == 0) a.host_gathered[(size_t)a.R * a.n] = 0ull; /* status: no time-out */
    for (int h = tid; h < a.H; h += nt) {
        unsigned long long run = 0;
        for (int r = 0; r < a.rank; ++r) run += ld_relaxed_sys_u64(mine + ((size_t)parity * a.R + r) * XCHG_SLOT + 8 + h);
        a.offsets[h] = run;
        a.host_offsets[h] = run;
    }
    /* the host reads its copies after the stream has drained (cc_exchange_counts): no system-wide fence */
}

__global__ void __launch_bounds__(256) k4_exchange_peer(const XchgArgs a) {
    __shared__ int s_timeout;
    xchg_post(a);
    xchg_finish(a, &s_timeout);
}

/* ---- small cutouts: the four kernels above as ONE single-CTA launch -------------------------------------------
 * A cluster-sized cutout leaves a few hundred survivors; k3_scan_counts + k3_bucket + k3_rank + k3_gather_halo are
 * then four launch-latency-bound kernels (~20 us of a 250 us step).  When the previous cutout on the context left at
 * most K3_SORT_MAX survivors the host speculatively enqueues this kernel instead: one CTA scans the halo counts,
 * sorts ALL records by (halo, theta_u, index) with a bitonic network in shared memory -- the sorted position IS the
 * output row -- and gathers.  If the record count turns out larger it sets misc[4] and does nothing; the host then
 * re-runs the cutout through the general kernels. */
constexpr int K3S_THREADS = 512;
constexpr int K3S_RANK_MAX = 1024; /* up to here rows come from counting, beyond from the bitonic network */
constexpr size_t K3S_STAGE_OFF = (size_t)K3_SORT_MAX * (sizeof(unsigned long long) + 2 * sizeof(unsigned short)) +
                                 (size_t)K3S_THREADS * sizeof(unsigned long long);
/* counting path: every record and the columns gathered for it wait in shared memory at the record's ROW (64 B each) */
struct __align__(16) K3sRow { Rec rec; Payload cols; };
constexpr size_t K3S_SMEM_BYTES = K3S_STAGE_OFF + (size_t)K3S_RANK_MAX * sizeof(K3sRow);
struct SmallArgs {
    GatherArgs g;                        /* recs = the UNbucketed records (rec_a) */
    const unsigned long long* counts;    /* [H] survivors per halo */
    unsigned long long* seg_begin;       /* [H+1] out */
    unsigned long long* misc;            /* [4] = 1 when the record count does not fit */
    unsigned long long* host_out;        /* pinned host memory: misc[8], counts[2H], seg_begin[H+1] -- what the host reads */
    unsigned int state_elems;            /* u64 words of the device state starting at misc: zeroed for the next cutout */
    unsigned long long* xchg_send;       /* [2H] {counts, rough counts} kept for the count exchange (capi.cu) */
    int H;
    unsigned long long seq;              /* written to host_out[7] once everything the host reads is on its way: the host
                                            spins on it instead of waiting for the stream to drain (capi.cu wait_small) */
    int do_xchg;                         /* the count exchange over the peer mailboxes is part of this kernel: posted first
                                            (the counts are final), collected after the rows are written */
    XchgArgs x;
};
/* last step of k3_small (all threads): results to the host, device state zeroed */
__device__ __forceinline__ void k3s_publish(const SmallArgs& a) {
    __syncthreads();
    const unsigned int nout = 8u + 3u * (unsigned int)a.H + 1u;
    for (unsigned int i = threadIdx.x; i < nout; i += K3S_THREADS) a.host_out[i] = a.misc[i];
    for (unsigned int i = threadIdx.x; i < 2u * (unsigned int)a.H; i += K3S_THREADS) a.xchg_send[i] = a.counts[i];
    /* one thread makes the CTA's stores to the host visible (the barrier orders them before its fence) and raises the
     * flag the host spins on; the others go on zeroing the device state */
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(a.host_out + 7) = a.seq;
    }
    for (unsigned int i = threadIdx.x; i < a.state_elems; i += K3S_THREADS) a.misc[i] = 0ull;
}
__global__ void __launch_bounds__(K3S_THREADS) k3_small(const SmallArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_raw);
    unsigned short* s_halo = reinterpret_cast<unsigned short*>(s_raw + (size_t)K3_SORT_MAX * 8);
    unsigned short* s_idx = s_halo + K3_SORT_MAX;
    unsigned long long* s_part = reinterpret_cast<unsigned long long*>(s_raw + (size_t)K3_SORT_MAX * 12);
    const unsigned int tid = threadIdx.x;
    __shared__ int s_xchg_timeout;
    /* one CTA on one SM: the kernel is a chain of memory round trips, so everything that can travel together does.
     * The record count, the first K3S_RANK_MAX records (speculatively: the buffer is at least that large or the load
     * is skipped) and -- exchange=eager -- the NVLink stores of the counts to every peer all leave before anything
     * waits; the peers' posts arrive while this CTA orders and gathers. */
    const unsigned long long nrec = *a.g.rec_count;
    K3sRow* s_row = reinterpret_cast<K3sRow*>(s_raw + K3S_STAGE_OFF);
    constexpr int PER = K3S_RANK_MAX / K3S_THREADS;
    Rec rec[PER];
    Payload cols[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const unsigned int i = tid + r * K3S_THREADS;
        rec[r] = Rec{0ull, 0u, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (i < a.g.rec_cap) rec[r] = a.g.recs[i];
    }
    if (a.do_xchg) xchg_post(a.x);
    if (nrec > K3_SORT_MAX || nrec > a.g.rec_cap) {
        if (tid == 0) a.misc[4] = 1ull;
        if (a.do_xchg) xchg_finish(a.x, &s_xchg_timeout);
        k3s_publish(a);
        return;
    }
    const bool ranked = nrec <= K3S_RANK_MAX;
    const unsigned int n8 = ((unsigned int)nrec + 7u) & ~7u;
    if (ranked) {
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const unsigned int i = tid + r * K3S_THREADS;
            if (i < nrec) {
                s_key[i] = rec[r].key;
                s_halo[i] = (unsigned short)rec[r].halo;
            } else if (i < n8) { s_key[i] = ~0ull; s_halo[i] = 0xffffu; }
        }
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const unsigned int i = tid + r * K3S_THREADS;
            cols[r] = Payload{0.f, 0.f, 0.f, 0.f, 0ll, 0, 0};
            if (i < nrec) {
                const long long g = (long long)(rec[r].key & 0xffffffffull);
                if (a.g.payload) cols[r] = load_payload(a.g.payload + g);
                else if (a.g.in.a) {
                    cols[r].a = a.g.in.a[g];
                    cols[r].id = a.g.in.id[g];
                    if (!a.g.pos_only) {
                        cols[r].vx = a.g.in.vx[g]; cols[r].vy = a.g.in.vy[g]; cols[r].vz = a.g.in.vz[g];
                        cols[r].rot = a.g.in.rot[g]; cols[r].rep = a.g.in.rep[g];
                    }
                }
            }
        }
    }
    /* exclusive scan of the halo counts -> seg_begin (read back by the host) */
    {
        const int per = (a.H + K3S_THREADS - 1) / K3S_THREADS;
        unsigned long long sum = 0;
        for (int i = tid * per; i < min(a.H, (int)(tid + 1) * per); ++i) sum += a.counts[i];
        /* block-wide exclusive scan of the per-thread sums: shuffles inside the warp, then the 16 warp totals
         * (a serial loop over 512 shared-memory entries here cost 6 us of a 220 us step) */
        const unsigned int lane = tid & 31u, warp = tid >> 5;
        unsigned long long inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned int)d) inc += t;
        }
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = lane < K3S_THREADS / 32 ? s_part[lane] : 0ull;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= (unsigned int)d) w += t;
            }
            if (lane < K3S_THREADS / 32) s_part[lane] = w; /* inclusive warp totals */
        }
        __syncthreads();
        unsigned long long run = inc - sum + (warp ? s_part[warp - 1] : 0ull);
        for (int i = tid * per; i < min(a.H, (int)(tid + 1) * per); ++i) { a.seg_begin[i] = run; run += a.counts[i]; }
        if (tid == K3S_THREADS - 1) a.seg_begin[a.H] = run;
    }
    if (ranked) {
        /* a cluster-sized cutout (a few hundred records): every record counts the records in front of it -- (halo, key)
         * pairs are unique -- and that count IS its row.  The kernel is one chain of dependent memory round trips on a
         * single SM, so the chain is kept short: the thread that owns a record loads it once, asks for its gather
         * columns BEFORE the counting loop (their DRAM latency hides behind it), and parks both in shared memory at the
         * record's row; the row-major writer below then touches no global memory but the output. */
        __syncthreads();
        unsigned int row[PER];
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const unsigned int i = tid + r * K3S_THREADS;
            unsigned int before = 0;
            if (i < nrec) {
                const unsigned long long ki = rec[r].key;
                if (a.H == 1) {
                    /* one halo: keys alone order the records; eight independent comparisons per trip */
                    for (unsigned int j = 0; j < n8; j += 8) {
                        const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(s_key + j);
                        const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(s_key + j + 2);
                        const ulonglong2 k2 = *reinterpret_cast<const ulonglong2*>(s_key + j + 4);
                        const ulonglong2 k3 = *reinterpret_cast<const ulonglong2*>(s_key + j + 6);
                        before += (unsigned int)(k0.x < ki) + (unsigned int)(k0.y < ki) + (unsigned int)(k1.x < ki) +
                                  (unsigned int)(k1.y < ki) + (unsigned int)(k2.x < ki) + (unsigned int)(k2.y < ki) +
                                  (unsigned int)(k3.x < ki) + (unsigned int)(k3.y < ki);
                    }
                } else {
                    const unsigned int hi = rec[r].halo & 0xffffu;
                    for (unsigned int j = 0; j < n8; j += 4) {
                        const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(s_key + j);
                        const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(s_key + j + 2);
                        const uint2 hh = *reinterpret_cast<const uint2*>(s_halo + j);
                        const unsigned int h0 = hh.x & 0xffffu, h1 = hh.x >> 16, h2 = hh.y & 0xffffu, h3 = hh.y >> 16;
                        before += (unsigned int)((h0 < hi) | ((h0 == hi) & (k0.x < ki))) +
                                  (unsigned int)((h1 < hi) | ((h1 == hi) & (k0.y < ki))) +
                                  (unsigned int)((h2 < hi) | ((h2 == hi) & (k1.x < ki))) +
                                  (unsigned int)((h3 < hi) | ((h3 == hi) & (k1.y < ki)));
                    }
                }
            }
            row[r] = before;
        }
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const unsigned int i = tid + r * K3S_THREADS;
            if (i < nrec) { s_row[row[r]].rec = rec[r]; s_row[row[r]].cols = cols[r]; }
        }
        __syncthreads();
        for (unsigned int o = tid; o < nrec; o += K3S_THREADS) {
            if ((long long)o >= a.g.out_cap) continue;
            const Rec r = s_row[o].rec;
            const Payload p = s_row[o].cols;
            a.g.out.theta[o] = r.th;
            a.g.out.phi[o] = r.ph;
            a.g.out.x[o] = r.x;
            a.g.out.y[o] = r.y;
            a.g.out.z[o] = r.z;
            if (a.g.payload || a.g.in.a) {
                a.g.out.redshift[o] = a_to_z_dev(p.a);
                a.g.out.id[o] = p.id;
                if (!a.g.pos_only) { a.g.out.vx[o] = p.vx; a.g.out.vy[o] = p.vy; a.g.out.vz[o] = p.vz; a.g.out.rot[o] = p.rot; a.g.out.rep[o] = p.rep; }
            }
            a.g.out.theta_u[o] = __int_as_float((int)(r.key >> 32));
            a.g.out.index[o] = a.g.index_base + (long long)(r.key & 0xffffffffull);
        }
        if (a.do_xchg) xchg_finish(a.x, &s_xchg_timeout);
        k3s_publish(a);
        return;
    }
    unsigned int n = 32;
    while (n < nrec) n <<= 1;
    for (unsigned int i = tid; i < n; i += K3S_THREADS) {
        if (i < nrec) { s_key[i] = a.g.recs[i].key; s_halo[i] = (unsigned short)a.g.recs[i].halo; }
        else { s_key[i] = ~0ull; s_halo[i] = 0xffffu; }
        s_idx[i] = (unsigned short)i;
    }
    __syncthreads();
    for (unsigned int k = 2; k <= n; k <<= 1) {
        for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            for (unsigned int t = tid; t < (n >> 1); t += K3S_THREADS) {
                const unsigned int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const unsigned long long ki = s_key[i], kl = s_key[l];
                const unsigned short hi = s_halo[i], hl = s_halo[l];
                const bool greater = (hi != hl) ? (hi > hl) : (ki > kl);
                const bool up = (i & k) == 0;
                if (greater == up) {
                    s_key[i] = kl; s_key[l] = ki;
                    s_halo[i] = hl; s_halo[l] = hi;
                    const unsigned short t2 = s_idx[i]; s_idx[i] = s_idx[l]; s_idx[l] = t2;
                }
            }
            __syncthreads();
        }
    }
    /* sorted position = output row (segments are contiguous in halo order) */
    for (unsigned int o = tid; o < nrec; o += K3S_THREADS) {
        if ((long long)o >= a.g.out_cap) continue;
        const Rec r = a.g.recs[s_idx[o]];
        const long long g = (long long)(r.key & 0xffffffffull);
        a.g.out.theta[o] = r.th;
        a.g.out.phi[o] = r.ph;
        a.g.out.x[o] = r.x;
        a.g.out.y[o] = r.y;
        a.g.out.z[o] = r.z;
        if (a.g.payload) {
            const Payload p = load_payload(a.g.payload + g);
            a.g.out.redshift[o] = a_to_z_dev(p.a);
            a.g.out.id[o] = p.id;
            if (!a.g.pos_only) { a.g.out.vx[o] = p.vx; a.g.out.vy[o] = p.vy; a.g.out.vz[o] = p.vz; a.g.out.rot[o] = p.rot; a.g.out.rep[o] = p.rep; }
        } else if (a.g.in.a) {
            a.g.out.redshift[o] = a_to_z_dev(a.g.in.a[g]);
            a.g.out.id[o] = a.g.in.id[g];
            if (!a.g.pos_only) {
                a.g.out.vx[o] = a.g.in.vx[g];
                a.g.out.vy[o] = a.g.in.vy[g];
                a.g.out.vz[o] = a.g.in.vz[g];
                a.g.out.rot[o] = a.g.in.rot[g];
                a.g.out.rep[o] = a.g.in.rep[g];
            }
        }
        a.g.out.theta_u[o] = __int_as_float((int)(r.key >> 32));
        a.g.out.index[o] = a.g.index_base + g;
    }
    if (a.do_xchg) xchg_finish(a.x, &s_xchg_timeout);
    k3s_publish(a);
}

/* ================================================================================================
 * K4  offsets from the all-gathered counts: offsets[h] = sum_{r < rank} counts[r][h]
 *     gathered layout: [rank][2*H] (counts then rough counts)
 * ================================================================================================ */
__global__ void k4_offsets(const unsigned long long* gathered, unsigned long long* offsets, int H, int rank) {
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
        unsigned long long run = 0;
        for (int r = 0; r < rank; ++r) run += gathered[(size_t)r * 2 * H + h];
        offsets[h] = run;
    }
}

/* ================================================================================================
 * synthetic lightcone generator (counter-based; cc_synth_host in capi.cu computes the same bits)
 * ================================================================================================ */
__host__ __device__ inline unsigned long long synth_mix(unsigned long long seed, unsigned long long col,
                                                        unsigned long long g) {
    unsigned long long z = seed + 0x9e3779b97f4a7c15ull * (g + 1) + 0xd1b54a32d192ed03ull * (col + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
/* uniform in (0,1): (2k+1) * 2^-24, k in [0, 2^23) -- exact in float, never 0 or 1 */
__host__ __device__ inline float synth_unit(unsigned long long h) {
    return (float)(2u * (unsigned int)(h >> 41) + 1u) * 5.9604644775390625e-08f;
}
struct SynthCols {
    float *x, *y, *z, *vx, *vy, *vz, *a;
    long long* id;
    int *rot, *rep;
};
__host__ __device__ inline void synth_one(unsigned long long seed, long long g, int kind, float box, float& x,
                                          float& y, float& z, float& vx, float& vy, float& vz, float& a,
                                          long long& id, int& rot, int& rep) {
    float u[7];
    for (int c = 0; c < 7; ++c) u[c] = synth_unit(synth_mix(seed, (unsigned long long)c, (unsigned long long)g));
    if (kind == 0) { /* first octant */
        x = CC_FMUL(u[0], box); y = CC_FMUL(u[1], box); z = CC_FMUL(u[2], box);
    } else { /* full cube: (2u-1) is exact */
        x = CC_FMUL(CC_FSUB(CC_FMUL(2.0f, u[0]), 1.0f), box);
        y = CC_FMUL(CC_FSUB(CC_FMUL(2.0f, u[1]), 1.0f), box);
        z = CC_FMUL(CC_FSUB(CC_FMUL(2.0f, u[2]), 1.0f), box);
    }
    vx = CC_FMUL(CC_FSUB(u[3], 0.5f), 600.0f);
    vy = CC_FMUL(CC_FSUB(u[4], 0.5f), 600.0f);
    vz = CC_FMUL(CC_FSUB(u[5], 0.5f), 600.0f);
    a = CC_FADD(0.5f, CC_FMUL(0.5f, u[6]));
    const unsigned long long hh = synth_mix(seed, 7ull, (unsigned long long)g);
    id = g;
    rot = (int)((hh >> 33) % 6ull);
    rep = (int)(hh & 0xfffffull);
}
__global__ void k_synth(SynthCols c, long long n, unsigned long long seed, long long index_base, int kind, float box) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float x, y, z, vx, vy, vz, a;
        long long id;
        int rot, rep;
        synth_one(seed, index_base + i, kind, box, x, y, z, vx, vy, vz, a, id, rot, rep);
        c.x[i] = x; c.y[i] = y; c.z[i] = z;
        if (c.vx) { c.vx[i] = vx; c.vy[i] = vy; c.vz[i] = vz; c.rot[i] = rot; c.rep[i] = rep; }
        c.a[i] = a; c.id[i] = id;
    }
}

/* ================================================================================================
 * test hook: element-wise evaluation of the device restatement
 * ================================================================================================ */
__global__ void k_debug_eval(int op, long long n, const float* in0, const float* in1, const float* in2, float* out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float r;
        switch (op) {
            case 0: r = ccref::acosf_ref(in0[i]); break;
            case 1: r = ccref::atanf_ref(in0[i]); break;
            case 2: r = ccref::theta_arcsec(in0[i], in1[i], in2[i]); break;
            case 3: r = ccref::phi_arcsec_plain(in0[i], in1[i]); break;
            case 4: r = ccref::phi_arcsec_guarded(in0[i], in1[i]); break;
            case 5: r = a_to_z_dev(in0[i]); break;
            default: r = ccref::to_arcsec_fast(in0[i]); break; /* 6: the device's division-free arcsecond scaling */
        }
        out[i] = r;
    }
}

}  // namespace cck
