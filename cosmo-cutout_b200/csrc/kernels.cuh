/*
 * kernels.cuh -- sm_100a kernels of the lightcone cutout path.
 *
 *   k1_cutout_const   use case 1: fused octant test -> pre-filter -> exact (theta,phi) -> window test ->
 *                     order-preserving single-pass compaction (warp ballot/popc + decoupled look-back over
 *                     tiles) -> 12-column gather.            replaces processLC.cpp:276-310
 *   k2_cutout_halo    use case 2: one pass over x,y,z for a batch of halos: pre-filter -> exact un-rotated
 *                     (theta_u,phi_u) -> rough test -> R.v -> exact rotated angles -> fine test -> survivor
 *                     records.                               replaces processLC.cpp:897-909, :1334-1340, :1383-1481
 *   k3_*              survivor ordering by (theta_u, input index) per halo
 *                                                            replaces processLC.cpp:1214-1221 (restricted to survivors)
 *   k3_gather_halo    column gather of the ordered survivors replaces processLC.cpp:1446-1461
 *   k4_offsets        exclusive scan over ranks of the all-gathered counts
 *                                                            replaces processLC.cpp:331-337, :1569-1583
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

/* x86 SSE quiets and propagates a NaN operand's payload; the GPU returns the canonical NaN.  redshift is
 * the only output column computed from a possibly-NaN input, so mirror the host there. */
__device__ __forceinline__ float a_to_z_dev(float a) {
    if (a != a) return __int_as_float(__float_as_int(a) | 0x00400000);
    return ccref::a_to_z(a);
}

/* ================================================================================================
 * K1  use case 1: warp-autonomous tiles
 *
 * Every warp owns tiles of 256 consecutive particles (8 per lane: two coalesced float4 chunks per column) and
 * runs the whole chain for a tile by itself -- load, pre-filter, ballot/popc compaction of the candidates into
 * its private shared-memory queue, exact reference arithmetic on the queued candidates, ballot/popc compaction
 * of the survivors, decoupled look-back over the per-tile states for the global row offset, 12-column gather.
 * There is no block-level barrier anywhere, so a warp stalled on a long-latency phase (fp64 divide chain, the
 * scattered gather loads, a look-back wait) never holds up the 30+ other warps of its SM that keep HBM busy.
 *
 * Tiles are assigned statically (tile = iteration * total_warps + global_warp): the look-back of tile t only
 * waits on tiles < t, which belong to warps that are resident and never wait on higher tiles -- provided every
 * warp of the grid IS resident, which is why the kernel is launched cooperatively with an occupancy-sized grid.
 * ================================================================================================ */
constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_TILE = 256; /* particles per warp tile */
constexpr unsigned long long TS_AGG = 1ull << 62;
constexpr unsigned long long TS_PREFIX = 2ull << 62;
constexpr unsigned long long TS_MASK = (1ull << 62) - 1;

struct ConstArgs {
    ColsIn in;
    ColsOut out;
    long long n;          /* particles in this shard */
    long long cap;        /* rows available in `out` */
    long long index_base; /* global index of in[0] */
    float th_lo, th_hi, ph_lo, ph_hi; /* exact, strict (processLC.cpp:287-288) */
    float chi2, clo2, tlo, thi;       /* pre-filter (prefilter.h) */
    unsigned long long* tile_state;   /* [ntiles], zeroed before launch */
    unsigned long long* result;       /* [0] survivors so far: read as the row base of tile 0 (zero for a resident
                                         step, the running total for a streamed one), written back by the last
                                         tile; [1] pre-filter candidates */
    unsigned int ntiles;
};

__device__ __forceinline__ bool k1_prefilter(float x, float y, float z, const ConstArgs& a) {
    const float mn = fminf(x, fminf(y, z));
    const float mx = fmaxf(x, fmaxf(y, z));
    if (!(mn > 0.0f)) return false;                            /* octant test, processLC.cpp:279 */
    if (!(mn >= CC_SANE_LO && mx <= CC_SANE_HI)) return true;  /* let the exact path decide */
    const float r2 = fmaf(z, z, fmaf(y, y, x * x));
    const float z2 = z * z;
    return (z2 < a.chi2 * r2) & (z2 > a.clo2 * r2) & (y > a.tlo * x) & (y < a.thi * x);
}

struct K1WarpQueue { /* one per warp */
    float x[K1_TILE], y[K1_TILE], z[K1_TILE], th[K1_TILE], ph[K1_TILE];
    unsigned char idx[K1_TILE];
};

template <bool VEC>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_cutout_const(const ConstArgs a) {
    __shared__ K1WarpQueue s_q[K1_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K1WarpQueue& q = s_q[warp];
    const unsigned int total_warps = gridDim.x * K1_WARPS;
    const unsigned int lane_lt = (1u << lane) - 1u;
    unsigned long long n_cand = 0;

    for (unsigned int tile = blockIdx.x * K1_WARPS + warp; tile < a.ntiles; tile += total_warps) {
        const long long base = (long long)tile * K1_TILE;
        /* ---- load 8 particles: chunk A = [base+4*lane, +4), chunk B = [base+128+4*lane, +4) ---- */
        float px[8], py[8], pz[8];
        unsigned int valid = 0xffu;
        if (VEC && base + K1_TILE <= a.n) {
            const float4 xa = ldg_stream4(a.in.x + base + 4 * lane), xb = ldg_stream4(a.in.x + base + 128 + 4 * lane);
            const float4 ya = ldg_stream4(a.in.y + base + 4 * lane), yb = ldg_stream4(a.in.y + base + 128 + 4 * lane);
            const float4 za = ldg_stream4(a.in.z + base + 4 * lane), zb = ldg_stream4(a.in.z + base + 128 + 4 * lane);
            px[0] = xa.x; px[1] = xa.y; px[2] = xa.z; px[3] = xa.w; px[4] = xb.x; px[5] = xb.y; px[6] = xb.z; px[7] = xb.w;
            py[0] = ya.x; py[1] = ya.y; py[2] = ya.z; py[3] = ya.w; py[4] = yb.x; py[5] = yb.y; py[6] = yb.z; py[7] = yb.w;
            pz[0] = za.x; pz[1] = za.y; pz[2] = za.z; pz[3] = za.w; pz[4] = zb.x; pz[5] = zb.y; pz[6] = zb.z; pz[7] = zb.w;
        } else {
            valid = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long g = base + (j >> 2) * 128 + 4 * lane + (j & 3);
                if (g < a.n) {
                    px[j] = __ldcs(a.in.x + g); py[j] = __ldcs(a.in.y + g); pz[j] = __ldcs(a.in.z + g);
                    valid |= 1u << j;
                } else {
                    px[j] = py[j] = pz[j] = 0.0f;
                }
            }
        }
        /* ---- pre-filter ---- */
        unsigned int m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) m |= (unsigned)k1_prefilter(px[j], py[j], pz[j], a) << j;
        m &= valid;
        /* ---- ordered candidate compaction into the warp's queue (order: chunk, lane, element = input order) ---- */
        const unsigned int cnt = __popc(m & 0xfu) | (__popc(m >> 4) << 16);
        unsigned int inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const unsigned int total = __shfl_sync(0xffffffffu, inc, 31);
        const unsigned int totalA = total & 0xffffu, C = totalA + (total >> 16);
        unsigned long long exclusive = 0;
        unsigned int bits[K1_TILE / 32];
        unsigned int K = 0;
        if (C) { /* warp-uniform */
            const unsigned int ex = inc - cnt;
            unsigned int posA = ex & 0xffffu, posB = totalA + (ex >> 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (m & (1u << j)) {
                    q.x[posA] = px[j]; q.y[posA] = py[j]; q.z[posA] = pz[j];
                    q.idx[posA] = (unsigned char)(4 * lane + j);
                    ++posA;
                }
#pragma unroll
            for (int j = 4; j < 8; ++j)
                if (m & (1u << j)) {
                    q.x[posB] = px[j]; q.y[posB] = py[j]; q.z[posB] = pz[j];
                    q.idx[posB] = (unsigned char)(128 + 4 * lane + (j - 4));
                    ++posB;
                }
            __syncwarp();
            /* ---- exact path on the densely packed candidates: the reference's arithmetic ---- */
#pragma unroll
            for (int w = 0; w < K1_TILE / 32; ++w) {
                bits[w] = 0;
                if ((unsigned)(w * 32) < C) {
                    const unsigned int j = w * 32 + lane;
                    bool pass = false;
                    if (j < C) {
                        const float x = q.x[j], y = q.y[j], z = q.z[j];
                        if (x > 0.0f && y > 0.0f && z > 0.0f) { /* processLC.cpp:279 */
                            const float th = ccref::theta_arcsec(x, y, z);
                            const float ph = ccref::phi_arcsec_plain(x, y);
                            pass = th > a.th_lo && th < a.th_hi && ph > a.ph_lo && ph < a.ph_hi; /* :287-288 */
                            q.th[j] = th;
                            q.ph[j] = ph;
                        }
                    }
                    bits[w] = __ballot_sync(0xffffffffu, pass);
                    K += __popc(bits[w]);
                }
            }
            n_cand += C;
        }
        /* ---- decoupled look-back: row offset of this tile's first survivor ---- */
        if (tile == 0) {
            exclusive = ld_relaxed_u64(a.result); /* rows written by earlier chunks of a streamed step */
            if (lane == 0) st_relaxed_u64(a.tile_state, TS_PREFIX | (exclusive + K));
        } else {
            if (lane == 0) st_relaxed_u64(a.tile_state + tile, TS_AGG | (unsigned long long)K);
            long long look = (long long)tile - 1;
            while (true) {
                const long long idx = look - lane;
                unsigned long long st = TS_PREFIX; /* before tile 0: prefix 0 (tile 0 itself carries the base) */
                if (idx >= 0) {
                    st = ld_relaxed_u64(a.tile_state + idx);
                    while ((st >> 62) == 0) st = ld_relaxed_u64(a.tile_state + idx);
                }
                const unsigned int pmask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                const int first = pmask ? (__ffs(pmask) - 1) : 31;
                unsigned long long v = (lane <= first) ? (st & TS_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                exclusive += v;
                if (pmask) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_u64(a.tile_state + tile, TS_PREFIX | (exclusive + K));
        }
        if (tile == a.ntiles - 1 && lane == 0) a.result[0] = exclusive + K;
        /* ---- gather: survivors in input order; 12 columns (processLC.cpp:291-307) ---- */
        if (K) { /* warp-uniform */
            unsigned int before = 0;
#pragma unroll
            for (int w = 0; w < K1_TILE / 32; ++w) {
                if ((unsigned)(w * 32) < C) {
                    const unsigned int b = bits[w];
                    if ((b >> lane) & 1u) {
                        const unsigned int j = w * 32 + lane;
                        const long long r = (long long)(exclusive + before + __popc(b & lane_lt));
                        if (r < a.cap) {
                            const long long g = base + q.idx[j];
                            a.out.theta[r] = q.th[j];
                            a.out.phi[r] = q.ph[j];
                            a.out.x[r] = q.x[j];
                            a.out.y[r] = q.y[j];
                            a.out.z[r] = q.z[j];
                            a.out.vx[r] = a.in.vx[g];
                            a.out.vy[r] = a.in.vy[g];
                            a.out.vz[r] = a.in.vz[g];
                            a.out.redshift[r] = a_to_z_dev(a.in.a[g]);
                            a.out.id[r] = a.in.id[g];
                            a.out.rot[r] = a.in.rot[g];
                            a.out.rep[r] = a.in.rep[g];
                            if (a.out.index) a.out.index[r] = a.index_base + g;
                        }
                    }
                    before += __popc(b);
                }
            }
        }
        __syncwarp(); /* the queue is rewritten by the next tile */
    }
    if (lane == 0 && n_cand) atomicAdd(a.result + 1, n_cand);
}

/* ================================================================================================
 * K2  use case 2: one pass, a batch of halos
 * ================================================================================================ */
constexpr int K2_THREADS = 256;
constexpr int K2_MAX_HALOS = 1024; /* halos per launch (pre-filter parameters live in shared memory) */

struct HaloDev { /* exact parameters of one halo */
    float R[9];
    float th_lo, th_hi, ph_lo, ph_hi;     /* fine, strict (processLC.cpp:1439-1440) */
    float rth_lo, rth_hi, rph_lo, rph_hi; /* rough: lo <= theta_u < hi (:1334-1337), lo < phi_u < hi (:1400) */
    float pad[3];
};
struct Rec {                /* one survivor */
    unsigned long long key; /* (bits of theta_u) << 32 | local particle index: the reference's sort order */
    unsigned int halo;
    float th, ph;           /* rotated, halo-centric angles (processLC.cpp:1446-1447) */
    unsigned int pad;
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
    int total_halos;
};

/* exact path for one (particle, halo) pair; noinline keeps the scan loop small.  Arguments are passed by
 * value (registers) so the kernel parameter block never has to be copied to local memory. */
__device__ __noinline__ void k2_exact(const HaloDev* __restrict__ Hp, unsigned int halo_global, Rec* recs,
                                      unsigned long long rec_cap, unsigned long long* rec_count,
                                      unsigned long long* counts, int total_halos, unsigned long long* n_cand,
                                      float x, float y, float z, unsigned int local_n) {
    const float thu = ccref::theta_arcsec(x, y, z);       /* processLC.cpp:899-900 */
    const float phu = ccref::phi_arcsec_guarded(x, y);    /* :903-908 */
    atomicAdd(n_cand, 1ull);
    if (!(thu >= Hp->rth_lo && thu < Hp->rth_hi)) return; /* :1334-1340 (two lower_bounds) */
    if (!(phu > Hp->rph_lo && phu < Hp->rph_hi)) return;  /* :1400 */
    atomicAdd(counts + total_halos + halo_global, 1ull);  /* rough_cutout_size, :1436 */
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Hp->R[i];
    float rx, ry, rz;
    ccref::mat_vec(R, x, y, z, rx, ry, rz);               /* :1419-1422 */
    const float th = ccref::theta_arcsec(rx, ry, rz);     /* :1425-1426 */
    const float ph = ccref::phi_arcsec_guarded(rx, ry);   /* :1430-1435 */
    if (th > Hp->th_lo && th < Hp->th_hi && ph > Hp->ph_lo && ph < Hp->ph_hi) { /* :1439-1440 */
        atomicAdd(counts + halo_global, 1ull);
        const unsigned long long slot = atomicAdd(rec_count, 1ull);
        if (slot < rec_cap) {
            Rec r;
            r.key = ((unsigned long long)(unsigned int)__float_as_int(thu) << 32) | local_n;
            r.halo = halo_global;
            r.th = th;
            r.ph = ph;
            r.pad = 0;
            recs[slot] = r;
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(K2_THREADS, 4) k2_cutout_halo(const HaloArgs a) {
    __shared__ float4 s_pre[K2_MAX_HALOS];
    for (int h = threadIdx.x; h < a.nhalos; h += K2_THREADS) s_pre[h] = a.pre[h];
    __syncthreads();
    const long long nquads = (a.n + 3) >> 2;
    for (long long q = (long long)blockIdx.x * K2_THREADS + threadIdx.x; q < nquads;
         q += (long long)gridDim.x * K2_THREADS) {
        const long long g0 = q << 2;
        float px[4], py[4], pz[4];
        unsigned int valid = 0xfu;
        if (VEC && g0 + 4 <= a.n) {
            const float4 vx = ldg_stream4(a.x + g0), vy = ldg_stream4(a.y + g0), vz = ldg_stream4(a.z + g0);
            px[0] = vx.x; px[1] = vx.y; px[2] = vx.z; px[3] = vx.w;
            py[0] = vy.x; py[1] = vy.y; py[2] = vy.z; py[3] = vy.w;
            pz[0] = vz.x; pz[1] = vz.y; pz[2] = vz.z; pz[3] = vz.w;
        } else {
            valid = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (g0 + j < a.n) {
                    px[j] = __ldcs(a.x + g0 + j); py[j] = __ldcs(a.y + g0 + j); pz[j] = __ldcs(a.z + g0 + j);
                    valid |= 1u << j;
                } else {
                    px[j] = py[j] = pz[j] = 1.0f;
                }
            }
        }
        /* per-particle quantities shared by all halos: r = |v| (approx), t = y/x (approx), range flag */
        float pr[4], pt[4];
        unsigned int insane = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float ax = fabsf(px[j]), ay = fabsf(py[j]), az = fabsf(pz[j]);
            const float mx = fmaxf(ax, fmaxf(ay, az));
            /* x must be in range for the approximate reciprocal; r needs only the largest component */
            if (!(ax >= CC_SANE_LO && mx <= CC_SANE_HI)) insane |= 1u << j;
            const float r2 = fmaf(pz[j], pz[j], fmaf(py[j], py[j], px[j] * px[j]));
            pr[j] = r2 * rsqrtf(r2);
            pt[j] = __fdividef(py[j], px[j]);
        }
        insane &= valid;
        for (int h = 0; h < a.nhalos; ++h) {
            const float4 P = s_pre[h];
            unsigned int m = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool c = (pz[j] <= P.x * pr[j]) & (pz[j] >= P.y * pr[j]) & (pt[j] >= P.z) & (pt[j] <= P.w);
                m |= (unsigned)c << j;
            }
            m = (m & valid) | insane;
            if (m) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m & (1u << j))
                        k2_exact(a.halos + h, (unsigned int)(a.halo0 + h), a.recs, a.rec_cap, a.rec_count, a.counts,
                                 a.total_halos, a.n_cand, px[j], py[j], pz[j], (unsigned int)(a.local_base + g0 + j));
            }
        }
    }
}

/* ================================================================================================
 * K3  survivor ordering + gather (use case 2)
 * ================================================================================================ */
/* scatter records into per-halo segments (unordered inside a segment) */
struct BucketArgs {
    const Rec* recs;
    unsigned long long nrec;
    const unsigned long long* seg_begin; /* [H] exclusive scan of counts */
    unsigned long long* cursor;          /* [H] zeroed */
    Rec* out;
};
__global__ void k3_bucket(const BucketArgs a) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrec;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const Rec r = a.recs[i];
        const unsigned long long p = a.seg_begin[r.halo] + atomicAdd(a.cursor + r.halo, 1ull);
        a.out[p] = r;
    }
}

/* exclusive scan of H counts (H is small: one CTA, serial chunks) */
__global__ void k3_scan_counts(const unsigned long long* counts, unsigned long long* seg_begin, int H) {
    __shared__ unsigned long long s_part[256];
    const int tid = threadIdx.x;
    const int per = (H + 255) / 256;
    unsigned long long sum = 0;
    for (int i = tid * per; i < min(H, (tid + 1) * per); ++i) sum += counts[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 256; ++i) { const unsigned long long t = s_part[i]; s_part[i] = run; run += t; }
    }
    __syncthreads();
    unsigned long long run = s_part[tid];
    for (int i = tid * per; i < min(H, (tid + 1) * per); ++i) { seg_begin[i] = run; run += counts[i]; }
    if (tid == 255) seg_begin[H] = run;
}

/* rank of every record inside its halo segment by counting smaller keys (keys are unique: the low 32 bits
 * are the particle index).  One CTA column per segment chunk; the segment streams through shared memory. */
constexpr int K3_THREADS = 256;
struct RankArgs {
    const Rec* recs;                     /* bucketed */
    const unsigned long long* seg_begin; /* [H+1] */
    unsigned int* rank;                  /* [nrec] rank inside the segment */
    int H;
};
__global__ void __launch_bounds__(K3_THREADS) k3_rank(const RankArgs a) {
    __shared__ unsigned long long s_key[K3_THREADS];
    const int h = blockIdx.y;
    const unsigned long long b = a.seg_begin[h], e = a.seg_begin[h + 1];
    const unsigned long long len = e - b;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * K3_THREADS; i0 < len;
         i0 += (unsigned long long)gridDim.x * K3_THREADS) {
        const unsigned long long i = i0 + threadIdx.x;
        const unsigned long long mykey = (i < len) ? a.recs[b + i].key : 0ull;
        unsigned int r = 0;
        for (unsigned long long j0 = 0; j0 < len; j0 += K3_THREADS) {
            const unsigned long long j = j0 + threadIdx.x;
            __syncthreads();
            s_key[threadIdx.x] = (j < len) ? a.recs[b + j].key : ~0ull;
            __syncthreads();
            const int lim = (int)min((unsigned long long)K3_THREADS, len - j0);
            for (int k = 0; k < lim; ++k) r += (s_key[k] < mykey);
        }
        if (i < len) a.rank[b + i] = r;
    }
}

struct GatherArgs {
    const Rec* recs;
    const unsigned int* rank;
    const unsigned long long* seg_begin;
    unsigned long long nrec;
    ColsIn in;
    ColsOut out;
    long long index_base;
    int pos_only;
};
__global__ void k3_gather_halo(const GatherArgs a) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrec;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const Rec r = a.recs[i];
        const unsigned long long o = a.seg_begin[r.halo] + a.rank[i];
        const long long g = (long long)(r.key & 0xffffffffull);
        a.out.theta[o] = r.th;                        /* processLC.cpp:1446-1447 */
        a.out.phi[o] = r.ph;
        a.out.redshift[o] = a_to_z_dev(a.in.a[g]);    /* :1450 */
        a.out.x[o] = a.in.x[g];
        a.out.y[o] = a.in.y[g];
        a.out.z[o] = a.in.z[g];
        a.out.id[o] = a.in.id[g];
        if (!a.pos_only) {                            /* :1455-1461 */
            a.out.vx[o] = a.in.vx[g];
            a.out.vy[o] = a.in.vy[g];
            a.out.vz[o] = a.in.vz[g];
            a.out.rot[o] = a.in.rot[g];
            a.out.rep[o] = a.in.rep[g];
        }
        a.out.theta_u[o] = __int_as_float((int)(r.key >> 32));
        a.out.index[o] = a.index_base + g;
    }
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
            default: r = a_to_z_dev(in0[i]); break;
        }
        out[i] = r;
    }
}

}  // namespace cck
