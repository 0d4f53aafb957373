// mb_gather.cu -- what does a random 32-byte read cost on B200?  (profiles/r02: input to the gather design)
// A survivor's payload row is 32 B at a random (increasing) index; density ~4.5 % (wide-field config).
// Variants: L2 fetch granularity limit 32/64/128, load flavour (ld.global.nc, ld.global.cs, L2::64B hint... ).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long mix(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31);
}
__global__ void make_idx(unsigned int* idx, long long k, int stride) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (long long)gridDim.x * blockDim.x)
        idx[i] = (unsigned int)(i * stride + (long long)(mix(i + 1) % (unsigned)stride));
}
template <int MODE>
__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 v;
    if (MODE == 0) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int MODE, int BYTES>
__global__ void gather(const uint4* __restrict__ pay, const unsigned int* __restrict__ idx, unsigned int* out, long long k) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (long long)gridDim.x * blockDim.x) {
        const uint4* p = pay + 2ll * idx[i];
        uint4 a = ld16<MODE>(p);
        unsigned int s = a.x ^ a.y ^ a.z ^ a.w;
        if (BYTES == 32) { uint4 b = ld16<MODE>(p + 1); s ^= b.x ^ b.y ^ b.z ^ b.w; }
        out[i] = s;
    }
}
// streaming reference: read the same number of payload rows contiguously
__global__ void stream_rows(const uint4* __restrict__ pay, unsigned int* out, long long k) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (long long)gridDim.x * blockDim.x) {
        uint4 a = pay[2 * i], b = pay[2 * i + 1];
        out[i] = a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    }
}
int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 500000000ll;
    const int stride = argc > 2 ? atoi(argv[2]) : 22;
    const long long k = n / stride;
    uint4* pay; unsigned int *idx, *out;
    CK(cudaMalloc(&pay, (size_t)n * 32)); CK(cudaMalloc(&idx, (size_t)k * 4)); CK(cudaMalloc(&out, (size_t)k * 4));
    CK(cudaMemset(pay, 1, (size_t)n * 32));
    make_idx<<<1184, 256>>>(idx, k, stride);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, auto launch) {
        for (int w = 0; w < 2; ++w) launch();
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("%-44s %8.3f ms  %7.1f ns/row-per-us  useful %7.1f GB/s (32 B rows)\n", name, best, 0.0, (double)k * 32 / best / 1e6);
    };
    printf("payload rows n=%lld, gathered k=%lld (1 of %d), random increasing\n", n, k, stride);
    size_t lim = 0; CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity)); printf("default cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    const int grids[2] = {148 * 8, 148 * 32};
    for (size_t g : {(size_t)0, (size_t)32, (size_t)64, (size_t)128}) {
        if (g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); printf("-- set fetch granularity %zu: %s\n", g, cudaGetErrorString(e)); }
        for (int gi = 0; gi < 2; ++gi) {
            const int G = grids[gi];
            char nm[96];
            snprintf(nm, 96, "nc 32B grid %d", G); timeit(nm, [&] { gather<0, 32><<<G, 256>>>(pay, idx, out, k); });
            if (g == 0 || g == 32) {
                snprintf(nm, 96, "cs 32B grid %d", G); timeit(nm, [&] { gather<1, 32><<<G, 256>>>(pay, idx, out, k); });
                snprintf(nm, 96, "nc.no_allocate 32B grid %d", G); timeit(nm, [&] { gather<2, 32><<<G, 256>>>(pay, idx, out, k); });
                snprintf(nm, 96, "nc.L2::64B 32B grid %d", G); timeit(nm, [&] { gather<3, 32><<<G, 256>>>(pay, idx, out, k); });
                snprintf(nm, 96, "cg 32B grid %d", G); timeit(nm, [&] { gather<4, 32><<<G, 256>>>(pay, idx, out, k); });
                snprintf(nm, 96, "nc 16B-only grid %d", G); timeit(nm, [&] { gather<0, 16><<<G, 256>>>(pay, idx, out, k); });
            }
        }
    }
    timeit("stream k rows contiguous", [&] { stream_rows<<<148 * 16, 256>>>(pay, out, k); });
    timeit("stream 4k rows contiguous (128 B per gathered row)", [&] { stream_rows<<<148 * 16, 256>>>(pay, out, 4 * k > n ? n : 4 * k); });
    return 0;
}
