// mb_h2d.cu -- host->device copy bandwidth with 1, 2, 4, 8 GPUs copying at once (profiles/r02: why the end-to-end
// number stops scaling past 2 GPUs).  One host thread per GPU, each copies its own pinned buffer to its own GPU.
// Variants of the host allocation: cudaHostAlloc (default flags), write-combined, malloc + first touch on a chosen
// NUMA node's cores + cudaHostRegister.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static std::vector<int> cpus_of_node(int node) {
    std::vector<int> out;
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return out;
    char buf[4096];
    if (fgets(buf, sizeof(buf), f)) {
        char* p = buf;
        while (*p) {
            int a = strtol(p, &p, 10), b = a;
            if (*p == '-') b = strtol(p + 1, &p, 10);
            for (int c = a; c <= b; ++c) out.push_back(c);
            if (*p == ',') ++p; else break;
        }
    }
    fclose(f);
    return out;
}
static int numa_nodes() { int n = 0; while (!cpus_of_node(n).empty()) ++n; return n; }

struct Barrier {
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) { count = 0; gen.fetch_add(1); }
        else while (gen.load() == g) std::this_thread::yield();
    }
};

int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? atoll(argv[1]) : 1200) * 1000000ull;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const int nodes = numa_nodes();
    printf("GPUs %d, NUMA nodes %d, online cpus %ld, buffer %zu MB per GPU\n", ndev, nodes, sysconf(_SC_NPROCESSORS_ONLN), bytes / 1000000);
    for (int d = 0; d < ndev; ++d) {
        char bus[32] = "";
        cudaDeviceGetPCIBusId(bus, sizeof(bus), d);
        std::string p = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
        for (auto& c : p) c = tolower(c);
        int node = -9;
        if (FILE* f = fopen(p.c_str(), "r")) { if (fscanf(f, "%d", &node) != 1) node = -9; fclose(f); }
        printf("  gpu %d pci %s numa_node %d\n", d, bus, node);
    }
    const char* modes[] = {"cudaHostAlloc", "write-combined", "malloc+touch(node0)+register", "malloc+touch(last node)+register"};
    for (int mode = 0; mode < 4; ++mode) {
        if (mode == 3 && nodes < 2) continue;
        for (int k = 1; k <= ndev; k *= 2) {
            std::vector<double> gbs(k, 0.0);
            Barrier bar(k);
            std::vector<std::thread> th;
            std::vector<double> t_all(k);
            for (int g = 0; g < k; ++g)
                th.emplace_back([&, g] {
                    CK(cudaSetDevice(g));
                    void *h = nullptr, *d = nullptr;
                    bool registered = false;
                    if (mode == 0) CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
                    else if (mode == 1) CK(cudaHostAlloc(&h, bytes, cudaHostAllocWriteCombined));
                    else {
                        const std::vector<int> cpus = cpus_of_node(mode == 2 ? 0 : nodes - 1);
                        if (!cpus.empty()) {
                            cpu_set_t set; CPU_ZERO(&set);
                            for (int c : cpus) CPU_SET(c, &set);
                            pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
                        }
                        if (posix_memalign(&h, 4096, bytes)) exit(2);
                        memset(h, 1, bytes); /* first touch on the chosen node */
                        CK(cudaHostRegister(h, bytes, cudaHostRegisterDefault));
                        registered = true;
                    }
                    if (mode < 2) memset(h, 1, bytes);
                    CK(cudaMalloc(&d, bytes));
                    cudaStream_t s; CK(cudaStreamCreate(&s));
                    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
                    CK(cudaStreamSynchronize(s));
                    bar.wait();
                    const auto t0 = std::chrono::steady_clock::now();
                    const int reps = 4;
                    for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
                    CK(cudaStreamSynchronize(s));
                    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    gbs[g] = reps * (double)bytes / dt / 1e9;
                    bar.wait();
                    CK(cudaFree(d));
                    if (registered) { CK(cudaHostUnregister(h)); free(h); } else CK(cudaFreeHost(h));
                    CK(cudaStreamDestroy(s));
                });
            for (auto& t : th) t.join();
            double sum = 0, mn = 1e9;
            for (double v : gbs) { sum += v; if (v < mn) mn = v; }
            printf("%-34s %d GPU(s) at once: aggregate %7.1f GB/s, slowest GPU %6.1f GB/s, per GPU:", modes[mode], k, sum, mn);
            for (double v : gbs) printf(" %.1f", v);
            printf("\n");
            fflush(stdout);
        }
    }
    return 0;
}
