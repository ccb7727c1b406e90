#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void touch(char *p, size_t n, int nt) {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=]() { size_t a = n * t / nt, b = n * (t + 1) / nt; memset(p + a, 1, b - a); });
    for (auto &x : th) x.join();
}
int main() {
    cudaFree(0);
    FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r"); char buf[128] = {0}; if (f) { fgets(buf, 127, f); fclose(f); } printf("THP: %s", buf);
    for (size_t mb : { (size_t)62, (size_t)772 }) {
        size_t n = mb << 20;
        for (int mode = 0; mode < 3; mode++) {
            char *p = nullptr; double t0 = now();
            if (mode == 0) p = (char *)calloc(n, 1);
            else { p = (char *)aligned_alloc(2 << 20, (n + (2 << 20) - 1) / (2 << 20) * (2 << 20)); if (mode == 2) madvise(p, n, MADV_HUGEPAGE); }
            double t1 = now(); touch(p, n, 8); double t2 = now();
            cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterDefault); double t3 = now();
            void *d; cudaMalloc(&d, n); cudaMemcpy(d, p, n, cudaMemcpyHostToDevice); double t4 = now(); cudaMemcpy(d, p, n, cudaMemcpyHostToDevice); double t5 = now();
            cudaHostUnregister(p); double t6 = now();
            printf("%zu MB mode %d (%s): alloc %.1f touch(8thr) %.1f register %.1f (%s) h2d %.1f/%.1f unregister %.1f\n", mb, mode, mode == 0 ? "calloc" : mode == 1 ? "aligned" : "aligned+THP", t1 - t0, t2 - t1, t3 - t2, cudaGetErrorString(e), t4 - t3, t5 - t4, t6 - t5);
            // pageable copy for comparison
            double t7 = now(); cudaMemcpy(d, p, n, cudaMemcpyHostToDevice); double t8 = now();
            printf("    pageable h2d %.1f ms\n", t8 - t7);
            cudaFree(d); free(p);
        }
        // cudaHostAlloc
        double t0 = now(); void *hp; cudaHostAlloc(&hp, n, cudaHostAllocDefault); double t1 = now(); touch((char *)hp, n, 8); double t2 = now();
        printf("%zu MB cudaHostAlloc %.1f touch %.1f\n", mb, t1 - t0, t2 - t1); cudaFreeHost(hp);
    }
    // single-thread memcpy speed into pinned
    { size_t n = 256u << 20; char *src = (char *)malloc(n); memset(src, 1, n); void *hp; cudaHostAlloc(&hp, 32u << 20, cudaHostAllocDefault);
      double t0 = now(); for (size_t o = 0; o < n; o += 32u << 20) memcpy(hp, src + o, 32u << 20); double t1 = now();
      printf("memcpy pageable->pinned 1 thread: %.1f GB/s\n", n / (t1 - t0) / 1e6); }
    return 0;
}
