// staged_upload.h -- uploads of a packed genome plane that lives in ordinary pageable host memory (internal)
#pragma once
#include "kgma_internal.h"
#include <atomic>
#include <thread>
#include <vector>
#include <cstring>
#include <sys/mman.h>

namespace kgma {

// Upload of a genome whose packed plane is ordinary pageable memory.  Page-locking 772 MB costs ~100 ms (the driver
// pins page by page) and a plain cudaMemcpy from pageable memory runs at ~10 GB/s through the driver's own single-threaded
// staging; here a few host threads copy 1 MB pieces into a small page-locked ring and the calling thread queues one
// asynchronous copy per piece as it becomes ready, which keeps the link busy while the prefilter chases the data as usual.
// The plane itself is only page-locked on request (kgma_genome_make_resident, or KGMA_NO_STAGING=1).
struct StagedUpload {
    static constexpr int NS = 16;                          // ring slots
    static constexpr size_t SB = (size_t)1 << 20;          // bytes per slot (16 MB ring: cheap to page-lock on first use)
    kgma_ctx *ctx = nullptr;
    const char *src = nullptr; size_t total = 0, nsub = 0;
    std::atomic<size_t> next{0}, issued{0};
    std::vector<std::atomic<int>> filled;
    std::atomic<bool> abort{false};
    std::vector<std::thread> th;

    static int prepare_ring(kgma_ctx *ctx)
    {
        if (ctx->stage) return KGMA_OK;
        const size_t bytes = NS * SB;
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) return set_err(ctx, KGMA_E_CAPACITY, "staging ring allocation failed");
        madvise(p, bytes, MADV_HUGEPAGE);
        memset(p, 0, bytes);
        if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
            cudaGetLastError(); munmap(p, bytes);
            return set_err(ctx, KGMA_E_CUDA, "cudaHostRegister of the staging ring failed");
        }
        // the ring is published only once every event exists: a failed cudaEventCreate must not leave it marked ready
        for (auto &e : ctx->stage_ev) {
            if (e) continue;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                e = nullptr; cudaGetLastError(); cudaHostUnregister(p); munmap(p, bytes);
                return set_err(ctx, KGMA_E_CUDA, "cudaEventCreate for the staging ring failed");
            }
        }
        ctx->stage = p; ctx->stage_bytes = bytes;
        return KGMA_OK;
    }
    void start(kgma_ctx *c, const char *s, size_t bytes)
    {
        ctx = c; src = s; total = bytes; nsub = (bytes + SB - 1) / SB;
        filled = std::vector<std::atomic<int>>(nsub);
        for (auto &f : filled) f.store(0, std::memory_order_relaxed);
        const int nt = (int)std::min<size_t>(nsub, std::min(6u, std::max(1u, std::thread::hardware_concurrency() / 2)));   // (NS - 2 at most: slots must free up)
        for (int t = 0; t < nt; t++) th.emplace_back([this]() {
            cudaSetDevice(ctx->device);
            for (;;) {
                const size_t j = next.fetch_add(1);
                if (j >= nsub) return;
                if (j >= (size_t)NS) {                     // the slot's previous piece must have left for the device
                    while (issued.load(std::memory_order_acquire) < j - NS + 1) { if (abort.load()) return; std::this_thread::yield(); }
                    cudaEventSynchronize(ctx->stage_ev[j % NS]);
                }
                if (abort.load()) return;
                memcpy((char *)ctx->stage + (j % NS) * SB, src + j * SB, std::min(SB, total - j * SB));
                filled[j].store(1, std::memory_order_release);
            }
        });
    }
    // queue pieces [j0, j1) on the copy stream, in order, as the workers deliver them
    int issue(size_t j0, size_t j1, char *dst, cudaStream_t sp)
    {
        for (size_t j = j0; j < j1; j++) {
            while (!filled[j].load(std::memory_order_acquire)) std::this_thread::yield();
            KGMA_CUDA(ctx, cudaMemcpyAsync(dst + j * SB, (char *)ctx->stage + (j % NS) * SB, std::min(SB, total - j * SB), cudaMemcpyHostToDevice, sp));
            KGMA_CUDA(ctx, cudaEventRecord(ctx->stage_ev[j % NS], sp));
            issued.store(j + 1, std::memory_order_release);
        }
        return KGMA_OK;
    }
    ~StagedUpload() { abort.store(true); for (auto &t : th) t.join(); }
};


}  // namespace kgma
