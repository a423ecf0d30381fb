// Process-wide cache of device and pinned-host blocks.  performGlobalIlluminationCl is a one-shot,
// host-buffer call (global_illumination_cl.c:275-321 creates and destroys every OpenCL object per
// call); caching the raw CUDA allocations between calls keeps that contract for the caller (no
// visible state survives) while taking cudaMalloc / cudaMallocHost / cudaFree — each of which
// costs 0.1-1 ms and, for cudaFree, a device synchronisation — off the repeated-call path.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <vector>

namespace fmgi {

class MemPool {
public:
    static MemPool &get() { static MemPool p; return p; }

    cudaError_t alloc(void **out, size_t bytes, bool pinned)
    {
        *out = nullptr;
        int dev = -1;
        if (!pinned) {
            cudaError_t e = cudaGetDevice(&dev);
            if (e != cudaSuccess) return e;
        }
        const size_t want = round_up(bytes);
        {
            std::lock_guard<std::mutex> lock(mu_);
            size_t best = (size_t)-1;
            for (size_t i = 0; i < free_.size(); i++) {
                const Block &b = free_[i];
                if (b.pinned == pinned && b.device == dev && b.bytes >= want && b.bytes <= 2 * want + (64 << 10) &&
                    (best == (size_t)-1 || b.bytes < free_[best].bytes))
                    best = i;
            }
            if (best != (size_t)-1) {
                Block b = free_[best];
                free_.erase(free_.begin() + best);
                live_[b.ptr] = b;
                *out = b.ptr;
                return cudaSuccess;
            }
        }
        void *p = nullptr;
        cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            release();                      // give cached blocks back and retry once
            cudaGetLastError();
            e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
            if (e != cudaSuccess) return e;
        }
        std::lock_guard<std::mutex> lock(mu_);
        live_[p] = Block{p, want, dev, pinned};
        *out = p;
        return cudaSuccess;
    }

    void free(void *p)
    {
        if (!p) return;
        std::lock_guard<std::mutex> lock(mu_);
        auto it = live_.find(p);
        if (it == live_.end()) return;
        free_.push_back(it->second);
        live_.erase(it);
    }

    // Returns every cached (not live) block to the driver.
    void release()
    {
        std::vector<Block> blocks;
        {
            std::lock_guard<std::mutex> lock(mu_);
            blocks.swap(free_);
        }
        int prev = -1;
        cudaGetDevice(&prev);
        for (const Block &b : blocks) {
            if (b.pinned) cudaFreeHost(b.ptr);
            else { cudaSetDevice(b.device); cudaFree(b.ptr); }
        }
        if (prev >= 0) cudaSetDevice(prev);
    }

    // Releases cached (not live) blocks, largest first, until at most `keep_bytes` stay cached PER DEVICE (pinned
    // host blocks count as one more "device").  Called at the end of every host-buffer entry point:
    // performGlobalIlluminationCl keeps 256 MB (small tables and staging buffers stay, a 1.8 GB atlas does not),
    // fmgi_bake keeps FMGI_CACHE_MB (default 8 GB) because cudaFree of gigabyte blocks was measured at up to 0.5 s.
    void trim(size_t keep_bytes)
    {
        std::vector<Block> victims;
        {
            std::lock_guard<std::mutex> lock(mu_);
            std::map<int, size_t> cached;             // device (-1: pinned) -> cached bytes
            for (const Block &b : free_) cached[b.pinned ? -1 : b.device] += b.bytes;
            for (auto &kv : cached) {
                while (kv.second > keep_bytes) {
                    size_t big = (size_t)-1;
                    for (size_t i = 0; i < free_.size(); i++)
                        if ((free_[i].pinned ? -1 : free_[i].device) == kv.first &&
                            (big == (size_t)-1 || free_[i].bytes > free_[big].bytes))
                            big = i;
                    if (big == (size_t)-1) break;
                    kv.second -= free_[big].bytes;
                    victims.push_back(free_[big]);
                    free_.erase(free_.begin() + big);
                }
            }
        }
        if (victims.empty()) return;
        int prev = -1;
        cudaGetDevice(&prev);
        for (const Block &b : victims) {
            if (b.pinned) cudaFreeHost(b.ptr);
            else { cudaSetDevice(b.device); cudaFree(b.ptr); }
        }
        if (prev >= 0) cudaSetDevice(prev);
    }

    size_t cached_bytes()
    {
        std::lock_guard<std::mutex> lock(mu_);
        size_t cached = 0;
        for (const Block &b : free_) cached += b.bytes;
        return cached;
    }

private:
    struct Block { void *ptr; size_t bytes; int device; bool pinned; };
    static size_t round_up(size_t b)
    {
        if (b < 256) return 256;
        const size_t g = b < (1u << 20) ? 256 : (64u << 10);
        return (b + g - 1) / g * g;
    }
    std::mutex mu_;
    std::map<void *, Block> live_;
    std::vector<Block> free_;
};

}  // namespace fmgi
