// scratch.h -- per-thread staging for the host-pointer entry points of the matchers.
//
// A matcher call at SLAM-frame size moves a few hundred KB; a cudaMalloc/cudaFree/cudaMemcpy per
// argument array would cost milliseconds.  Every host thread instead keeps ONE grow-only device
// slab with a pinned host mirror and a stream: a call lays its arrays out in the slab, packs the
// small inputs into the mirror and uploads them with one copy, runs its kernels on the stream and
// brings all outputs back with one copy.  Entry points are synchronous, so the slab is free again
// when they return; threads never share a slab (ORBmatcher is called from the tracking, local
// mapping and loop closing threads concurrently).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <vector>

struct OrbfeArena {
    int device = -1;
    char* d = nullptr;
    char* h = nullptr;        // pinned mirror of the first hcap bytes (small calls are packed through it)
    size_t cap = 0, hcap = 0;
    cudaStream_t st = nullptr;
    static constexpr size_t kMirrorMax = ((size_t)32 << 20) + ((size_t)1 << 20);

    cudaError_t reserve(int dev, size_t bytes) {
        cudaError_t e = cudaSetDevice(dev);
        if (e != cudaSuccess) return e;
        if (dev != device) {
            release();
            device = dev;
            e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
        }
        if (bytes > cap) {
            if (d) cudaFree(d);
            if (h) cudaFreeHost(h);
            d = nullptr; h = nullptr; cap = 0; hcap = 0;
            const size_t want = bytes + bytes / 4 + (1 << 20);
            e = cudaMalloc(&d, want);
            if (e != cudaSuccess) return e;
            const size_t hw = want < kMirrorMax ? want : kMirrorMax;
            e = cudaMallocHost(&h, hw);
            if (e != cudaSuccess) return e;
            cap = want;
            hcap = hw;
        }
        return cudaSuccess;
    }
    void release() {
        if (device >= 0) cudaSetDevice(device);
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        if (st) cudaStreamDestroy(st);
        d = nullptr; h = nullptr; cap = 0; hcap = 0; st = nullptr; device = -1;
    }
    ~OrbfeArena() {}   // freed with the context at process exit (thread exit order vs. CUDA teardown is unsafe)
};

inline OrbfeArena& orbfe_arena() {
    static thread_local OrbfeArena a;
    return a;
}

// Layout of one call inside the arena: inputs first (uploaded together), then work arrays, then
// outputs (downloaded together).
class OrbfeStage {
   public:
    size_t in(const void* src, size_t bytes) { return add(src, nullptr, bytes, 0); }
    size_t work(size_t bytes) { return add(nullptr, nullptr, bytes, 1); }
    size_t out(void* dst, size_t bytes) { return add(nullptr, dst, bytes, 2); }
    // in/out: uploaded from `io`, downloaded to `io`
    size_t inout(void* io, size_t bytes) { return add(io, io, bytes, 3); }

    cudaError_t commit(int device) {
        // order: kind 0 (in), 3 (inout), 1 (work), 2 (out); inout sits between so that both the upload
        // range [in..inout] and the download range [inout..out] -- minus the work arrays -- are contiguous
        size_t off = 0;
        const int order[4] = {0, 3, 1, 2};
        for (int k = 0; k < 4; k++) {
            if (order[k] == 3) ioBegin_ = off;
            if (order[k] == 1) upEnd_ = off;
            if (order[k] == 2) outBegin_ = off;
            for (auto& it : items_)
                if (it.kind == order[k]) { it.off = off; off += (it.bytes + 255) & ~(size_t)255; }
        }
        total_ = off;
        a_ = &orbfe_arena();
        return a_->reserve(device, total_ ? total_ : 256);
    }
    template <class T> T* ptr(size_t id) const { return reinterpret_cast<T*>(a_->d + items_[id].off); }
    cudaStream_t stream() const { return a_->st; }

    cudaError_t upload() {
        if (upEnd_ == 0) return cudaSuccess;
        if (upEnd_ <= a_->hcap) {
            for (auto& it : items_)
                if ((it.kind == 0 || it.kind == 3) && it.src && it.bytes) memcpy(a_->h + it.off, it.src, it.bytes);
            return cudaMemcpyAsync(a_->d, a_->h, upEnd_, cudaMemcpyHostToDevice, a_->st);
        }
        for (auto& it : items_)
            if ((it.kind == 0 || it.kind == 3) && it.src && it.bytes) {
                cudaError_t e = cudaMemcpyAsync(a_->d + it.off, it.src, it.bytes, cudaMemcpyHostToDevice, a_->st);
                if (e != cudaSuccess) return e;
            }
        return cudaSuccess;
    }
    // Copies inout + out arrays back and synchronises the stream.
    cudaError_t download() {
        const bool small = total_ <= a_->hcap;
        cudaError_t e = cudaSuccess;
        if (small) {
            // two ranges: [ioBegin_, upEnd_) and [outBegin_, total_)
            if (upEnd_ > ioBegin_) e = cudaMemcpyAsync(a_->h + ioBegin_, a_->d + ioBegin_, upEnd_ - ioBegin_, cudaMemcpyDeviceToHost, a_->st);
            if (e == cudaSuccess && total_ > outBegin_)
                e = cudaMemcpyAsync(a_->h + outBegin_, a_->d + outBegin_, total_ - outBegin_, cudaMemcpyDeviceToHost, a_->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(a_->st);
            if (e != cudaSuccess) return e;
            for (auto& it : items_)
                if ((it.kind == 2 || it.kind == 3) && it.dst && it.bytes) memcpy(it.dst, a_->h + it.off, it.bytes);
            return cudaSuccess;
        }
        for (auto& it : items_)
            if ((it.kind == 2 || it.kind == 3) && it.dst && it.bytes) {
                e = cudaMemcpyAsync(it.dst, a_->d + it.off, it.bytes, cudaMemcpyDeviceToHost, a_->st);
                if (e != cudaSuccess) return e;
            }
        return cudaStreamSynchronize(a_->st);
    }

   private:
    struct Item { const void* src; void* dst; size_t bytes; int kind; size_t off; };
    size_t add(const void* src, void* dst, size_t bytes, int kind) {
        items_.push_back({src, dst, bytes, kind, 0});
        return items_.size() - 1;
    }
    std::vector<Item> items_;
    OrbfeArena* a_ = nullptr;
    size_t total_ = 0, upEnd_ = 0, ioBegin_ = 0, outBegin_ = 0;
};
