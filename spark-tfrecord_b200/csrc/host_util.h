// host_util.h -- small host-side helpers for api.cu: growable device buffers, a pinned host block
// pool, and the Arrow C Data / Device Interface structs (restated from the Arrow ABI specification,
// identical in layout to arrow/c/abi.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <memory>
#include <vector>

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  // grows (never shrinks); contents are NOT preserved
  cudaError_t ensure_raw(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  int32_t ensure(size_t bytes);
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinnedPool {
  struct Block { void* p; size_t cap; bool used; };
  std::vector<Block> blocks;
  void* acquire(size_t bytes) {
    int best = -1;
    for (size_t i = 0; i < blocks.size(); ++i)
      if (!blocks[i].used && blocks[i].cap >= bytes && (best < 0 || blocks[i].cap < blocks[best].cap)) best = (int)i;
    if (best >= 0) { blocks[best].used = true; return blocks[best].p; }
    // drop free blocks that are too small so the pool does not grow without bound
    for (size_t i = 0; i < blocks.size();) {
      if (!blocks[i].used) { cudaFreeHost(blocks[i].p); blocks.erase(blocks.begin() + i); } else ++i;
    }
    void* p = nullptr;
    size_t cap = bytes + bytes / 8 + 4096;
    if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    blocks.push_back({p, cap, true});
    return p;
  }
  void give_back(void* p) { for (auto& b : blocks) if (b.p == p) b.used = false; }
  void release_all() { for (auto& b : blocks) cudaFreeHost(b.p); blocks.clear(); }
};

// device blocks reused across batches.  A block goes back to the pool when its batch is released, possibly while kernels that
// write it are still queued on the decode stream: `ready` is recorded there at that moment, and whoever takes the block for
// work on ANOTHER stream waits for it.  Free blocks are handed out oldest-release-first, so that in a steady pipeline the
// block a new batch gets was released two or three batches ago and its event has long fired.
struct DevPool {
  struct Block { void* p; size_t cap; bool used; cudaEvent_t ready; unsigned long long stamp; };
  std::vector<Block> blocks;
  unsigned long long clock = 0;
  void* acquire(size_t bytes, cudaEvent_t* ready_out = nullptr) {
    int best = -1;
    for (size_t i = 0; i < blocks.size(); ++i) {
      const Block& b = blocks[i];
      if (b.used || b.cap < bytes) continue;
      if (best < 0) { best = (int)i; continue; }
      const Block& c = blocks[best];
      // a block up to 1/8 larger than the smallest fit counts as the same size class: among those, the oldest release wins
      const bool same_class = b.cap <= c.cap + c.cap / 8 && c.cap <= b.cap + b.cap / 8;
      if (same_class ? b.stamp < c.stamp : b.cap < c.cap) best = (int)i;
    }
    if (best >= 0) { blocks[best].used = true; if (ready_out) *ready_out = blocks[best].stamp ? blocks[best].ready : nullptr; return blocks[best].p; }
    for (size_t i = 0; i < blocks.size();) {
      if (!blocks[i].used && blocks.size() > 8) { cudaFree(blocks[i].p); if (blocks[i].ready) cudaEventDestroy(blocks[i].ready); blocks.erase(blocks.begin() + i); } else ++i;
    }
    void* p = nullptr;
    size_t cap = bytes + bytes / 8 + 4096;
    if (cudaMalloc(&p, cap) != cudaSuccess) { cap = bytes; if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr; }
    blocks.push_back({p, cap, true, nullptr, 0});
    if (ready_out) *ready_out = nullptr;
    return p;
  }
  // `st`: the stream whose queued work may still touch the block
  void give_back(void* p, cudaStream_t st = nullptr) {
    for (auto& b : blocks)
      if (b.p == p) {
        b.used = false;
        b.stamp = ++clock;
        if (!b.ready) cudaEventCreateWithFlags(&b.ready, cudaEventDisableTiming);
        if (b.ready) cudaEventRecord(b.ready, st);
      }
  }
  void release_all() { for (auto& b : blocks) { cudaFree(b.p); if (b.ready) cudaEventDestroy(b.ready); } blocks.clear(); }
};

extern "C" {
struct ArrowSchema {
  const char* format; const char* name; const char* metadata; int64_t flags; int64_t n_children;
  struct ArrowSchema** children; struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*); void* private_data;
};
struct ArrowArray {
  int64_t length; int64_t null_count; int64_t offset; int64_t n_buffers; int64_t n_children;
  const void** buffers; struct ArrowArray** children; struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*); void* private_data;
};
struct ArrowDeviceArray {
  struct ArrowArray array; int64_t device_id; int32_t device_type; void* sync_event; int64_t reserved[3];
};
}
