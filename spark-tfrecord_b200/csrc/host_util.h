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

// device blocks reused across batches (all use is ordered on the owning decoder's stream)
struct DevPool {
  struct Block { void* p; size_t cap; bool used; };
  std::vector<Block> blocks;
  void* acquire(size_t bytes) {
    int best = -1;
    for (size_t i = 0; i < blocks.size(); ++i)
      if (!blocks[i].used && blocks[i].cap >= bytes && (best < 0 || blocks[i].cap < blocks[best].cap)) best = (int)i;
    if (best >= 0) { blocks[best].used = true; return blocks[best].p; }
    for (size_t i = 0; i < blocks.size();) {
      if (!blocks[i].used && blocks.size() > 6) { cudaFree(blocks[i].p); blocks.erase(blocks.begin() + i); } else ++i;
    }
    void* p = nullptr;
    size_t cap = bytes + bytes / 8 + 4096;
    if (cudaMalloc(&p, cap) != cudaSuccess) { cap = bytes; if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr; }
    blocks.push_back({p, cap, true});
    return p;
  }
  void give_back(void* p) { for (auto& b : blocks) if (b.p == p) b.used = false; }
  void release_all() { for (auto& b : blocks) cudaFree(b.p); blocks.clear(); }
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
