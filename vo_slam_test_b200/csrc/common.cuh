// common.cuh -- shared helpers for the sm_100a kernels of libvoslam_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <string>

#include "../../include/orb_b200.h"

namespace orbx {

void set_error(const std::string& msg);

#define ORBX_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      orbx::set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                           \
      return ORBX_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

// Programmatic dependent launch (PDL): every kernel of the per-chunk chain starts with pdl_prologue().  launch_dependents
// lets the NEXT kernel's CTAs become resident while this grid's last wave drains; wait blocks this CTA until the PREVIOUS
// grid has completed and its writes are visible.  Nothing touches global memory before the wait, so the only thing that
// overlaps is launch latency + CTA start-up.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

int pdl_enabled();   // orb_capi.cu: ORBX_PDL = 0 (off), 1 (default: the pyramid chain only, where 7 short launches follow each other;
                     // measured +2 % there, -2 % when the long kernels are chained too), 2 (every kernel of the chunk chain)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
static inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- host-side scratch shared by the host entry points (sbp.cu, hamming.cu, frame.cu, bow.cu) ---------------------
struct DevArena {      // grow-only per-thread device scratch so repeated searches do not pay cudaMalloc
  uint8_t* base = nullptr; size_t cap = 0, used = 0; int device = -1;
  int reserve(size_t bytes, int dev) {
    if (dev != device || bytes > cap) {
      if (base) { cudaSetDevice(device < 0 ? dev : device); cudaFree(base); base = nullptr; cap = 0; }
      if (cudaSetDevice(dev) != cudaSuccess) return ORBX_ERR_CUDA;
      size_t want = std::max(bytes + bytes / 2, (size_t)8 << 20);
      if (cudaMalloc(&base, want) != cudaSuccess) { base = nullptr; return ORBX_ERR_CUDA; }
      cap = want; device = dev;
    } else if (cudaSetDevice(dev) != cudaSuccess) {
      return ORBX_ERR_CUDA;
    }
    used = 0;
    return ORBX_OK;
  }
  template <typename T> T* take(size_t count) {
    used = align_up_sz(used, 256);
    T* p = reinterpret_cast<T*>(base + used);
    used += sizeof(T) * count;
    return p;
  }
};
struct HostArena {     // pinned staging so that all inputs of one search go up in a single copy
  uint8_t* base = nullptr; size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes > cap) {
      if (base) cudaFreeHost(base);
      base = nullptr; cap = 0;
      size_t want = std::max(bytes + bytes / 2, (size_t)4 << 20);
      if (cudaMallocHost(&base, want) != cudaSuccess) { base = nullptr; return ORBX_ERR_CUDA; }
      cap = want;
    }
    return ORBX_OK;
  }
};

// Candidate / keypoint packing used between the stages: x:12 | y:12 | score:8 (region coordinates).
__host__ __device__ static inline uint32_t pack_key(int x, int y, int s) {
  return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24);
}
__host__ __device__ static inline int key_x(uint32_t k) { return (int)(k & 0xFFFu); }
__host__ __device__ static inline int key_y(uint32_t k) { return (int)((k >> 12) & 0xFFFu); }
__host__ __device__ static inline int key_s(uint32_t k) { return (int)(k >> 24); }

// Exclusive scan of `n` ints held in shared memory, in place, by the whole block (blockDim.x threads,
// a multiple of 32, at most 1024).  Returns the total.  `warp_sums` is a 33-int shared scratch.  Contains barriers.
__device__ static inline int block_exclusive_scan(int* data, int n, int* warp_sums) {
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (n + T - 1) / T;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += data[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < (T >> 5)) ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    warp_sums[lane] = wi - w;            // exclusive warp offsets
    if (lane == 31) warp_sums[32] = wi;  // grand total
  }
  __syncthreads();
  int run = warp_sums[wid] + incl - sum;
  for (int i = lo; i < hi; ++i) {
    int v = data[i];
    data[i] = run;
    run += v;
  }
  int total = warp_sums[32];
  __syncthreads();
  return total;
}

}  // namespace orbx
