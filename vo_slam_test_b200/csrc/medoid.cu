// medoid.cu -- MapPoint::computeDescriptor (mappoint.cpp:118-179) for many map points at once: the representative
// descriptor of a map point is the observation whose MEDIAN Hamming distance to the other observations is smallest
// (median = sorted row element int(0.5*(N-1)), row includes the 0 self distance; strict '<' so the first row wins ties).
// One warp per map point; lanes own rows; the k-th smallest of a row is found by counting (N is a few tens at most).
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace orbx {

__device__ __forceinline__ int hamm32(const uint4 a0, const uint4 a1, const uint8_t* b) {
  const uint4* q = reinterpret_cast<const uint4*>(b);
  const uint4 b0 = __ldg(q), b1 = __ldg(q + 1);
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256) medoid_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ start, int P,
                                                     int32_t* __restrict__ best) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const int s = __ldg(start + p), N = __ldg(start + p + 1) - s;
  if (N <= 0) { if (lane == 0) best[p] = -1; return; }
  const int k = (int)(0.5 * (N - 1));                 // mappoint.cpp:165
  const uint8_t* D = desc + (size_t)s * 32;
  uint32_t bestKey = 0xFFFFFFFFu;                     // mid << 16 | row
  for (int i = lane; i < N; i += 32) {
    const uint4* q = reinterpret_cast<const uint4*>(D + (size_t)i * 32);
    const uint4 a0 = __ldg(q), a1 = __ldg(q + 1);
    // k-th smallest of row i: smallest v with #{j : d_ij <= v} > k, found by walking candidates v = d_ij
    int mid = 256;
    for (int j = 0; j < N; ++j) {
      const int v = (j == i) ? 0 : hamm32(a0, a1, D + (size_t)j * 32);
      if (v >= mid) continue;
      int le = 0;
      for (int t = 0; t < N; ++t) le += ((t == i ? 0 : hamm32(a0, a1, D + (size_t)t * 32)) <= v) ? 1 : 0;
      if (le > k) mid = v;
    }
    const uint32_t key = ((uint32_t)mid << 16) | (uint32_t)i;
    if (mid < 256) bestKey = min(bestKey, key);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bestKey = min(bestKey, __shfl_xor_sync(0xffffffffu, bestKey, o));
  if (lane == 0) best[p] = bestKey == 0xFFFFFFFFu ? 0 : (int)(bestKey & 0xFFFFu);    // bestIdx starts at 0 (:159)
}

}  // namespace orbx

using namespace orbx;

// desc: all observations back to back (32 B each); start[P+1]: CSR of the observations of each map point;
// best[p] = index (within the point's observations) of the descriptor the reference would keep, -1 for a point without
// observations (the reference returns early and keeps the old descriptor).
extern "C" int orbx_medoid_descriptors(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best, int device) {
  if (!start || !best || npoints < 0 || (npoints > 0 && start[npoints] > 0 && !desc)) { set_error("bad argument"); return ORBX_ERR_ARG; }
  if (npoints == 0) return ORBX_OK;
  for (int p = 0; p < npoints; ++p)
    if (start[p + 1] - start[p] > 65535) { set_error("more than 65535 observations for one map point"); return ORBX_ERR_ARG; }
  ORBX_CUDA(cudaSetDevice(device));
  const size_t total = (size_t)start[npoints];
  uint8_t* d_desc = nullptr; int32_t* d_start = nullptr; int32_t* d_best = nullptr;
  auto cleanup = [&]() { cudaFree(d_desc); cudaFree(d_start); cudaFree(d_best); };
#define MK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error(cudaGetErrorString(e_)); cleanup(); return ORBX_ERR_CUDA; } } while (0)
  MK(cudaMalloc(&d_desc, std::max<size_t>(total * 32, 32)));
  MK(cudaMalloc(&d_start, sizeof(int32_t) * (npoints + 1)));
  MK(cudaMalloc(&d_best, sizeof(int32_t) * npoints));
  if (total) MK(cudaMemcpy(d_desc, desc, total * 32, cudaMemcpyHostToDevice));
  MK(cudaMemcpy(d_start, start, sizeof(int32_t) * (npoints + 1), cudaMemcpyHostToDevice));
  medoid_kernel<<<(npoints + 7) / 8, 256>>>(d_desc, d_start, npoints, d_best);
  MK(cudaGetLastError());
  MK(cudaMemcpy(best, d_best, sizeof(int32_t) * npoints, cudaMemcpyDeviceToHost));
#undef MK
  cleanup();
  return ORBX_OK;
}
