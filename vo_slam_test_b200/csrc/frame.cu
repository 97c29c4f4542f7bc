// frame.cu -- what Frame::Frame does between ORB extraction and matching (frame.cpp:22-32), for a batch of frames that
// are already resident on the device (removes the D2H -> H2D round trip between the extractor
// and the grid-window searches):
//   undistortKeyPoints   frame.cpp:36-70   cv::undistortPoints(pts, pts, K, D, noArray(), K): 5 fixed-point iterations
//                                          in double, no FMA contraction (pinned bit-exact to cv2 4.13.0)
//   findDepth            frame.cpp:108-133 depth_[i] = D(v,u) at the ORIGINAL keypoint, uRight_[i] = unKp.x - bf/d
//   assignFeaturesToGrid frame.cpp:72-97   64x48 grid over the undistorted points, ids ascending inside a cell
// One CTA per frame.  The per-keypoint arithmetic is embarrassingly parallel; the ordered placement into the CSR is
// done by one warp that walks the keypoints 32 at a time and ranks equal cells with match_any (32 steps per 1000
// keypoints) instead of an O(n^2) rank.
#include <string.h>

#include <algorithm>

#include "frame_handle.cuh"

namespace orbx {

constexpr int FGC = ORBX_GRID_COLS, FGR = ORBX_GRID_ROWS, FNCELL = FGC * FGR;
constexpr int kFrameThreads = 1024;    // one CTA per frame: 1000 keypoints = one undistortion per thread (latency of the tracked frame)

struct CamDev {
  double fx, fy, cx, cy, ifx, ify;
  double k[8];
  int undistort;            // distCoef.at<float>(0) != 0 (frame.cpp:41)
  float bf;
  float xmin, ymin, gw, gh;
};

// cvUndistortPointsInternal with criteria (MAX_ITER, 5) and R = I, P = K.  Every operation is an explicit IEEE
// round-to-nearest double op (the file is compiled with -fmad=false as well).
__device__ __forceinline__ void undistort_point(const CamDev& C, float uin, float vin, float& uo, float& vo) {
  const double u = (double)uin, v = (double)vin;
  double x = __dmul_rn(__dsub_rn(u, C.cx), C.ifx), y = __dmul_rn(__dsub_rn(v, C.cy), C.ify);
  const double x0 = x, y0 = y;
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y);
    const double r2 = __dadd_rn(xx, yy);
    const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(C.k[7], r2), C.k[6]), r2), C.k[5]), r2));
    const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(C.k[4], r2), C.k[1]), r2), C.k[0]), r2));
    const double icdist = __ddiv_rn(num, den);
    if (icdist < 0) {
      x = __dmul_rn(__dsub_rn(u, C.cx), C.ifx);
      y = __dmul_rn(__dsub_rn(v, C.cy), C.ify);
      break;
    }
    // 2*k[2]*x*y + k[3]*(r2 + 2*x*x)   and   k[2]*(r2 + 2*y*y) + 2*k[3]*x*y   (left-to-right products)
    const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, C.k[2]), x), y),
                                __dmul_rn(C.k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
    const double dY = __dadd_rn(__dmul_rn(C.k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                __dmul_rn(__dmul_rn(__dmul_rn(2.0, C.k[3]), x), y));
    x = __dmul_rn(__dsub_rn(x0, dX), icdist);
    y = __dmul_rn(__dsub_rn(y0, dY), icdist);
  }
  uo = (float)__dadd_rn(__dmul_rn(C.fx, x), C.cx);
  vo = (float)__dadd_rn(__dmul_rn(C.fy, y), C.cy);
}

__global__ void __launch_bounds__(kFrameThreads) frame_finish_kernel(
    const CamDev C, const orbx_keypoint* __restrict__ kps, const int32_t* __restrict__ counts, int cap,
    const float* __restrict__ depth, int W, int H, size_t depthRow, size_t depthFrame, orbx_keypoint* __restrict__ unkps,
    float* __restrict__ uright, float* __restrict__ depthOut, int32_t* __restrict__ cellStart, int32_t* __restrict__ ids,
    float4* __restrict__ feat, float* __restrict__ angle) {
  extern __shared__ int16_t cellOf[];     // [cap]
  __shared__ int cnt[FNCELL];
  __shared__ int ws[40];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = min(max(counts[f], 0), cap);
  kps += (size_t)f * cap; unkps += (size_t)f * cap; uright += (size_t)f * cap; depthOut += (size_t)f * cap;
  ids += (size_t)f * cap; cellStart += (size_t)f * (FNCELL + 1);
  if (feat) feat += (size_t)f * cap;
  if (angle) angle += (size_t)f * cap;
  const char* dimg = depth ? (const char*)depth + (size_t)f * depthFrame : nullptr;
  for (int c = tid; c < FNCELL; c += kFrameThreads) cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kFrameThreads) {
    orbx_keypoint k = kps[i];
    const float u0 = k.x, v0 = k.y;
    if (C.undistort) undistort_point(C, u0, v0, k.x, k.y);
    unkps[i] = k;
    float d = -1.f, ur = -1.f;                                        // frame.cpp:113-114
    if (dimg) {
      const int col = min(max((int)u0, 0), W - 1), row = min(max((int)v0, 0), H - 1);   // at<float>(v,u): truncation
      const float dv = *(const float*)(dimg + (size_t)row * depthRow + (size_t)col * sizeof(float));
      if (dv > 0) { d = dv; ur = __fsub_rn(k.x, __fdiv_rn(C.bf, dv)); }                 // :126-130
    }
    uright[i] = ur; depthOut[i] = d;
    if (feat) feat[i] = make_float4(k.x, k.y, __int_as_float(k.octave), ur);    // the searches' per-feature record (sbp.cu FrameDev::feat)
    if (angle) angle[i] = k.angle;
    const int gx = (int)roundf(__fmul_rn(__fsub_rn(k.x, C.xmin), C.gw));                // frame.cpp:83-84
    const int gy = (int)roundf(__fmul_rn(__fsub_rn(k.y, C.ymin), C.gh));
    int c = -1;
    if (gx >= 0 && gx < FGC && gy >= 0 && gy < FGR) { c = gx * FGR + gy; atomicAdd(&cnt[c], 1); }   // :91-97
    cellOf[i] = (int16_t)c;
  }
  __syncthreads();
  const int total = block_exclusive_scan(cnt, FNCELL, ws);
  for (int c = tid; c < FNCELL; c += kFrameThreads) cellStart[c] = cnt[c];
  if (tid == 0) cellStart[FNCELL] = total;
  __syncthreads();
  if (tid < 32) {     // push_back order of the reference loop: ascending keypoint index inside every cell
    const unsigned lt = (1u << tid) - 1u;
    for (int base = 0; base < n; base += 32) {
      const int i = base + tid;
      const int c = i < n ? (int)cellOf[i] : -1;
      const unsigned m = __match_any_sync(0xffffffffu, c);
      int pos = 0;
      if (c >= 0) { pos = cnt[c] + __popc(m & lt); ids[pos] = i; }
      __syncwarp();
      if (c >= 0 && (m >> tid) == 1u) cnt[c] = pos + 1;               // highest lane of the group advances the cursor
      __syncwarp();
    }
  }
}

static int make_cam(const orbx_camera* cam, CamDev& C) {
  if (!cam || cam->ndist < 0 || cam->ndist > 8) { set_error("bad camera (ndist must be 0..8)"); return ORBX_ERR_ARG; }
  if (!(cam->xmax > cam->xmin) || !(cam->ymax > cam->ymin)) { set_error("bad image bounds"); return ORBX_ERR_ARG; }
  memset(&C, 0, sizeof(C));
  C.fx = cam->fx; C.fy = cam->fy; C.cx = cam->cx; C.cy = cam->cy;
  C.ifx = 1.0 / C.fx; C.ify = 1.0 / C.fy;
  for (int i = 0; i < cam->ndist; ++i) C.k[i] = (double)cam->dist[i];
  C.undistort = (cam->ndist > 0 && cam->dist[0] != 0.0f) ? 1 : 0;
  C.bf = cam->bf;
  C.xmin = cam->xmin; C.ymin = cam->ymin;
  C.gw = (float)FGC / (cam->xmax - cam->xmin);                         // camera.cpp:47-48
  C.gh = (float)FGR / (cam->ymax - cam->ymin);
  return ORBX_OK;
}

static int launch_frame_finish(const CamDev& C, const orbx_keypoint* d_kps, const int32_t* d_counts, int nframes, int cap,
                               const float* d_depth, int w, int h, size_t depthRow, size_t depthFrame, orbx_keypoint* d_unkps,
                               float* d_uright, float* d_depthOut, int32_t* d_cellStart, int32_t* d_ids, cudaStream_t st,
                               float4* d_feat = nullptr, float* d_angle = nullptr) {
  const size_t smem = sizeof(int16_t) * (size_t)cap;
  if (smem > 160 * 1024) { set_error("cap too large for the frame kernel (max 81920 keypoints per frame)"); return ORBX_ERR_ARG; }
  if (smem > 32 * 1024)
    ORBX_CUDA(cudaFuncSetAttribute(frame_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  frame_finish_kernel<<<nframes, kFrameThreads, smem, st>>>(C, d_kps, d_counts, cap, d_depth, w, h, depthRow, depthFrame, d_unkps,
                                                           d_uright, d_depthOut, d_cellStart, d_ids, d_feat, d_angle);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

int frame_finish_launch(const orbx_camera* cam, const orbx_keypoint* d_kps, const int32_t* d_counts, int nframes, int cap,
                        const float* d_depth, int w, int h, size_t depthRow, size_t depthFrame, orbx_keypoint* d_unkps, float* d_uright,
                        float* d_depthOut, int32_t* d_cellStart, int32_t* d_ids, float4* d_feat, float* d_angle, cudaStream_t st) {
  CamDev C;
  if (int rc = make_cam(cam, C)) return rc;
  return launch_frame_finish(C, d_kps, d_counts, nframes, cap, d_depth, w, h, depthRow, depthFrame, d_unkps, d_uright, d_depthOut,
                             d_cellStart, d_ids, st, d_feat, d_angle);
}

// feat[i].w = uright[i] after the host filled uRight_ from a host-side depth image (orbx_frame_create)
// `src` may be the frame's pinned host mirror (read over PCIe, 4 KB): the values then also land in the device copy `d_uright`,
// and the separate host->device copy of uRight_ is saved.
__global__ void patch_uright_kernel(const float* __restrict__ src, float* __restrict__ d_uright, float4* __restrict__ feat, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = src[i];
    if (d_uright != src) d_uright[i] = v;
    feat[i].w = v;
  }
}
int frame_patch_uright(const float* src, float* d_uright, float4* d_feat, int n, cudaStream_t st) {
  if (n > 0) patch_uright_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, d_uright, d_feat, n);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

}  // namespace orbx

using namespace orbx;

extern "C" int orbx_frame_finish_device(const orbx_camera* cam, const orbx_keypoint* d_kps, const int32_t* d_counts, int nframes,
                                        int cap, const float* d_depth, int w, int height, size_t depth_row_stride,
                                        size_t depth_frame_stride, orbx_keypoint* d_unkps, float* d_uright, float* d_depth_out,
                                        int32_t* d_cell_start, int32_t* d_ids, int device, void* stream) {
  CamDev C;
  if (int rc = make_cam(cam, C)) return rc;
  if (nframes < 0 || cap <= 0 || !d_kps || !d_counts || !d_unkps || !d_uright || !d_depth_out || !d_cell_start || !d_ids) {
    set_error("bad argument"); return ORBX_ERR_ARG;
  }
  if (d_depth && (w <= 0 || height <= 0 || depth_row_stride < sizeof(float) * (size_t)w || depth_row_stride % sizeof(float))) {
    set_error("bad depth image geometry"); return ORBX_ERR_ARG;
  }
  if (nframes == 0) return ORBX_OK;
  ORBX_CUDA(cudaSetDevice(device));
  return launch_frame_finish(C, d_kps, d_counts, nframes, cap, d_depth, w, height, depth_row_stride, depth_frame_stride, d_unkps,
                             d_uright, d_depth_out, d_cell_start, d_ids, (cudaStream_t)stream);
}

extern "C" int orbx_frame_finish(const orbx_camera* cam, const orbx_keypoint* kps, const int32_t* counts, int nframes, int cap,
                                 const float* depth, int w, int height, size_t depth_row_stride, size_t depth_frame_stride,
                                 orbx_keypoint* unkps, float* uright, float* depth_out, int32_t* cell_start, int32_t* ids,
                                 int device) {
  CamDev C;
  if (int rc = make_cam(cam, C)) return rc;
  if (nframes < 0 || cap <= 0 || !kps || !counts || !unkps || !uright || !depth_out || !cell_start || !ids) {
    set_error("bad argument"); return ORBX_ERR_ARG;
  }
  if (depth && (w <= 0 || height <= 0 || depth_row_stride < sizeof(float) * (size_t)w || depth_row_stride % sizeof(float) ||
                (nframes > 1 && depth_frame_stride < depth_row_stride * (size_t)height))) {
    set_error("bad depth image geometry"); return ORBX_ERR_ARG;
  }
  if (nframes == 0) return ORBX_OK;
  ORBX_CUDA(cudaSetDevice(device));
  const size_t nk = (size_t)nframes * cap;
  const size_t depthBytes = depth ? (size_t)(nframes - 1) * depth_frame_stride + (size_t)height * depth_row_stride : 0;
  const size_t o_kps = 0, o_un = align_up_sz(o_kps + nk * sizeof(orbx_keypoint), 256), o_ur = align_up_sz(o_un + nk * sizeof(orbx_keypoint), 256),
               o_dp = align_up_sz(o_ur + nk * 4, 256), o_ids = align_up_sz(o_dp + nk * 4, 256), o_cs = align_up_sz(o_ids + nk * 4, 256),
               o_cnt = align_up_sz(o_cs + (size_t)nframes * (FNCELL + 1) * 4, 256), o_depth = align_up_sz(o_cnt + (size_t)nframes * 4, 256),
               totalBytes = o_depth + depthBytes + 256;
  static thread_local DevArena arena;      // grow-only: the tracking thread calls this once per frame
  if (arena.reserve(totalBytes, device)) { set_error("scratch allocation failed"); return ORBX_ERR_CUDA; }
  char* db = (char*)arena.take<uint8_t>(totalBytes);
  int rc = ORBX_OK;
  auto run = [&]() -> int {
    ORBX_CUDA(cudaMemcpy(db + o_kps, kps, nk * sizeof(orbx_keypoint), cudaMemcpyHostToDevice));
    ORBX_CUDA(cudaMemcpy(db + o_cnt, counts, (size_t)nframes * 4, cudaMemcpyHostToDevice));
    if (depth) ORBX_CUDA(cudaMemcpy(db + o_depth, depth, depthBytes, cudaMemcpyHostToDevice));
    if (int r = launch_frame_finish(C, (const orbx_keypoint*)(db + o_kps), (const int32_t*)(db + o_cnt), nframes, cap,
                                    depth ? (const float*)(db + o_depth) : nullptr, w, height, depth_row_stride, depth_frame_stride,
                                    (orbx_keypoint*)(db + o_un), (float*)(db + o_ur), (float*)(db + o_dp), (int32_t*)(db + o_cs),
                                    (int32_t*)(db + o_ids), 0))
      return r;
    ORBX_CUDA(cudaMemcpy(unkps, db + o_un, nk * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost));
    ORBX_CUDA(cudaMemcpy(uright, db + o_ur, nk * 4, cudaMemcpyDeviceToHost));
    ORBX_CUDA(cudaMemcpy(depth_out, db + o_dp, nk * 4, cudaMemcpyDeviceToHost));
    ORBX_CUDA(cudaMemcpy(ids, db + o_ids, nk * 4, cudaMemcpyDeviceToHost));
    ORBX_CUDA(cudaMemcpy(cell_start, db + o_cs, (size_t)nframes * (FNCELL + 1) * 4, cudaMemcpyDeviceToHost));
    return ORBX_OK;
  };
  rc = run();
  return rc;
}
