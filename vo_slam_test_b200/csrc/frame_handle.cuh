// frame_handle.cuh -- the device-resident frame (orbx_frame_t): what Frame::Frame builds (frame.cpp:22-32: extractor output,
// undistorted keypoints, depth / uRight, 64x48 grid) kept in HBM so that the tracking thread's searches
// (visualOdometry.cpp:240,265,329,354) read it in place instead of re-uploading it per call.
#pragma once
#include "common.cuh"

struct orbx_extractor;

struct orbx_frame {
  orbx_extractor* owner = nullptr;
  int device = 0, cap = 0, n = 0, nlevels = 0;
  float xmin = 0, xmax = 0, ymin = 0, ymax = 0;
  uint8_t* d_block = nullptr;          // one device allocation, carved below
  uint8_t* h_mirror = nullptr;         // pinned host copy of the packed region (one D2H per frame)
  size_t packed_bytes = 0;
  // packed region: [count (64 B) | kps | desc | unkps | uright | depth]
  int32_t* d_count = nullptr; orbx_keypoint* d_kps = nullptr; uint8_t* d_desc = nullptr; orbx_keypoint* d_unkps = nullptr;
  float* d_uright = nullptr; float* d_depth = nullptr;
  // device-only: what the searches read
  float* d_angle = nullptr; float* d_scale = nullptr; int32_t* d_cellStart = nullptr; int32_t* d_ids = nullptr; float4* d_feat = nullptr;
  // the uRight_ patch of orbx_frame_create runs on the legacy default stream (where the searches run) and reads the mirror in
  // place: the next frame built in this block (on the extractor's own stream) waits for it
  cudaEvent_t patched = nullptr; bool patchPending = false;
  // the captured device chain of orbx_frame_create for this block (same outputs every time) and what it was captured for
  cudaGraphExec_t graph = nullptr; unsigned long long graphKey = 0; int graphLaunches = 0;
};

namespace orbx {

// frame.cu: undistortKeyPoints + findDepth + assignFeaturesToGrid for `nframes` frames on `st`; feat / angle (optional,
// [nframes*cap]) = the per-feature records of the grid-window searches (x, y, octave bits, uRight) and unKeypoints_[i].angle
int frame_finish_launch(const orbx_camera* cam, const orbx_keypoint* d_kps, const int32_t* d_counts, int nframes, int cap,
                        const float* d_depth, int w, int h, size_t depthRow, size_t depthFrame, orbx_keypoint* d_unkps, float* d_uright,
                        float* d_depthOut, int32_t* d_cellStart, int32_t* d_ids, float4* d_feat, float* d_angle, cudaStream_t st);

int frame_patch_uright(const float* src, float* d_uright, float4* d_feat, int n, cudaStream_t st);

}  // namespace orbx
