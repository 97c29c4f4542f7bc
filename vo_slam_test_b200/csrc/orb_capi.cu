// orb_capi.cu -- C ABI of the extractor (include/orb_b200.h): host tables, geometry, workspaces, launch
// sequencing.  Host-side tables mirror the reference constructor (ORBextractor.cpp:414-476) and
// ComputePyramid's sizes (ORBextractor.cpp:1119-1120); everything per-pixel runs in orb_kernels.cu.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "frame_handle.cuh"
#include "orb_extract.cuh"
#include "orb_pattern.h"

namespace orbx {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

// kernels / launchers (orb_kernels.cu)
void launch_resize(const uint8_t* src, int sw, int sh, int spitch, size_t sframe, uint8_t* dst, int dw, int dh,
                   int dpitch, size_t dframe, const ResizeTaps& T, const ResizeTma& R, const CUtensorMap* map, int z0, int nframes,
                   cudaStream_t st);
cudaError_t configure_kernels(const Geom& G);
size_t fast_smem_bytes(const Geom& G);
size_t octree_smem_bytes(const Geom& G);
void launch_pyramid_fused(const PyrLevel* d_plan, const PyrLaunch& P, int* d_sync, size_t smem, bool pdl, cudaStream_t st);
bool pyramid_pdl();
void launch_fast(const Geom& G, const Bufs& B, const TmaSet& TM, int nframes, cudaStream_t st, int level0 = 0, int nlev = -1);
void launch_octree(const Geom& G, const Bufs& B, int nframes, cudaStream_t st, int level0 = 0, int nlev = -1);
void launch_blur(const Geom& G, const Bufs& B, const TmaSet& TM, int nframes, cudaStream_t st);
void launch_orient_desc(const Geom& G, const Bufs& B, const TmaSet& TM, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts,
                        int frame0, int nframes, cudaStream_t st);

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DevTaps {
  ResizeTaps t{};
  ResizeTma rt{};          // box geometry of the TMA-staged resize; rt.use is set when the source map could be encoded
  bool rtFits = false;     // tile origins / box fit the limits (independent of the descriptor)
  void* block = nullptr;
};

}  // namespace orbx

using namespace orbx;

namespace orbx {
int pdl_enabled() {
  static const int v = getenv("ORBX_PDL") ? atoi(getenv("ORBX_PDL")) : 1;
  return v;
}
}  // namespace orbx

struct orbx_extractor {
  orbx_params p{};
  // reference-constructor tables
  std::vector<float> scale, invScale;
  std::vector<int> nfeat;
  int umax[16];
  int maxKp = 0;
  // geometry of the currently configured frame size
  Geom G{};
  bool haveGeom = false;
  int chunk = 0;             // frames the workspace can hold
  // device state
  uint8_t* d_pyr = nullptr; uint8_t* d_blur = nullptr;
  int* d_slotKeyBase = nullptr; int* d_slotCount = nullptr; uint32_t* d_slotKeys = nullptr;
  uint32_t* d_flatKeys = nullptr; uint16_t* d_nodeOf = nullptr; int* d_candCount = nullptr;
  uint32_t* d_sel = nullptr; int* d_selCount = nullptr;
  float2* d_pattern = nullptr; int* d_umax = nullptr;
  std::vector<DevTaps> taps;
  TmaSet tma{};              // FAST tile descriptors (levels >= 1 fixed after configure, level 0 per call)
  CUtensorMap hostMaps[5 * kMaxLevels];   // host copies; d_maps mirrors them in device memory (layout: TmaSet::map)
  bool patchU[kMaxLevels] = {}, patchB[kMaxLevels] = {}, blurSrc[kMaxLevels] = {};   // which orientation / descriptor patch maps are valid
  CUtensorMap* d_maps = nullptr;
  EncodeTiledFn encode = nullptr;
  // staging for the host entry points
  uint8_t* d_in = nullptr; size_t d_in_bytes = 0;
  orbx_keypoint* d_kps = nullptr; uint8_t* d_desc = nullptr; int32_t* d_counts = nullptr;
  uint8_t* d_one = nullptr; uint8_t* h_one = nullptr; int one_cap = 0;   // single-frame path: packed [count | kps | desc], pinned mirror
  size_t d_out_frames = 0; int d_out_cap = 0;
  cudaStream_t stream = nullptr, copyStream = nullptr, backStream = nullptr;
  int32_t* d_qf2 = nullptr; size_t qf_frames = 0;     // pair index 0, 1, 2, ... of the resident extract + match entry point
  PyrLevel* d_plan = nullptr; int* d_pyrSync = nullptr; size_t pyrSyncInts = 0, pyrSmem = 0; bool planDirty = true, planOk = false;   // fused pyramid
  cudaStream_t auxStream = nullptr;           // latency mode: level 0's FAST + quadtree and the blur run here beside the pyramid
  cudaEvent_t evIn = nullptr, evPyr = nullptr, evAux = nullptr;
  int32_t *d_midx = nullptr, *d_md1 = nullptr, *d_md2 = nullptr, *d_qf = nullptr; uint8_t* d_mok = nullptr;
  size_t m_frames = 0; int m_cap = 0;
  long long launches = 0;
  unsigned long long geomEpoch = 0;          // bumped whenever the workspace or the level-0 map is rebuilt: captured graphs of older epochs are stale
  cudaGraphExec_t oneGraph = nullptr; unsigned long long oneKey = 0; int oneLaunches = 0;   // orbx_extract: captured single-frame chain
  bool graphBroken = false;                  // stream capture of the single-frame chain failed once: stay eager
  const uint8_t* map0_base = nullptr; size_t map0_row = 0, map0_frame = 0; int map0_n = 0, map0_want = 0;
  // second lane of the device-resident batch path: a sibling extractor (own workspace, own stream) that takes the second
  // half of a large batch concurrently, so that kernel tails and the latency-bound stages of one lane fill under the
  // throughput-bound stages of the other (measured: +2.4 % on 4096 VGA frames, tools/lanes_probe.py)
  orbx_extractor* lane2 = nullptr;             // first sibling (also the second lane of the host pipeline)
  orbx_extractor* laneX[2] = {nullptr, nullptr};   // third / fourth lane (ORBX_LANES=3|4, A/B only)
  cudaEvent_t laneJoinX[2] = {nullptr, nullptr};
  cudaEvent_t laneFork = nullptr, laneJoin = nullptr;
  // device-resident frames (orbx_frame_t): blocks handed back by orbx_frame_destroy are reused by the next orbx_frame_create
  std::vector<orbx_frame*> framePool, framesAll;      // framesAll: every block not yet freed (live ones included)
  int framesLive = 0;
  // last chunk info for the debug taps
  const uint8_t* last_img0 = nullptr; size_t last_rowStride = 0, last_frameStride = 0; int last_frames = 0;
};

namespace {

void free_frame_block(orbx_frame* f);

// Frames per chunk.  Device-resident batches: 512 (fewer, larger launches: +3 % over 256, no gain beyond).  Host-pipelined
// batches: 128 on two compute lanes, because the upload of chunk c+1 overlaps the kernels of chunk c and the pipeline fill/drain
// grows with the chunk (measured end to end, ms per 4096-frame step: 24.5 at 128, 25.0 at 256, 26.6 at 512; one lane alone cannot
// keep up with the link below 256).  ORBX_CHUNK overrides both, ORBX_CHUNK_HOST the latter.
constexpr int kDefaultChunk = 512, kDefaultHostChunk = 128;   // host pipeline: 128-frame chunks on two lanes (tools/e2e_probe.py)

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// The workspace (pyramid, blurred pyramid, candidate lists) is about 8 bytes per input pixel per frame of the chunk; the
// default chunk is capped so that it stays below ~4 GiB (VGA: 512 frames = 1 GB, 1080p: 256, 4K: 64).  An explicit
// ORBX_CHUNK / ORBX_CHUNK_HOST is taken as given.
int capped_chunk(int dflt, int w, int hgt) {
  const long long per_frame = 8LL * w * hgt;
  const long long cap = std::max(16LL, (4LL << 30) / std::max(per_frame, 1LL));
  int c = (int)std::min<long long>(dflt, cap);
  if (c >= 64) c &= ~31;
  return std::max(1, c);
}
constexpr int kSplitFrames = 8;       // latency mode of run_chunk (two streams) up to this many frames per call
constexpr int kLaneMinFrames = 32;     // below 2 x this many frames a second lane is not worth its fork/join
int resident_chunk(int w, int hgt) {
  const int e = env_int("ORBX_CHUNK", 0);
  return e > 0 ? e : capped_chunk(kDefaultChunk, w, hgt);
}
int host_chunk(int w, int hgt) {
  const int e = env_int("ORBX_CHUNK_HOST", env_int("ORBX_CHUNK", 0));
  return e > 0 ? e : capped_chunk(kDefaultHostChunk, w, hgt);
}

// Chunk schedule of the host-pipelined path: a short first chunk (kernels start after a quarter-chunk upload) and a short
// last one (only a quarter chunk of kernels + download is left when the last upload ends), full chunks in between.
std::vector<std::pair<int, int>> host_schedule(int nframes, int chunk) {
  std::vector<std::pair<int, int>> v;
  const int q = std::max(1, chunk / 4);
  if (nframes <= chunk + 2 * q || !env_int("ORBX_HOST_RAMP", 1)) {
    for (int f0 = 0; f0 < nframes; f0 += chunk) v.emplace_back(f0, std::min(chunk, nframes - f0));
    return v;
  }
  int f = 0;
  v.emplace_back(f, q); f += q;
  while (nframes - f - q > chunk) { v.emplace_back(f, chunk); f += chunk; }
  if (nframes - f - q > 0) { v.emplace_back(f, nframes - f - q); f = nframes - q; }
  v.emplace_back(f, q);
  return v;
}

void free_workspace(orbx_extractor* h) {
  cudaFree(h->d_pyr); cudaFree(h->d_blur); cudaFree(h->d_slotKeyBase); cudaFree(h->d_slotCount);
  cudaFree(h->d_slotKeys); cudaFree(h->d_flatKeys); cudaFree(h->d_nodeOf); cudaFree(h->d_candCount);
  cudaFree(h->d_sel); cudaFree(h->d_selCount);
  h->d_pyr = h->d_blur = nullptr; h->d_slotKeyBase = nullptr; h->d_slotCount = nullptr; h->d_slotKeys = nullptr;
  h->d_flatKeys = nullptr; h->d_nodeOf = nullptr; h->d_candCount = nullptr; h->d_sel = nullptr; h->d_selCount = nullptr;
  for (auto& t : h->taps) cudaFree(t.block);
  h->taps.clear();
  cudaFree(h->d_plan); cudaFree(h->d_pyrSync); h->d_plan = nullptr; h->d_pyrSync = nullptr; h->pyrSyncInts = 0; h->planDirty = true; h->planOk = false;
  h->haveGeom = false; h->chunk = 0;
  h->geomEpoch++;
  h->map0_base = nullptr; h->map0_n = 0;
  for (int l = 0; l < kMaxLevels; ++l) { h->tma.use[l] = 0; h->patchU[l] = h->patchB[l] = h->blurSrc[l] = false; }
  h->tma.usePatch = 0; h->tma.useBlur = 0;
}

// cv::resize tap tables for one axis
void axis_taps(int ssize, int dsize, std::vector<int>& ofs, std::vector<short>& a0, std::vector<short>& a1) {
  ofs.resize(dsize); a0.resize(dsize); a1.resize(dsize);
  const double scale = 1.0 / ((double)dsize / ssize);
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    ofs[d] = s;
    a0[d] = (short)cv_round_f((1.f - f) * 2048.f);
    a1[d] = (short)cv_round_f(f * 2048.f);
  }
}

// One rank-3 (x bytes, y rows, frame) tiled tensor map over a w x h x nframes byte image into slot `idx` of d_maps.
bool encode_map(orbx_extractor* h, int idx, const uint8_t* base, int w, int hgt, size_t pitch, size_t frameStride, int nframes,
                int boxW, int boxH) {
  if (!h->encode || !h->d_maps || env_int("ORBX_NO_TMA", 0)) return false;
  if ((((uintptr_t)base) | pitch | frameStride) & 15) return false;      // TMA needs 16-byte aligned base and strides
  if (boxW > 256 || boxH > 256) return false;
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)hgt, (cuuint64_t)std::max(nframes, 1)};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameStride};
  cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = h->encode(&h->hostMaps[idx], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  // stream-ordered upload on the legacy default stream would serialise; a synchronous copy is fine (rare event)
  return cudaMemcpy(h->d_maps + idx, &h->hostMaps[idx], sizeof(CUtensorMap), cudaMemcpyHostToDevice) == cudaSuccess;
}

void refresh_use_patch(orbx_extractor* h) {
  int ok = 1;
  int okb = 1;
  for (int l = 0; l < h->G.nlevels; ++l) { ok &= (h->patchU[l] && h->patchB[l]) ? 1 : 0; okb &= h->blurSrc[l] ? 1 : 0; }
  h->tma.usePatch = ok;
  h->tma.useBlur = okb;
}

// Descriptors of an UNBLURRED level: the FAST kernel's shared-memory tile and the 48x31 orientation patch.
bool encode_level_map(orbx_extractor* h, int l, const uint8_t* base, size_t pitch, size_t frameStride, int nframes) {
  const LevelGeom& L = h->G.L[l];
  h->tma.use[l] = encode_map(h, l, base, L.w, L.h, pitch, frameStride, nframes, h->G.fastTileW, h->G.fastTileH) ? 1 : 0;
  h->patchU[l] = encode_map(h, kMaxLevels + l, base, L.w, L.h, pitch, frameStride, nframes, kOdUW, kOdUH);
  h->blurSrc[l] = encode_map(h, 4 * kMaxLevels + l, base, L.w, L.h, pitch, frameStride, nframes, kBtBoxW, kBtBoxH);
  if (l + 1 < h->G.nlevels && l + 1 < (int)h->taps.size()) {       // this level is the resize source of level l + 1
    DevTaps& T = h->taps[l + 1];
    T.rt.use = T.rtFits && encode_map(h, 3 * kMaxLevels + l + 1, base, L.w, L.h, pitch, frameStride, nframes, T.rt.boxW, T.rt.boxH) ? 1 : 0;
    h->planDirty = true;
  }
  refresh_use_patch(h);
  return h->tma.use[l] != 0;
}

// Build the per-size geometry (cell grid, slots, capacities, workspace offsets) and allocate for `chunk` frames.
int configure(orbx_extractor* h, int W, int H, int chunk) {
  if (h->haveGeom && h->G.W == W && h->G.H == H && h->chunk >= chunk) return ORBX_OK;
  free_workspace(h);
  Geom& G = h->G;
  memset(&G, 0, sizeof(G));
  const int nl = h->p.nlevels;
  G.nlevels = nl; G.W = W; G.H = H; G.iniTh = h->p.ini_th_fast; G.minTh = h->p.min_th_fast;
  size_t pyrBytes = 0, blurBytes = 0;
  int slot = 0, keyBase = 0, selBase = 0, blurTile = 0, bwTile = 0, slotKeys = 0;
  std::vector<int> slotKeyBase;
  G.nodeCap = 0; G.maxSlotsPerLevel = 0; G.fastTileW = 0; G.fastTileH = 0;
  for (int l = 0; l < nl; ++l) {
    LevelGeom& L = G.L[l];
    const float s = h->invScale[l];
    L.w = cv_round_f((float)W * s);                    // ORBextractor.cpp:1119-1120
    L.h = cv_round_f((float)H * s);
    L.regW = L.w - 2 * kMinBorder; L.regH = L.h - 2 * kMinBorder;
    if (L.regW >= 30 && L.regH >= 30 && ((L.regW + L.regW / 30 - 1) / (L.regW / 30) > 63 || (L.regH + L.regH / 30 - 1) / (L.regH / 30) > 63)) { set_error("cell size out of range"); return ORBX_ERR_SHAPE; }
    if (L.regW < 30 || L.regH < 30) { set_error("image too small for the pyramid: a level's FAST region is < 30 px"); return ORBX_ERR_SHAPE; }
    if (L.regW > 4095 || L.regH > 4095) { set_error("image too large (level FAST region > 4095 px)"); return ORBX_ERR_SHAPE; }
    L.nCols = L.regW / 30; L.nRows = L.regH / 30;      // :791-792 (float division then truncation == integer division here)
    L.wCell = (L.regW + L.nCols - 1) / L.nCols;        // :793-794
    L.hCell = (L.regH + L.nRows - 1) / L.nRows;
    L.groups = (L.nCols + kCellsPerCta - 1) / kCellsPerCta;
    L.groupsMagic = 0xFFFFFFFFu / (uint32_t)L.groups + 1;
    L.slot0 = slot; L.nSlots = L.nRows * L.groups;
    G.maxSlotsPerLevel = std::max(G.maxSlotsPerLevel, L.nSlots);
    L.nFeat = h->nfeat[l];
    L.nIni = (int)roundf((float)L.regW / (float)L.regH);   // :549
    if (L.nIni < 1 || L.nIni > kMaxRoots) { set_error("unsupported aspect ratio (the reference's quadtree needs 1..16 root nodes)"); return ORBX_ERR_SHAPE; }
    L.hX = (float)L.regW / L.nIni;                         // :551
    for (int i = 0; i <= L.nIni; ++i) L.rootX[i] = (int)(L.hX * (float)i);   // :561-562
    L.selCap = std::max(L.nFeat + 3, 4 * L.nIni);
    L.selBase = selBase; selBase += L.selCap;
    G.nodeCap = std::max(G.nodeCap, L.selCap + 1);
    // slot capacities: NMS keeps no two 8-adjacent pixels inside a cell -> ceil(w/2)*ceil(h/2) per cell
    int levelKeys = 0;
    for (int ci = 0; ci < L.nRows; ++ci) {
      const int ih = std::max(0, std::min(ci * L.hCell + L.hCell + 6, L.regH) - ci * L.hCell - 6);
      for (int g = 0; g < L.groups; ++g) {
        int capSlot = 0;
        for (int j = g * kCellsPerCta; j < std::min((g + 1) * kCellsPerCta, L.nCols); ++j) {
          const int iw = std::max(0, std::min(j * L.wCell + L.wCell + 6, L.regW) - j * L.wCell - 6);
          capSlot += ((iw + 1) / 2) * ((ih + 1) / 2);
        }
        slotKeyBase.push_back(slotKeys);
        slotKeys += capSlot; levelKeys += capSlot;
        G.fastSurvCap = std::max(G.fastSurvCap, capSlot);
      }
    }
    slot += L.nSlots;
    L.keyBase = keyBase; L.keyCap = levelKeys; keyBase += levelKeys;
    L.pitch = align_up(L.w, 64); L.bpitch = align_up(L.w, 64);
    if (l > 0) { L.pyrOff = pyrBytes; pyrBytes += (size_t)chunk * L.h * L.pitch; }
    L.blurOff = blurBytes; blurBytes += (size_t)chunk * L.h * L.bpitch;
    L.blurTilesX = (L.w + 63) / 64; L.blurTilesY = (L.h + 25) / 26;   // kBlurTW x kBlurTH of orb_kernels.cu
    L.blurTile0 = blurTile; blurTile += L.blurTilesX;      // one CTA per 64-px column strip
    L.bwTilesX = (L.w + 63) / 64; L.bwTilesY = (L.h + 207) / 208;     // kBwTileH of orb_kernels.cu
    L.bwTile0 = bwTile; bwTile += L.bwTilesX * L.bwTilesY;
    L.scale = h->scale[l];
    L.kpSize = (float)(int)(31 * h->scale[l]);             // :845
    G.fastTileW = std::max(G.fastTileW, align_up(std::min(kCellsPerCta, L.nCols) * L.wCell + 6 + 15, 16));
    G.fastTileH = std::max(G.fastTileH, L.hCell + 6);
    G.fastScW = std::max(G.fastScW, align_up(std::min(kCellsPerCta, L.nCols) * (L.wCell + 2), 16));   // every cell framed by zeros
    G.fastScH = std::max(G.fastScH, L.hCell + 2);
    G.fastCellPix = std::max(G.fastCellPix, align_up(L.wCell * L.hCell, 8));
    G.fastCellQuads = std::max(G.fastCellQuads, align_up(((L.wCell + 3) / 4) * L.hCell, 8));
    G.fastCellSurv = std::max(G.fastCellSurv, align_up(((L.wCell + 1) / 2) * ((L.hCell + 1) / 2), 4));
  }
  slotKeyBase.push_back(slotKeys);
  G.totalSlots = slot; G.slotKeysPerFrame = slotKeys; G.keysPerFrame = keyBase; G.selPerFrame = selBase;
  G.blurTiles = blurTile;
  G.bwTiles = bwTile;
  if (G.nodeCap > 60000) { set_error("nfeatures per level too large"); return ORBX_ERR_ARG; }
  if (fast_smem_bytes(G) > 200 * 1024 || octree_smem_bytes(G) > 200 * 1024) { set_error("shape needs too much shared memory"); return ORBX_ERR_SHAPE; }

  ORBX_CUDA(cudaSetDevice(h->p.device));
  ORBX_CUDA(configure_kernels(G));
  ORBX_CUDA(cudaMalloc(&h->d_pyr, std::max<size_t>(pyrBytes, 256)));
  ORBX_CUDA(cudaMalloc(&h->d_blur, blurBytes));
  ORBX_CUDA(cudaMalloc(&h->d_slotKeyBase, sizeof(int) * slotKeyBase.size()));
  ORBX_CUDA(cudaMemcpy(h->d_slotKeyBase, slotKeyBase.data(), sizeof(int) * slotKeyBase.size(), cudaMemcpyHostToDevice));
  ORBX_CUDA(cudaMalloc(&h->d_slotCount, sizeof(int) * (size_t)chunk * G.totalSlots));
  ORBX_CUDA(cudaMalloc(&h->d_slotKeys, sizeof(uint32_t) * (size_t)chunk * G.slotKeysPerFrame));
  ORBX_CUDA(cudaMalloc(&h->d_flatKeys, sizeof(uint32_t) * (size_t)chunk * G.keysPerFrame));
  ORBX_CUDA(cudaMalloc(&h->d_nodeOf, sizeof(uint16_t) * (size_t)chunk * G.keysPerFrame));
  ORBX_CUDA(cudaMalloc(&h->d_candCount, sizeof(int) * (size_t)chunk * nl));
  ORBX_CUDA(cudaMalloc(&h->d_sel, sizeof(uint32_t) * (size_t)chunk * G.selPerFrame));
  ORBX_CUDA(cudaMalloc(&h->d_selCount, sizeof(int) * (size_t)chunk * nl));
  // resize tap tables per level >= 1
  h->taps.resize(nl);
  for (int l = 1; l < nl; ++l) {
    const LevelGeom& S = G.L[l - 1];
    const LevelGeom& D = G.L[l];
    std::vector<int> xo, yo; std::vector<short> xa0, xa1, yb0, yb1;
    axis_taps(S.w, D.w, xo, xa0, xa1);
    axis_taps(S.h, D.h, yo, yb0, yb1);
    const size_t bx = align_up_sz(sizeof(int) * D.w, 256), bs = align_up_sz(sizeof(short) * D.w, 256);
    const size_t by = align_up_sz(sizeof(int) * D.h, 256), bt = align_up_sz(sizeof(short) * D.h, 256);
    const int nquad = (D.w + 3) / 4;
    const size_t bq = align_up_sz(sizeof(int4) * nquad, 256), bw = align_up_sz(sizeof(uint4) * nquad, 256);
    std::vector<int4> quad(nquad);
    std::vector<uint4> xw(nquad);
    int quadOk = 1;
    for (int q = 0; q < nquad; ++q) {
      const int s0 = xo[4 * q];
      uint32_t sels = 0, wv[4];
      for (int i = 0; i < 4; ++i) {
        const int dx = std::min(4 * q + i, D.w - 1);
        const int o = xo[dx] - s0;
        if (o < 0 || o + 1 > 7) quadOk = 0;
        sels |= (uint32_t)((o & 7) | (((o + 1) & 7) << 4)) << (8 * i);
        wv[i] = (uint32_t)(uint16_t)xa0[dx] | ((uint32_t)(uint16_t)xa1[dx] << 16);
      }
      quad[q] = make_int4(s0 & ~3, (s0 & 3) * 8, (int)sels, 0);
      xw[q] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    {   // TMA-staged resize: source box origin of every 128-column / 64-row destination tile, and the box that covers all
      DevTaps& T = h->taps[l];
      T.rt = ResizeTma{};
      // small levels: half-height tiles, so that a chunk still makes a few waves of CTAs (levels 6-7 of a VGA pyramid were
      // single partial waves: issue 65 % against 85-90 % on the large levels)
      const int ntx = (D.w + kRzTileW - 1) / kRzTileW;
      const long long ctasFull = (long long)ntx * ((D.h + kRzTileH - 1) / kRzTileH) * chunk;
      T.rt.rows = (ctasFull < 6000 && env_int("ORBX_RESIZE_SMALL_TILES", 1)) ? kRwRows / 2 : kRwRows;
      // latency mode (a handful of frames per call): the 7 dependent launches are each one partial wave, so their duration is
      // the per-thread walk; quarter-height tiles shorten it (ORBX_RESIZE_LAT_ROWS, A/B)
      if (chunk <= kSplitFrames) T.rt.rows = std::max(1, std::min(kRwRows, env_int("ORBX_RESIZE_LAT_ROWS", kRwRows / 4)));
      while (T.rt.rows < kRwRows && (D.h + 4 * T.rt.rows - 1) / (4 * T.rt.rows) > kRzMaxTY) T.rt.rows *= 2;   // tall level: keep the origin table in range
      const int tileH = 4 * T.rt.rows;
      const int nty = (D.h + tileH - 1) / tileH;
      bool fits = quadOk && ntx <= kRzMaxTX && nty <= kRzMaxTY;
      int boxW = 16, boxH = 1;
      for (int tx = 0; fits && tx < ntx; ++tx) {
        const int q0 = tx * (kRzTileW / 4), q1 = std::min(q0 + kRzTileW / 4, nquad) - 1;
        const int x0 = quad[q0].x & ~15, x1 = quad[q1].x + 12;
        if (x0 > 32767) fits = false;
        T.rt.x0[tx] = (short)x0;
        boxW = std::max(boxW, align_up(x1 - x0, 16));
      }
      for (int ty = 0; fits && ty < nty; ++ty) {
        const int d0 = ty * tileH, d1 = std::min(d0 + tileH, D.h) - 1;
        const int y0 = yo[d0], y1 = std::min(yo[d1] + 1, S.h - 1);
        if (y0 > 32767) fits = false;
        T.rt.y0[ty] = (short)y0;
        boxH = std::max(boxH, y1 - y0 + 1);
      }
      T.rt.boxW = boxW; T.rt.boxH = boxH;
      T.rtFits = fits && boxW <= 256 && boxH <= 256 && (size_t)boxW * boxH <= 48 * 1024;
    }
    std::vector<uint8_t> blob(bx + 2 * bs + by + 2 * bt + bq + bw, 0);
    memcpy(blob.data(), xo.data(), sizeof(int) * D.w);
    memcpy(blob.data() + bx, xa0.data(), sizeof(short) * D.w);
    memcpy(blob.data() + bx + bs, xa1.data(), sizeof(short) * D.w);
    memcpy(blob.data() + bx + 2 * bs, yo.data(), sizeof(int) * D.h);
    memcpy(blob.data() + bx + 2 * bs + by, yb0.data(), sizeof(short) * D.h);
    memcpy(blob.data() + bx + 2 * bs + by + bt, yb1.data(), sizeof(short) * D.h);
    memcpy(blob.data() + bx + 2 * bs + by + 2 * bt, quad.data(), sizeof(int4) * nquad);
    memcpy(blob.data() + bx + 2 * bs + by + 2 * bt + bq, xw.data(), sizeof(uint4) * nquad);
    DevTaps& T = h->taps[l];
    ORBX_CUDA(cudaMalloc(&T.block, blob.size()));
    ORBX_CUDA(cudaMemcpy(T.block, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    uint8_t* b = (uint8_t*)T.block;
    T.t.xofs = (int*)b; T.t.xa0 = (short*)(b + bx); T.t.xa1 = (short*)(b + bx + bs);
    T.t.yofs = (int*)(b + bx + 2 * bs); T.t.yb0 = (short*)(b + bx + 2 * bs + by); T.t.yb1 = (short*)(b + bx + 2 * bs + by + bt);
    T.t.quad = (int4*)(b + bx + 2 * bs + by + 2 * bt); T.t.xw = (uint4*)(b + bx + 2 * bs + by + 2 * bt + bq);
    T.t.quadOk = quadOk;
  }
  h->haveGeom = true;
  h->chunk = chunk;
  for (int l = 1; l < nl; ++l)
    encode_level_map(h, l, h->d_pyr + G.L[l].pyrOff, (size_t)G.L[l].pitch, (size_t)G.L[l].h * G.L[l].pitch, chunk);
  for (int l = 0; l < nl; ++l)                 // blurred levels: the 64x37 descriptor patch
    h->patchB[l] = encode_map(h, 2 * kMaxLevels + l, h->d_blur + G.L[l].blurOff, G.L[l].w, G.L[l].h, (size_t)G.L[l].bpitch,
                              (size_t)G.L[l].h * G.L[l].bpitch, chunk, kOdBW, kOdBH);
  refresh_use_patch(h);
  return ORBX_OK;
}

Bufs make_bufs(orbx_extractor* h, const uint8_t* img0, size_t rowStride, size_t frameStride) {
  Bufs B{};
  B.img0 = img0; B.rowStride0 = rowStride; B.frameStride0 = frameStride;
  B.pyr = h->d_pyr; B.blur = h->d_blur; B.slotKeyBase = h->d_slotKeyBase; B.slotCount = h->d_slotCount;
  B.slotKeys = h->d_slotKeys; B.flatKeys = h->d_flatKeys; B.nodeOf = h->d_nodeOf; B.candCount = h->d_candCount;
  B.sel = h->d_sel; B.selCount = h->d_selCount; B.pattern = h->d_pattern; B.umax = h->d_umax;
  return B;
}

// (Re)build the device plan of the fused pyramid kernel: one record per destination level.  planOk: every level can be staged
// by TMA (descriptor encodable, tile origins in range); otherwise the per-level launches are used.
int refresh_plan(orbx_extractor* h) {
  const Geom& G = h->G;
  const int nlev = G.nlevels - 1;
  h->planDirty = false;
  h->planOk = false;
  if (nlev < 1 || !h->d_maps) return ORBX_OK;
  std::vector<PyrLevel> plan(nlev);
  size_t smem = 0;
  for (int l = 1; l < G.nlevels; ++l) {
    const DevTaps& T = h->taps[l];
    if (!T.t.quadOk || !T.rt.use) return ORBX_OK;
    const LevelGeom& D = G.L[l];
    PyrLevel& P = plan[l - 1];
    memset(&P, 0, sizeof(P));
    P.map = h->d_maps + 3 * kMaxLevels + l;
    P.dst = h->d_pyr + D.pyrOff; P.dframe = (size_t)D.h * D.pitch;
    P.dw = D.w; P.dh = D.h; P.dpitch = D.pitch; P.sh = G.L[l - 1].h;
    P.ntx = (D.w + kRzTileW - 1) / kRzTileW; P.nty = (D.h + 4 * T.rt.rows - 1) / (4 * T.rt.rows);
    if (P.nty > kPyrSyncStride) return ORBX_OK;
    if (l > 1) { P.srcTileH = 4 * h->taps[l - 1].rt.rows; P.srcNtx = plan[l - 2].ntx; } else { P.srcTileH = 1; P.srcNtx = 0; }
    P.T = T.t; P.R = T.rt;
    smem = std::max(smem, (size_t)T.rt.boxW * T.rt.boxH);
  }
  if (!h->d_plan) ORBX_CUDA(cudaMalloc(&h->d_plan, sizeof(PyrLevel) * kMaxLevels));
  const size_t ints = 1 + (size_t)h->chunk * nlev * kPyrSyncStride;
  if (h->pyrSyncInts < ints) {
    cudaFree(h->d_pyrSync); h->d_pyrSync = nullptr; h->pyrSyncInts = 0;
    ORBX_CUDA(cudaMalloc(&h->d_pyrSync, sizeof(int) * ints));
    h->pyrSyncInts = ints;
  }
  // earlier launches of this handle may still read the old plan: drain them (only when the geometry or the input buffer changes)
  ORBX_CUDA(cudaDeviceSynchronize());
  ORBX_CUDA(cudaMemcpy(h->d_plan, plan.data(), sizeof(PyrLevel) * nlev, cudaMemcpyHostToDevice));
  h->pyrSmem = smem;
  h->planOk = true;
  return ORBX_OK;
}

// The whole extractor for frames [frame0, frame0+n) of a device-resident batch (n <= h->chunk).
int run_chunk(orbx_extractor* h, const uint8_t* d_imgs, size_t rowStride, size_t frameStride, int frame0, int n,
              orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, cudaStream_t st,
              cudaEvent_t* ev = nullptr /* 6 events: before pyramid, fast, octree, blur, orient_desc, end */) {
  const Geom& G = h->G;
  if (h->map0_base != d_imgs || h->map0_row != rowStride || h->map0_frame != frameStride || h->map0_n < frame0 + n) {
    // level 0 is the caller's buffer: (re)describe it; the map covers frames [0, frame0+n) at least.  The descriptors live
    // in device memory and are rewritten in place, so earlier launches of this handle (any stream) that may still read
    // the old ones are drained first; this only happens when the caller switches input buffers.
    if (h->map0_base) cudaDeviceSynchronize();
    encode_level_map(h, 0, d_imgs, rowStride, frameStride, std::max(frame0 + n, h->map0_want));
    h->geomEpoch++;
    h->map0_base = d_imgs; h->map0_row = rowStride; h->map0_frame = frameStride; h->map0_n = std::max(frame0 + n, h->map0_want);
  }
  const uint8_t* img0 = d_imgs + (size_t)frame0 * frameStride;
  Bufs B = make_bufs(h, img0, rowStride, frameStride);
  // Latency mode (a handful of frames per call: orbx_extract, orbx_frame_create): the chain is bound by the dependent launches,
  // not by throughput, and level 0 needs nothing from the pyramid.  Level 0's FAST + quadtree (the longest single CTA of the
  // chain) and then the blur of all levels run on a second stream beside the 7 resize launches and levels >= 1.
  static const bool splitOn = env_int("ORBX_SPLIT", 1) != 0;
  const bool split = splitOn && !ev && n <= kSplitFrames && G.nlevels > 1;
  if (split && !h->auxStream) {
    ORBX_CUDA(cudaStreamCreateWithFlags(&h->auxStream, cudaStreamNonBlocking));
    ORBX_CUDA(cudaEventCreateWithFlags(&h->evIn, cudaEventDisableTiming));
    ORBX_CUDA(cudaEventCreateWithFlags(&h->evPyr, cudaEventDisableTiming));
    ORBX_CUDA(cudaEventCreateWithFlags(&h->evAux, cudaEventDisableTiming));
  }
  h->tma.frame0 = frame0;
  if (split) {
    cudaStream_t sa = h->auxStream;
    ORBX_CUDA(cudaEventRecord(h->evIn, st));
    ORBX_CUDA(cudaStreamWaitEvent(sa, h->evIn, 0));
    launch_fast(G, B, h->tma, n, sa, 0, 1);
    launch_octree(G, B, n, sa, 0, 1);
  }
  if (ev) cudaEventRecord(ev[0], st);
  nvtxRangePushA("orbx:pyramid");          // NVTX ranges per stage (host side of the launches): nsys timelines of the lanes
  static const int fusedMode = env_int("ORBX_PYR_FUSED", 0);   // 0: per-level launches, 1: one launch, 2: one launch in latency mode only
  bool fused = false;
  if (fusedMode == 1 || (fusedMode == 2 && n <= kSplitFrames)) {
    if (h->planDirty) { if (int rc = refresh_plan(h)) return rc; }
    fused = h->planOk;
  }
  if (fused) {                                 // ComputePyramid (ORBextractor.cpp:1114-1135) as ONE kernel
    PyrLaunch P{};
    P.nlev = G.nlevels - 1; P.nframes = n; P.z0 = frame0;
    int t = 0;
    for (int l = 1; l < G.nlevels; ++l) {
      P.start[l - 1] = t;
      const int rows = h->taps[l].rt.rows;
      t += n * ((G.L[l].w + kRzTileW - 1) / kRzTileW) * ((G.L[l].h + 4 * rows - 1) / (4 * rows));
    }
    P.start[P.nlev] = t;
    ORBX_CUDA(cudaMemsetAsync(h->d_pyrSync, 0, sizeof(int) * (1 + (size_t)n * P.nlev * kPyrSyncStride), st));
    launch_pyramid_fused(h->d_plan, P, h->d_pyrSync, h->pyrSmem, pyramid_pdl(), st);
  }
  for (int l = 1; l < G.nlevels && !fused; ++l) {        // ComputePyramid: level l from level l-1 (ORBextractor.cpp:1129)
    const LevelGeom& S = G.L[l - 1];
    const LevelGeom& D = G.L[l];
    const uint8_t* src = (l == 1) ? img0 : h->d_pyr + S.pyrOff;
    const int spitch = (l == 1) ? (int)rowStride : S.pitch;
    const size_t sframe = (l == 1) ? frameStride : (size_t)S.h * S.pitch;
    launch_resize(src, S.w, S.h, spitch, sframe, h->d_pyr + D.pyrOff, D.w, D.h, D.pitch, (size_t)D.h * D.pitch,
                  h->taps[l].t, h->taps[l].rt, h->d_maps ? h->d_maps + 3 * kMaxLevels + l : nullptr, l == 1 ? frame0 : 0, n, st);
  }
  nvtxRangePop();
  if (split) {
    ORBX_CUDA(cudaEventRecord(h->evPyr, st));
    ORBX_CUDA(cudaStreamWaitEvent(h->auxStream, h->evPyr, 0));
    launch_blur(G, B, h->tma, n, h->auxStream);
    ORBX_CUDA(cudaEventRecord(h->evAux, h->auxStream));
    launch_fast(G, B, h->tma, n, st, 1, -1);
    launch_octree(G, B, n, st, 1, -1);
    ORBX_CUDA(cudaStreamWaitEvent(st, h->evAux, 0));
    launch_orient_desc(G, B, h->tma, d_kps, d_desc, cap, d_counts, frame0, n, st);
    h->launches += 2;
  } else {
  if (ev) cudaEventRecord(ev[1], st);
  nvtxRangePushA("orbx:fast");
  launch_fast(G, B, h->tma, n, st);
  nvtxRangePop();
  if (ev) cudaEventRecord(ev[2], st);
  nvtxRangePushA("orbx:quadtree");
  launch_octree(G, B, n, st);
  nvtxRangePop();
  if (ev) cudaEventRecord(ev[3], st);
  nvtxRangePushA("orbx:blur");
  launch_blur(G, B, h->tma, n, st);
  nvtxRangePop();
  if (ev) cudaEventRecord(ev[4], st);
  nvtxRangePushA("orbx:orient_desc");
  launch_orient_desc(G, B, h->tma, d_kps, d_desc, cap, d_counts, frame0, n, st);
  nvtxRangePop();
  if (ev) cudaEventRecord(ev[5], st);
  }
  h->launches += (fused ? 1 : G.nlevels - 1) + 4;
  h->last_img0 = img0; h->last_rowStride = rowStride; h->last_frameStride = frameStride; h->last_frames = n;
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

void free_frame_block(orbx_frame* f) {
  if (!f) return;
  if (f->patched) cudaEventDestroy(f->patched);
  if (f->graph) cudaGraphExecDestroy(f->graph);
  cudaFree(f->d_block);
  if (f->h_mirror) cudaFreeHost(f->h_mirror);
  delete f;
}

// carve one device block: [packed region | angle | scale | cellStart | ids | feat]
int alloc_frame_block(orbx_extractor* h, orbx_frame** out) {
  const int cap = h->maxKp, nl = h->p.nlevels;
  orbx_frame* f = new orbx_frame();
  f->owner = h; f->device = h->p.device; f->cap = cap; f->nlevels = nl;
  size_t used = 0;
  auto add = [&](size_t bytes) { used = align_up_sz(used, 256); size_t o = used; used += bytes; return o; };
  const size_t o_cnt = add(64), o_kps = add(sizeof(orbx_keypoint) * cap), o_desc = add((size_t)32 * cap), o_un = add(sizeof(orbx_keypoint) * cap),
               o_ur = add(sizeof(float) * cap), o_dp = add(sizeof(float) * cap),
               o_cs = add(sizeof(int32_t) * (ORBX_GRID_COLS * ORBX_GRID_ROWS + 1)), o_ids = add(sizeof(int32_t) * cap);
  f->packed_bytes = align_up_sz(used, 256);
  const size_t o_ang = add(sizeof(float) * cap), o_sc = add(sizeof(float) * nl), o_feat = add(sizeof(float4) * cap);
  if (cudaMalloc(&f->d_block, used + 256) != cudaSuccess || cudaHostAlloc(&f->h_mirror, f->packed_bytes, cudaHostAllocDefault) != cudaSuccess) {
    set_error("frame block allocation failed");
    free_frame_block(f);
    return ORBX_ERR_CUDA;
  }
  uint8_t* b = f->d_block;
  f->d_count = (int32_t*)(b + o_cnt); f->d_kps = (orbx_keypoint*)(b + o_kps); f->d_desc = b + o_desc; f->d_unkps = (orbx_keypoint*)(b + o_un);
  f->d_uright = (float*)(b + o_ur); f->d_depth = (float*)(b + o_dp); f->d_angle = (float*)(b + o_ang); f->d_scale = (float*)(b + o_sc);
  f->d_cellStart = (int32_t*)(b + o_cs); f->d_ids = (int32_t*)(b + o_ids); f->d_feat = (float4*)(b + o_feat);
  if (cudaMemcpy(f->d_scale, h->scale.data(), sizeof(float) * nl, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("frame block initialisation failed");
    free_frame_block(f);
    return ORBX_ERR_CUDA;
  }
  h->framesAll.push_back(f);
  *out = f;
  return ORBX_OK;
}

int check_handle(orbx_handle h) {
  if (!h) { set_error("null handle"); return ORBX_ERR_ARG; }
  return ORBX_OK;
}

}  // namespace

extern "C" {

const char* orbx_last_error(void) { return g_err.c_str(); }

int orbx_device_count(int* n) {
  if (!n) return ORBX_ERR_ARG;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *n = 0; set_error(cudaGetErrorString(e)); return ORBX_ERR_CUDA; }
  *n = c;
  return ORBX_OK;
}

int orbx_host_alloc(size_t bytes, int write_combined, void** out) {
  if (!out || bytes == 0) { set_error("bad argument"); return ORBX_ERR_ARG; }
  *out = nullptr;
  ORBX_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return ORBX_OK;
}
int orbx_host_free(void* p) {
  if (p) ORBX_CUDA(cudaFreeHost(p));
  return ORBX_OK;
}

int orbx_create(const orbx_params* p, orbx_handle* out) {
  if (!p || !out) { set_error("null argument"); return ORBX_ERR_ARG; }
  if (p->nlevels < 1 || p->nlevels > kMaxLevels || p->nfeatures < 1 || !(p->scale_factor > 1.0f) ||
      p->ini_th_fast < 1 || p->min_th_fast < 1 || p->ini_th_fast > 254 || p->min_th_fast > 254) {
    set_error("bad extractor parameters");
    return ORBX_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || p->device < 0 || p->device >= ndev) {
    set_error("no CUDA device (this library has no CPU fallback)");
    return ORBX_ERR_CUDA;
  }
  orbx_extractor* h = new orbx_extractor();
  h->p = *p;
  const int nl = p->nlevels;
  // ---- ORBextractor.cpp:419-451 (scaleFactor is stored as double, ORBextractor.h:98) ----------------
  const double sf = (double)p->scale_factor;
  h->scale.assign(nl, 1.f); h->invScale.assign(nl, 1.f);
  for (int i = 1; i < nl; ++i) h->scale[i] = (float)(h->scale[i - 1] * sf);
  for (int i = 0; i < nl; ++i) h->invScale[i] = 1.0f / h->scale[i];
  h->nfeat.assign(nl, 0);
  const float factor = (float)(1.0f / sf);
  float want = (float)(p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl)));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) {
    h->nfeat[l] = cv_round_f(want);
    sum += h->nfeat[l];
    want *= factor;
  }
  h->nfeat[nl - 1] = std::max(p->nfeatures - sum, 0);
  // ---- ORBextractor.cpp:457-475: patch row extents ---------------------------------------------------
  {
    const int hp = 15;
    const int vmax = (int)floor(hp * sqrt(2.f) / 2 + 1), vmin = (int)ceil(hp * sqrt(2.f) / 2);
    const double hp2 = hp * hp;
    for (int v = 0; v <= vmax; ++v) h->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (int v = hp, v0 = 0; v >= vmin; --v) {
      while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
      h->umax[v] = v0;
      ++v0;
    }
  }
  h->maxKp = 0;
  for (int l = 0; l < nl; ++l) h->maxKp += std::max(h->nfeat[l] + 3, 4 * kMaxRoots);
  if (cudaSetDevice(p->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaSetDevice/cudaStreamCreate failed");
    delete h;
    return ORBX_ERR_CUDA;
  }
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      h->encode = (EncodeTiledFn)fn;
    if (cudaMalloc(&h->d_maps, sizeof(CUtensorMap) * 5 * kMaxLevels) != cudaSuccess) h->d_maps = nullptr;
    h->tma.map = h->d_maps;
  }
  float2 pat[512];
  // transposed for coalesced warp reads: entry [k][lane] = point k (0..15) of descriptor byte `lane`
  for (int lane = 0; lane < 32; ++lane)
    for (int k = 0; k < 16; ++k) {
      const int i = lane * 16 + k;
      pat[k * 32 + lane] = make_float2((float)orb_bit_pattern_31[2 * i], (float)orb_bit_pattern_31[2 * i + 1]);
    }
  if (cudaMalloc(&h->d_pattern, sizeof(pat)) != cudaSuccess || cudaMalloc(&h->d_umax, sizeof(int) * 16) != cudaSuccess ||
      cudaMemcpy(h->d_pattern, pat, sizeof(pat), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->d_umax, h->umax, sizeof(int) * 16, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("device table upload failed");
    delete h;
    return ORBX_ERR_CUDA;
  }
  *out = h;
  return ORBX_OK;
}

int orbx_destroy(orbx_handle h) {
  if (!h) return ORBX_OK;
  cudaSetDevice(h->p.device);
  cudaDeviceSynchronize();
  for (orbx_frame* f : h->framesAll) free_frame_block(f);      // frames die with their extractor (documented in orb_b200.h)
  h->framesAll.clear(); h->framePool.clear();
  if (h->lane2) orbx_destroy(h->lane2);
  for (int k = 0; k < 2; ++k) { if (h->laneX[k]) orbx_destroy(h->laneX[k]); if (h->laneJoinX[k]) cudaEventDestroy(h->laneJoinX[k]); }
  if (h->laneFork) cudaEventDestroy(h->laneFork);
  if (h->laneJoin) cudaEventDestroy(h->laneJoin);
  free_workspace(h);
  cudaFree(h->d_maps);
  cudaFree(h->d_pattern); cudaFree(h->d_umax); cudaFree(h->d_in); cudaFree(h->d_kps); cudaFree(h->d_desc);
  cudaFree(h->d_counts);
  cudaFree(h->d_one); if (h->h_one) cudaFreeHost(h->h_one);
  if (h->oneGraph) cudaGraphExecDestroy(h->oneGraph);
  cudaFree(h->d_midx); cudaFree(h->d_md1); cudaFree(h->d_md2); cudaFree(h->d_mok); cudaFree(h->d_qf);
  cudaFree(h->d_qf2);
  if (h->auxStream) cudaStreamDestroy(h->auxStream);
  if (h->evIn) cudaEventDestroy(h->evIn);
  if (h->evPyr) cudaEventDestroy(h->evPyr);
  if (h->evAux) cudaEventDestroy(h->evAux);
  if (h->copyStream) cudaStreamDestroy(h->copyStream);
  if (h->backStream) cudaStreamDestroy(h->backStream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return ORBX_OK;
}

int orbx_get_levels(orbx_handle h, int* nlevels) {
  if (check_handle(h) || !nlevels) return ORBX_ERR_ARG;
  *nlevels = h->p.nlevels;
  return ORBX_OK;
}
int orbx_scale_factors(orbx_handle h, float* out, int n) {
  if (check_handle(h) || !out || n < h->p.nlevels) return ORBX_ERR_ARG;
  for (int i = 0; i < h->p.nlevels; ++i) out[i] = h->scale[i];
  return ORBX_OK;
}
int orbx_inv_scale_factors(orbx_handle h, float* out, int n) {
  if (check_handle(h) || !out || n < h->p.nlevels) return ORBX_ERR_ARG;
  for (int i = 0; i < h->p.nlevels; ++i) out[i] = h->invScale[i];
  return ORBX_OK;
}
int orbx_features_per_level(orbx_handle h, int* out, int n) {
  if (check_handle(h) || !out || n < h->p.nlevels) return ORBX_ERR_ARG;
  for (int i = 0; i < h->p.nlevels; ++i) out[i] = h->nfeat[i];
  return ORBX_OK;
}
int orbx_max_keypoints(orbx_handle h, int* cap) {
  if (check_handle(h) || !cap) return ORBX_ERR_ARG;
  *cap = h->maxKp;
  return ORBX_OK;
}
int orbx_launch_count(orbx_handle h, long long* n) {
  if (check_handle(h) || !n) return ORBX_ERR_ARG;
  *n = h->launches + (h->lane2 ? h->lane2->launches : 0) + (h->laneX[0] ? h->laneX[0]->launches : 0) + (h->laneX[1] ? h->laneX[1]->launches : 0);
  return ORBX_OK;
}

struct MatchOut { int th; float ratio; int32_t* idx; int32_t* d1; int32_t* d2; uint8_t* ok; };
static int ensure_pair_index(orbx_extractor* h, int nframes) {       // d_qf = 0, 1, 2, ... (frame index of pair p / of its train frame p + 1)
  if (h->qf_frames >= (size_t)nframes + 1) return ORBX_OK;
  cudaFree(h->d_qf2); h->d_qf2 = nullptr; h->qf_frames = 0;
  ORBX_CUDA(cudaMalloc(&h->d_qf2, sizeof(int32_t) * ((size_t)nframes + 1)));
  std::vector<int32_t> iota((size_t)nframes + 1);
  for (int i = 0; i <= nframes; ++i) iota[i] = i;
  ORBX_CUDA(cudaMemcpy(h->d_qf2, iota.data(), sizeof(int32_t) * ((size_t)nframes + 1), cudaMemcpyHostToDevice));
  h->qf_frames = (size_t)nframes + 1;
  return ORBX_OK;
}
// pairs [p0, p1) (pair p: queries = frame p, train = frame p + 1) on `st`
static int launch_pairs(orbx_extractor* h, const uint8_t* d_desc, const int32_t* d_counts, int cap, int p0, int p1, const MatchOut& M,
                        cudaStream_t st) {
  if (p1 <= p0) return ORBX_OK;      // (counted by hamm_launch_count)
  return hamm_knn2_pairs_device(d_desc, d_counts, cap, h->d_qf2 + p0, h->d_qf2 + p0 + 1, p1 - p0, M.th, M.ratio, M.idx + (size_t)p0 * cap,
                                M.d1 + (size_t)p0 * cap, M.d2 + (size_t)p0 * cap, M.ok + (size_t)p0 * cap, st);
}
static int extract_batch_device_impl(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height, size_t row_stride,
                                     size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts,
                                     void* stream, const MatchOut* M);

int orbx_extract_batch_device(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height, size_t row_stride,
                              size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts,
                              void* stream) {
  return extract_batch_device_impl(h, d_imgs, nframes, w, height, row_stride, frame_stride, d_kps, d_desc, cap, d_counts, stream, nullptr);
}

int orbx_extract_match_batch_device(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height, size_t row_stride,
                                    size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, int th,
                                    float ratio, int32_t* d_m_idx, int32_t* d_m_d1, int32_t* d_m_d2, uint8_t* d_m_ok, void* stream) {
  if (check_handle(h)) return ORBX_ERR_ARG;
  if (nframes > 1 && (!d_m_idx || !d_m_d1 || !d_m_d2 || !d_m_ok)) { set_error("bad argument"); return ORBX_ERR_ARG; }
  const MatchOut M{th, ratio, d_m_idx, d_m_d1, d_m_d2, d_m_ok};
  return extract_batch_device_impl(h, d_imgs, nframes, w, height, row_stride, frame_stride, d_kps, d_desc, cap, d_counts, stream, &M);
}

static int extract_batch_device_impl(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height, size_t row_stride,
                                     size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts,
                                     void* stream, const MatchOut* M) {
  if (check_handle(h)) return ORBX_ERR_ARG;
  if (!d_imgs || !d_kps || !d_desc || !d_counts || nframes < 0 || w <= 0 || height <= 0 || cap <= 0 ||
      row_stride < (size_t)w || frame_stride < row_stride * (size_t)(height - 1) + w) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  if (nframes == 0) return ORBX_OK;
  ORBX_CUDA(cudaSetDevice(h->p.device));
  if (M) { if (int rcq = ensure_pair_index(h, nframes)) return rcq; }
  const int want_lanes = env_int("ORBX_LANES", 2);
  // Equal chunks, an even number of them when two lanes run: a batch that is a little larger than a multiple of the chunk
  // size (a rank's block + its replicated boundary frame in the strong-scaling split: 513, 1025, 2049 frames) must not leave
  // one lane with a one-frame chunk, and a batch of at most one chunk (512 frames per GPU at N = 8) still gets both lanes.
  int chunk = std::min(nframes, resident_chunk(w, height));
  const int NL = std::max(1, std::min(want_lanes, 4));
  if (NL >= 2 && nframes >= NL * kLaneMinFrames) {
    const int groups = (nframes + NL * chunk - 1) / (NL * chunk);          // chunk groups needed at the configured chunk size
    chunk = (nframes + NL * groups - 1) / (NL * groups);
  }
  int rc = configure(h, w, height, chunk);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunks = (nframes + chunk - 1) / chunk;
  if (NL >= 2 && nchunks >= 2) {
    // lanes: lane 0 = this handle on the caller's stream, lanes 1.. = siblings (own workspace, own stream); lane k takes the
    // chunks [k * cpl, (k + 1) * cpl)
    const int nl = std::min(NL, nchunks), cpl = (nchunks + nl - 1) / nl;
    orbx_extractor* lane[4] = {h, nullptr, nullptr, nullptr};
    for (int k = 1; k < nl; ++k) {
      orbx_extractor*& slot = k == 1 ? h->lane2 : h->laneX[k - 2];
      if (!slot) {
        orbx_handle sib = nullptr;
        rc = orbx_create(&h->p, &sib);
        if (rc) return rc;
        slot = sib;
      }
      lane[k] = slot;
      rc = configure(slot, w, height, chunk);
      if (rc) return rc;
    }
    if (!h->laneFork) ORBX_CUDA(cudaEventCreateWithFlags(&h->laneFork, cudaEventDisableTiming));
    if (!h->laneJoin) ORBX_CUDA(cudaEventCreateWithFlags(&h->laneJoin, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) if (!h->laneJoinX[k]) ORBX_CUDA(cudaEventCreateWithFlags(&h->laneJoinX[k], cudaEventDisableTiming));
    ORBX_CUDA(cudaEventRecord(h->laneFork, st));
    int first[4], count[4];
    for (int k = 0; k < nl; ++k) {
      first[k] = std::min(k * cpl * chunk, nframes);
      count[k] = std::min((k + 1) * cpl * chunk, nframes) - first[k];
      lane[k]->map0_want = count[k];
      if (k) ORBX_CUDA(cudaStreamWaitEvent(lane[k]->stream, h->laneFork, 0));
    }
    for (int c = 0; c < cpl; ++c)                 // submissions alternate so that every stream has work from the start
      for (int k = 0; k < nl; ++k) {
        const int f0 = c * chunk;
        if (f0 >= count[k]) continue;
        const int nn = std::min(chunk, count[k] - f0);
        rc = run_chunk(lane[k], d_imgs + (size_t)first[k] * frame_stride, row_stride, frame_stride, f0, nn,
                       d_kps + (size_t)first[k] * cap, d_desc + (size_t)first[k] * cap * 32, cap, d_counts + first[k], k ? lane[k]->stream : st);
        if (rc) return rc;
        // the pairs whose train frame is in this chunk and whose query frame is in this lane: [first + f0 - 1, first + f0 + nn - 1),
        // without the pair that straddles the lane boundary (its query frame belongs to the previous lane: after the join)
        if (M) {
          const int g0 = first[k] + f0;
          rc = launch_pairs(h, d_desc, d_counts, cap, std::max(g0 - 1, first[k]), g0 + nn - 1, *M, k ? lane[k]->stream : st);
          if (rc) return rc;
        }
      }
    for (int k = 1; k < nl; ++k) {
      cudaEvent_t ev = k == 1 ? h->laneJoin : h->laneJoinX[k - 2];
      ORBX_CUDA(cudaEventRecord(ev, lane[k]->stream));
      ORBX_CUDA(cudaStreamWaitEvent(st, ev, 0));
    }
    if (M)
      for (int k = 1; k < nl; ++k) {                // the nl - 1 pairs across the lane boundaries
        rc = launch_pairs(h, d_desc, d_counts, cap, first[k] - 1, first[k], *M, st);
        if (rc) return rc;
      }
    return ORBX_OK;
  }
  h->map0_want = nframes;
  for (int f0 = 0; f0 < nframes; f0 += chunk) {
    const int n = std::min(chunk, nframes - f0);
    rc = run_chunk(h, d_imgs, row_stride, frame_stride, f0, n, d_kps, d_desc, cap, d_counts, st);
    if (rc) return rc;
    if (M) { rc = launch_pairs(h, d_desc, d_counts, cap, std::max(f0 - 1, 0), f0 + n - 1, *M, st); if (rc) return rc; }
  }
  return ORBX_OK;
}

int orbx_extract_batch(orbx_handle h, const uint8_t* imgs, int nframes, int w, int height, size_t row_stride,
                       size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts) {
  if (check_handle(h)) return ORBX_ERR_ARG;
  if (!imgs || !kps || !desc || !counts || nframes < 0 || w <= 0 || height <= 0 || cap <= 0 || row_stride < (size_t)w) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  if (nframes == 0) return ORBX_OK;
  ORBX_CUDA(cudaSetDevice(h->p.device));
  const size_t pitch = align_up_sz((size_t)w, 64), fbytes = pitch * height;
  if (h->d_in_bytes < fbytes * nframes) {
    cudaFree(h->d_in);
    h->d_in = nullptr; h->d_in_bytes = 0;
    ORBX_CUDA(cudaMalloc(&h->d_in, fbytes * nframes));
    h->d_in_bytes = fbytes * nframes;
  }
  if (nframes == 1) {
    // The drop-in operator() case: results are packed as [count | keypoints | descriptors] in one device block and come
    // back in ONE copy into a pinned mirror (three copies into the caller's pageable arrays cost ~3x the latency); only
    // the valid entries are then copied out on the host.
    const size_t offK = 64, offD = offK + sizeof(orbx_keypoint) * (size_t)cap, total = offD + (size_t)32 * cap;
    if (h->one_cap < cap) {
      cudaFree(h->d_one); if (h->h_one) cudaFreeHost(h->h_one);
      if (h->oneGraph) { cudaGraphExecDestroy(h->oneGraph); h->oneGraph = nullptr; }
      h->d_one = nullptr; h->h_one = nullptr; h->one_cap = 0;
      ORBX_CUDA(cudaMalloc(&h->d_one, total));
      ORBX_CUDA(cudaHostAlloc(&h->h_one, total, cudaHostAllocDefault));
      h->one_cap = cap;
    }
    cudaStream_t st1 = h->stream;
    if (pitch == row_stride) ORBX_CUDA(cudaMemcpyAsync(h->d_in, imgs, fbytes, cudaMemcpyHostToDevice, st1));
    else ORBX_CUDA(cudaMemcpy2DAsync(h->d_in, pitch, imgs, row_stride, w, height, cudaMemcpyHostToDevice, st1));
    // the chain + the copy down as a CUDA graph, captured on the second call of a geometry (see orbx_frame_create)
    static const bool graphsOn = env_int("ORBX_GRAPH", 1) != 0;
    const unsigned long long key = ((unsigned long long)w << 40) ^ ((unsigned long long)height << 20) ^ (unsigned long long)cap ^ (h->geomEpoch << 52);
    const bool steady = graphsOn && !h->graphBroken && h->haveGeom && h->G.W == w && h->G.H == height && h->map0_base == h->d_in &&
                        h->map0_row == pitch && h->map0_frame == fbytes && h->auxStream != nullptr;
    bool replayed = false;
    if (steady && h->oneGraph && h->oneKey == key) {
      if (cudaGraphLaunch(h->oneGraph, st1) == cudaSuccess) { replayed = true; h->launches += h->oneLaunches; }
      else { cudaGetLastError(); cudaGraphExecDestroy(h->oneGraph); h->oneGraph = nullptr; h->graphBroken = true; }
    }
    if (!replayed) {
      const bool capture = steady && cudaStreamBeginCapture(st1, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      const long long l0 = h->launches;
      int rc1 = orbx_extract_batch_device(h, h->d_in, 1, w, height, pitch, fbytes, (orbx_keypoint*)(h->d_one + offK), h->d_one + offD, cap,
                                          (int32_t*)h->d_one, st1);
      cudaError_t ec = rc1 == ORBX_OK ? cudaMemcpyAsync(h->h_one, h->d_one, total, cudaMemcpyDeviceToHost, st1) : cudaSuccess;
      if (capture) {
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        cudaError_t e = cudaStreamEndCapture(st1, &g);
        if (e == cudaSuccess && rc1 == ORBX_OK && ec == cudaSuccess && g) e = cudaGraphInstantiate(&ge, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e == cudaSuccess && rc1 == ORBX_OK && ec == cudaSuccess && ge) {
          if (h->oneGraph) cudaGraphExecDestroy(h->oneGraph);
          h->oneGraph = ge; h->oneKey = key; h->oneLaunches = (int)(h->launches - l0);
          ORBX_CUDA(cudaGraphLaunch(h->oneGraph, st1));
        } else {
          cudaGetLastError();
          if (ge) cudaGraphExecDestroy(ge);
          h->graphBroken = true;                 // stay eager from now on; run this call again outside the capture
          if (rc1 == ORBX_OK) {
            rc1 = orbx_extract_batch_device(h, h->d_in, 1, w, height, pitch, fbytes, (orbx_keypoint*)(h->d_one + offK), h->d_one + offD,
                                            cap, (int32_t*)h->d_one, st1);
            ec = rc1 == ORBX_OK ? cudaMemcpyAsync(h->h_one, h->d_one, total, cudaMemcpyDeviceToHost, st1) : cudaSuccess;
          }
        }
      }
      if (rc1) { cudaStreamSynchronize(st1); return rc1; }
      ORBX_CUDA(ec);
    }
    ORBX_CUDA(cudaStreamSynchronize(st1));
    const int32_t c = *(const int32_t*)h->h_one;
    const int ncopy = std::max(0, std::min((int)c, cap));
    counts[0] = c;
    memcpy(kps, h->h_one + offK, sizeof(orbx_keypoint) * (size_t)ncopy);
    memcpy(desc, h->h_one + offD, (size_t)32 * ncopy);
    return ORBX_OK;
  }
  if (h->d_out_frames < (size_t)nframes || h->d_out_cap < cap) {
    cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_counts);
    h->d_kps = nullptr; h->d_desc = nullptr; h->d_counts = nullptr; h->d_out_frames = 0;
    ORBX_CUDA(cudaMalloc(&h->d_kps, sizeof(orbx_keypoint) * (size_t)nframes * cap));
    ORBX_CUDA(cudaMalloc(&h->d_desc, (size_t)32 * nframes * cap));
    ORBX_CUDA(cudaMalloc(&h->d_counts, sizeof(int32_t) * nframes));
    h->d_out_frames = nframes; h->d_out_cap = cap;
  }
  cudaStream_t st = h->stream;
  if (frame_stride == row_stride * (size_t)height && pitch == row_stride) {
    ORBX_CUDA(cudaMemcpyAsync(h->d_in, imgs, fbytes * nframes, cudaMemcpyHostToDevice, st));
  } else if (frame_stride == row_stride * (size_t)height) {
    ORBX_CUDA(cudaMemcpy2DAsync(h->d_in, pitch, imgs, row_stride, w, (size_t)height * nframes, cudaMemcpyHostToDevice, st));
  } else {
    for (int f = 0; f < nframes; ++f)
      ORBX_CUDA(cudaMemcpy2DAsync(h->d_in + f * fbytes, pitch, imgs + f * frame_stride, row_stride, w, height,
                                  cudaMemcpyHostToDevice, st));
  }
  int rc = orbx_extract_batch_device(h, h->d_in, nframes, w, height, pitch, fbytes, h->d_kps, h->d_desc, cap, h->d_counts, st);
  if (rc) return rc;
  ORBX_CUDA(cudaMemcpyAsync(counts, h->d_counts, sizeof(int32_t) * nframes, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaMemcpyAsync(kps, h->d_kps, sizeof(orbx_keypoint) * (size_t)nframes * cap, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaMemcpyAsync(desc, h->d_desc, (size_t)32 * nframes * cap, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  return ORBX_OK;
}

int orbx_extract(orbx_handle h, const uint8_t* img, int w, int height, size_t stride, orbx_keypoint* kps, uint8_t* desc,
                 int cap, int* n) {
  if (check_handle(h)) return ORBX_ERR_ARG;
  if (!img || w <= 0 || height <= 0) return ORBX_OK;   // `if(_image.empty()) return;` (ORBextractor.cpp:1054)
  if (!n) { set_error("null count pointer"); return ORBX_ERR_ARG; }
  int32_t c = 0;
  int rc = orbx_extract_batch(h, img, 1, w, height, stride, stride * (size_t)height, kps, desc, cap, &c);
  if (rc) return rc;
  *n = c;
  if (c > cap) { set_error("keypoint buffer too small"); return ORBX_ERR_CAPACITY; }
  return ORBX_OK;
}


// Per-stage device time of one batch (CUDA events on `stream` around every stage of every chunk; ms summed over
// chunks).  ms[0..4] = pyramid, FAST, quadtree, blur, orientation+descriptors; ms[5] = whole batch.
int orbx_profile_stages(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height, size_t row_stride,
                        size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts,
                        void* stream, float* ms) {
  if (check_handle(h)) return ORBX_ERR_ARG;
  if (!d_imgs || !d_kps || !d_desc || !d_counts || !ms || nframes <= 0 || w <= 0 || height <= 0 || cap <= 0) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  ORBX_CUDA(cudaSetDevice(h->p.device));
  const int chunk = std::min(nframes, resident_chunk(w, height));
  int rc = configure(h, w, height, chunk);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  h->map0_want = nframes;
  const int nchunks = (nframes + chunk - 1) / chunk;
  std::vector<cudaEvent_t> ev((size_t)nchunks * 6);
  for (auto& e : ev) ORBX_CUDA(cudaEventCreate(&e));
  for (int c = 0; c < nchunks; ++c) {
    const int f0 = c * chunk, n = std::min(chunk, nframes - f0);
    rc = run_chunk(h, d_imgs, row_stride, frame_stride, f0, n, d_kps, d_desc, cap, d_counts, st, &ev[(size_t)c * 6]);
    if (rc) break;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  for (int k = 0; k < 6; ++k) ms[k] = 0.f;
  if (rc == ORBX_OK && e == cudaSuccess) {
    for (int c = 0; c < nchunks; ++c)
      for (int k = 0; k < 5; ++k) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[(size_t)c * 6 + k], ev[(size_t)c * 6 + k + 1]);
        ms[k] += t;
      }
    cudaEventElapsedTime(&ms[5], ev[0], ev[(size_t)(nchunks - 1) * 6 + 5]);
  }
  for (auto& x : ev) cudaEventDestroy(x);
  if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return ORBX_ERR_CUDA; }
  return rc;
}

// Host-buffer batch: extraction of every frame + frame-to-frame Hamming top-2 (pair p: queries = frame p, train =
// frame p+1), BASELINE config 2.  Host->device copies of chunk c+1 overlap the kernels of chunk c; results stream back
// on a third stream.  imgs should be pinned for the copies to be asynchronous.  m_* are [(nframes-1) * cap].
int orbx_extract_match_batch(orbx_handle h, const uint8_t* imgs, int nframes, int w, int height, size_t row_stride,
                             size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts, int th,
                             float ratio, int32_t* m_idx, int32_t* m_d1, int32_t* m_d2, uint8_t* m_ok) {
  if (check_handle(h)) return ORBX_ERR_ARG;
  if (!imgs || !kps || !desc || !counts || !m_idx || !m_d1 || !m_d2 || !m_ok || nframes <= 0 || w <= 0 || height <= 0 ||
      cap <= 0 || row_stride < (size_t)w) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  ORBX_CUDA(cudaSetDevice(h->p.device));
  const std::vector<std::pair<int, int>> sched = host_schedule(nframes, std::min(nframes, host_chunk(w, height)));
  int rc = configure(h, w, height, std::min(nframes, host_chunk(w, height)));
  if (rc) return rc;
  const size_t pitch = align_up_sz((size_t)w, 64), fbytes = pitch * height;
  const int npairs = nframes - 1;
  if (h->d_in_bytes < fbytes * nframes) {
    cudaFree(h->d_in); h->d_in = nullptr; h->d_in_bytes = 0;
    ORBX_CUDA(cudaMalloc(&h->d_in, fbytes * nframes));
    h->d_in_bytes = fbytes * nframes;
  }
  if (h->d_out_frames < (size_t)nframes || h->d_out_cap < cap) {
    cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_counts);
    h->d_kps = nullptr; h->d_desc = nullptr; h->d_counts = nullptr; h->d_out_frames = 0;
    ORBX_CUDA(cudaMalloc(&h->d_kps, sizeof(orbx_keypoint) * (size_t)nframes * cap));
    ORBX_CUDA(cudaMalloc(&h->d_desc, (size_t)32 * nframes * cap));
    ORBX_CUDA(cudaMalloc(&h->d_counts, sizeof(int32_t) * nframes));
    h->d_out_frames = nframes; h->d_out_cap = cap;
  }
  if (h->m_frames < (size_t)nframes || h->m_cap < cap) {
    cudaFree(h->d_midx); cudaFree(h->d_md1); cudaFree(h->d_md2); cudaFree(h->d_mok); cudaFree(h->d_qf);
    h->d_midx = h->d_md1 = h->d_md2 = nullptr; h->d_mok = nullptr; h->d_qf = nullptr; h->m_frames = 0;
    const size_t np = std::max(npairs, 1);
    ORBX_CUDA(cudaMalloc(&h->d_midx, sizeof(int32_t) * np * cap));
    ORBX_CUDA(cudaMalloc(&h->d_md1, sizeof(int32_t) * np * cap));
    ORBX_CUDA(cudaMalloc(&h->d_md2, sizeof(int32_t) * np * cap));
    ORBX_CUDA(cudaMalloc(&h->d_mok, np * cap));
    ORBX_CUDA(cudaMalloc(&h->d_qf, sizeof(int32_t) * (nframes + 1)));
    std::vector<int32_t> iota(nframes + 1);
    for (int i = 0; i <= nframes; ++i) iota[i] = i;
    ORBX_CUDA(cudaMemcpy(h->d_qf, iota.data(), sizeof(int32_t) * (nframes + 1), cudaMemcpyHostToDevice));
    h->m_frames = nframes; h->m_cap = cap;
  }
  if (!h->copyStream) {
    ORBX_CUDA(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
    ORBX_CUDA(cudaStreamCreateWithFlags(&h->backStream, cudaStreamNonBlocking));
  }
  cudaStream_t sc = h->copyStream, sb = h->backStream;
  h->map0_want = nframes;
  const int nchunks = (int)sched.size();
  // Two compute lanes (this handle and its sibling, each with its own workspace and stream) take the chunks alternately: one
  // lane alone processes 256-frame chunks a little slower than the link delivers them, so the copies would wait for the kernels;
  // with two lanes the kernel tails of one chunk fill under the next chunk's kernels and the step is bound by the upload.
  // The pairs of chunk c need the last frame of chunk c-1, extracted on the other lane: its `ext` event orders them.
  const bool two = env_int("ORBX_HOST_LANES", 2) >= 2 && nchunks >= 2;
  orbx_extractor* g = h;
  if (two) {
    if (!h->lane2) {
      orbx_handle sib = nullptr;
      rc = orbx_create(&h->p, &sib);
      if (rc) return rc;
      h->lane2 = sib;
    }
    g = h->lane2;
    rc = configure(g, w, height, std::min(nframes, host_chunk(w, height)));
    if (rc) return rc;
    g->map0_want = nframes;
  }
  std::vector<cudaEvent_t> up(nchunks), ext(nchunks), done(nchunks);
  for (int c = 0; c < nchunks; ++c) {
    ORBX_CUDA(cudaEventCreateWithFlags(&up[c], cudaEventDisableTiming));
    ORBX_CUDA(cudaEventCreateWithFlags(&ext[c], cudaEventDisableTiming));
    ORBX_CUDA(cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming));
  }
  const bool dense = frame_stride == row_stride * (size_t)height;
  const int dbg = env_int("ORBX_E2E_DEBUG", 0);      // diagnosis only: 1 = no kernels, 2 = no downloads, 3 = neither
  for (int c = 0; c < nchunks && rc == ORBX_OK; ++c) {
    const int f0 = sched[c].first, n = sched[c].second;
    orbx_extractor* x = (two && (c & 1)) ? g : h;
    cudaStream_t sk = x->stream;
    if (dense && pitch == row_stride) {          // one contiguous block: a plain 1-D copy (no per-row DMA descriptors)
      cudaMemcpyAsync(h->d_in + f0 * fbytes, imgs + f0 * frame_stride, fbytes * n, cudaMemcpyHostToDevice, sc);
    } else if (dense) {
      cudaMemcpy2DAsync(h->d_in + f0 * fbytes, pitch, imgs + f0 * frame_stride, row_stride, w, (size_t)height * n,
                        cudaMemcpyHostToDevice, sc);
    } else {
      for (int f = f0; f < f0 + n; ++f)
        cudaMemcpy2DAsync(h->d_in + f * fbytes, pitch, imgs + f * frame_stride, row_stride, w, height, cudaMemcpyHostToDevice, sc);
    }
    cudaEventRecord(up[c], sc);
    cudaStreamWaitEvent(sk, up[c], 0);
    if (!(dbg & 1)) rc = run_chunk(x, h->d_in, pitch, fbytes, f0, n, h->d_kps, h->d_desc, cap, h->d_counts, sk);
    if (rc) break;
    cudaEventRecord(ext[c], sk);
    // pairs whose second frame is now available: p in [max(f0-1,0), f0+n-1)
    const int p0 = std::max(f0 - 1, 0), p1 = f0 + n - 1;
    if (p1 > p0 && !(dbg & 1)) {
      if (two && c > 0) cudaStreamWaitEvent(sk, ext[c - 1], 0);
      rc = hamm_knn2_pairs_device(h->d_desc, h->d_counts, cap, h->d_qf + p0, h->d_qf + p0 + 1, p1 - p0, th, ratio,
                                  h->d_midx + (size_t)p0 * cap, h->d_md1 + (size_t)p0 * cap, h->d_md2 + (size_t)p0 * cap,
                                  h->d_mok + (size_t)p0 * cap, sk);
    }
    cudaEventRecord(done[c], sk);
    cudaStreamWaitEvent(sb, done[c], 0);
    if (dbg & 2) continue;
    cudaMemcpyAsync(counts + f0, h->d_counts + f0, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, sb);
    cudaMemcpyAsync(kps + (size_t)f0 * cap, h->d_kps + (size_t)f0 * cap, sizeof(orbx_keypoint) * (size_t)n * cap, cudaMemcpyDeviceToHost, sb);
    cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, h->d_desc + (size_t)f0 * cap * 32, (size_t)32 * n * cap, cudaMemcpyDeviceToHost, sb);
    if (p1 > p0) {
      const size_t o = (size_t)p0 * cap, cnt = (size_t)(p1 - p0) * cap;
      cudaMemcpyAsync(m_idx + o, h->d_midx + o, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, sb);
      cudaMemcpyAsync(m_d1 + o, h->d_md1 + o, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, sb);
      cudaMemcpyAsync(m_d2 + o, h->d_md2 + o, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, sb);
      cudaMemcpyAsync(m_ok + o, h->d_mok + o, cnt, cudaMemcpyDeviceToHost, sb);
    }
  }
  cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(sb), e3 = cudaStreamSynchronize(sc);
  if (two) { const cudaError_t e4 = cudaStreamSynchronize(g->stream); if (e1 == cudaSuccess) e1 = e4; }
  for (int c = 0; c < nchunks; ++c) { cudaEventDestroy(up[c]); cudaEventDestroy(ext[c]); cudaEventDestroy(done[c]); }
  if (rc) return rc;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    set_error(cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return ORBX_ERR_CUDA;
  }
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

// ---- device-resident frame: Frame::Frame (frame.cpp:22-32) in one call -----------------------------------------------------
// image up -> the extractor's kernel chain -> frame_finish_kernel (undistortKeyPoints, findDepth, assignFeaturesToGrid, the
// searches' per-feature records) -> ONE packed copy down.  Everything is enqueued on the handle's stream before the single
// synchronisation; the depth image (when given) goes up on the copy stream under the extraction kernels.
int orbx_frame_create(orbx_handle h, const orbx_camera* cam, const uint8_t* img, int w, int height, size_t stride, const float* depth,
                      size_t depth_row_stride, orbx_frame_t* out, int* n) {
  if (check_handle(h)) return ORBX_ERR_ARG;
  if (!cam || !img || !out || w <= 0 || height <= 0 || stride < (size_t)w ||
      (depth && (depth_row_stride < sizeof(float) * (size_t)w || depth_row_stride % sizeof(float)))) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  *out = nullptr;
  ORBX_CUDA(cudaSetDevice(h->p.device));
  orbx_frame* f = nullptr;
  if (!h->framePool.empty()) { f = h->framePool.back(); h->framePool.pop_back(); }
  else if (int rc = alloc_frame_block(h, &f)) return rc;
  auto fail = [&](int rc) { h->framePool.push_back(f); return rc; };
  const size_t pitch = align_up_sz((size_t)w, 64), fbytes = pitch * height;
  if (h->d_in_bytes < fbytes) {
    cudaFree(h->d_in); h->d_in = nullptr; h->d_in_bytes = 0;
    if (cudaMalloc(&h->d_in, fbytes) != cudaSuccess) { set_error("input staging allocation failed"); return fail(ORBX_ERR_CUDA); }
    h->d_in_bytes = fbytes;
  }
  cudaStream_t st = h->stream;
  if (f->patchPending) { cudaStreamWaitEvent(st, f->patched, 0); f->patchPending = false; }
  static const bool phaseTiming = getenv("ORBX_CREATE_TIMING") != nullptr;   // debugging aid: where the host time of one call goes
  const auto tc0 = std::chrono::steady_clock::now();
  cudaError_t e = pitch == stride ? cudaMemcpyAsync(h->d_in, img, fbytes, cudaMemcpyHostToDevice, st)
                                  : cudaMemcpy2DAsync(h->d_in, pitch, img, stride, w, height, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return fail(ORBX_ERR_CUDA); }
  // findDepth (frame.cpp:108-133) needs the depth image only at the ~1000 keypoint positions, which exist after the extraction:
  // uploading the whole image (1.2 MB, pageable in the reference: 55 us) would cost more than the rest of the constructor's
  // copies together.  The kernel therefore leaves depth_/uRight_ at -1; after the one packed copy down the host samples the
  // image at the ORIGINAL keypoints with the reference's float->int truncation, computes uRight_ = unKp.x - bf/d in IEEE float
  // (one division, one subtraction: no contraction possible) and sends the 4 KB of uRight_ back for the searches.
  // The device chain of one frame (14 launches on two streams, their events, the finish kernel and the copy down) is the same
  // work descriptor for every frame built in this block with this camera: it is captured ONCE into a CUDA graph and replayed,
  // which takes the per-launch cost off the host and off the GPU front end.  Captured on the second frame of a block (the first
  // call runs eagerly: it may still configure the workspace and describe the staging buffer, which a capture cannot contain).
  static const bool graphsOn = env_int("ORBX_GRAPH", 1) != 0;
  unsigned long long key = 1469598103934665603ull;
  {
    auto mix = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; ++i) { key ^= b[i]; key *= 1099511628211ull; } };
    mix(cam, sizeof(*cam)); mix(&w, sizeof(w)); mix(&height, sizeof(height)); mix(&h->geomEpoch, sizeof(h->geomEpoch));
  }
  const bool steady = graphsOn && !h->graphBroken && h->haveGeom && h->G.W == w && h->G.H == height && h->map0_base == h->d_in &&
                      h->map0_row == pitch && h->map0_frame == fbytes && h->auxStream != nullptr;
  int rc = ORBX_OK;
  bool replayed = false;
  if (steady && f->graph && f->graphKey == key) {
    e = cudaGraphLaunch(f->graph, st);
    if (e == cudaSuccess) { replayed = true; h->launches += f->graphLaunches; }
    else { cudaGetLastError(); cudaGraphExecDestroy(f->graph); f->graph = nullptr; h->graphBroken = true; }
  }
  if (!replayed) {
    const bool capture = steady && cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    const long long l0 = h->launches;
    rc = orbx_extract_batch_device(h, h->d_in, 1, w, height, pitch, fbytes, f->d_kps, f->d_desc, f->cap, f->d_count, st);
    if (rc == ORBX_OK)
      rc = frame_finish_launch(cam, f->d_kps, f->d_count, 1, f->cap, nullptr, w, height, sizeof(float) * (size_t)w, 0,
                               f->d_unkps, f->d_uright, f->d_depth, f->d_cellStart, f->d_ids, f->d_feat, f->d_angle, st);
    if (rc == ORBX_OK) {
      h->launches += 1;
      if (cudaMemcpyAsync(f->h_mirror, f->d_block, f->packed_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = ORBX_ERR_CUDA;
    }
    if (capture) {
      cudaGraph_t g = nullptr;
      cudaGraphExec_t ge = nullptr;
      e = cudaStreamEndCapture(st, &g);
      if (e == cudaSuccess && rc == ORBX_OK && g) e = cudaGraphInstantiate(&ge, g, 0);
      if (g) cudaGraphDestroy(g);
      if (e == cudaSuccess && rc == ORBX_OK && ge) {
        if (f->graph) cudaGraphExecDestroy(f->graph);
        f->graph = ge; f->graphKey = key; f->graphLaunches = (int)(h->launches - l0);
        e = cudaGraphLaunch(f->graph, st);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = ORBX_ERR_CUDA; }
      } else {
        // this driver cannot capture the chain (or a launch failed inside it): run eagerly from now on
        cudaGetLastError();
        if (ge) cudaGraphExecDestroy(ge);
        h->graphBroken = true;
        if (rc == ORBX_OK) {
          rc = orbx_extract_batch_device(h, h->d_in, 1, w, height, pitch, fbytes, f->d_kps, f->d_desc, f->cap, f->d_count, st);
          if (rc == ORBX_OK)
            rc = frame_finish_launch(cam, f->d_kps, f->d_count, 1, f->cap, nullptr, w, height, sizeof(float) * (size_t)w, 0,
                                     f->d_unkps, f->d_uright, f->d_depth, f->d_cellStart, f->d_ids, f->d_feat, f->d_angle, st);
          if (rc == ORBX_OK && cudaMemcpyAsync(f->h_mirror, f->d_block, f->packed_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = ORBX_ERR_CUDA;
        }
      }
    }
  }
  if (rc == ORBX_OK) {
    e = cudaSuccess;
    const auto tc1 = std::chrono::steady_clock::now();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (phaseTiming) {
      const auto tc2 = std::chrono::steady_clock::now();
      fprintf(stderr, "[frame_create] enqueue %.1f us, wait %.1f us\n", std::chrono::duration<double, std::micro>(tc1 - tc0).count(),
              std::chrono::duration<double, std::micro>(tc2 - tc1).count());
    }
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = ORBX_ERR_CUDA; }
  } else {
    cudaStreamSynchronize(st);
  }
  if (rc) return fail(rc);
  const int32_t c = *(const int32_t*)f->h_mirror;
  f->n = std::max(0, std::min((int)c, f->cap));
  if (depth && f->n > 0) {
    uint8_t* m = f->h_mirror;
    const orbx_keypoint* hk = (const orbx_keypoint*)(m + ((uint8_t*)f->d_kps - f->d_block));
    const orbx_keypoint* hu = (const orbx_keypoint*)(m + ((uint8_t*)f->d_unkps - f->d_block));
    float* hur = (float*)(m + ((uint8_t*)f->d_uright - f->d_block));
    float* hdp = (float*)(m + ((uint8_t*)f->d_depth - f->d_block));
    const float bf = cam->bf;
    for (int i = 0; i < f->n; ++i) {
      const int col = std::min(std::max((int)hk[i].x, 0), w - 1), row = std::min(std::max((int)hk[i].y, 0), height - 1);   // at<float>(v,u)
      const volatile float dv = *(const float*)((const char*)depth + (size_t)row * depth_row_stride + (size_t)col * sizeof(float));
      if (dv > 0) {                                                     // frame.cpp:126-130
        volatile float q = bf / dv;
        hdp[i] = dv; hur[i] = hu[i].x - q;
      }
    }
    // the searches run on the legacy default stream: the 4 KB and the patch kernel are ordered in front of them there
    float* hurDev = nullptr;       // the pinned mirror as the device sees it: the patch kernel reads the 4 KB in place
    if (cudaHostGetDevicePointer((void**)&hurDev, hur, 0) == cudaSuccess && hurDev) {
      if (int r2 = frame_patch_uright(hurDev, f->d_uright, f->d_feat, f->n, nullptr)) return fail(r2);
      if (!f->patched) cudaEventCreateWithFlags(&f->patched, cudaEventDisableTiming);
      if (f->patched && cudaEventRecord(f->patched, nullptr) == cudaSuccess) f->patchPending = true;
    } else {
      e = cudaMemcpyAsync(f->d_uright, hur, sizeof(float) * (size_t)f->n, cudaMemcpyHostToDevice, nullptr);
      if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return fail(ORBX_ERR_CUDA); }
      if (int r2 = frame_patch_uright(f->d_uright, f->d_uright, f->d_feat, f->n, nullptr)) return fail(r2);
    }
    h->launches += 1;
  }
  f->xmin = cam->xmin; f->xmax = cam->xmax; f->ymin = cam->ymin; f->ymax = cam->ymax;
  h->framesLive++;
  if (n) *n = f->n;
  *out = f;
  return ORBX_OK;
}

int orbx_frame_size(orbx_frame_t f, int* n) {
  if (!f || !n) { set_error("null argument"); return ORBX_ERR_ARG; }
  *n = f->n;
  return ORBX_OK;
}

// Host copies of the frame's members out of the pinned mirror (no device traffic).  Any pointer may be NULL.
int orbx_frame_get(orbx_frame_t f, orbx_keypoint* kps, uint8_t* desc, orbx_keypoint* unkps, float* uright, float* depth, int cap) {
  if (!f || cap < 0) { set_error("bad argument"); return ORBX_ERR_ARG; }
  if (!f->h_mirror) { set_error("an uploaded frame has no host mirror (the caller holds the originals)"); return ORBX_ERR_ARG; }
  if (cap < f->n) { set_error("keypoint buffer too small"); return ORBX_ERR_CAPACITY; }
  const uint8_t* m = f->h_mirror;
  const size_t n = (size_t)f->n;
  if (kps) memcpy(kps, m + ((uint8_t*)f->d_kps - f->d_block), sizeof(orbx_keypoint) * n);
  if (desc) memcpy(desc, m + (f->d_desc - f->d_block), 32 * n);
  if (unkps) memcpy(unkps, m + ((uint8_t*)f->d_unkps - f->d_block), sizeof(orbx_keypoint) * n);
  if (uright) memcpy(uright, m + ((uint8_t*)f->d_uright - f->d_block), sizeof(float) * n);
  if (depth) memcpy(depth, m + ((uint8_t*)f->d_depth - f->d_block), sizeof(float) * n);
  return ORBX_OK;
}

// The 64x48 grid as the CSR of orbx_grid_build (cell_start[64*48+1], ids[n]) -- Frame::gridKeypoints_ for host code that still
// reads it.
int orbx_frame_grid(orbx_frame_t f, int32_t* cell_start, int32_t* ids, int cap) {
  if (!f || !cell_start || !ids) { set_error("bad argument"); return ORBX_ERR_ARG; }
  if (!f->h_mirror) { set_error("an uploaded frame has no host mirror"); return ORBX_ERR_ARG; }
  const int32_t* cs = (const int32_t*)(f->h_mirror + ((uint8_t*)f->d_cellStart - f->d_block));
  const int total = cs[ORBX_GRID_COLS * ORBX_GRID_ROWS];
  if (cap < total) { set_error("id buffer too small"); return ORBX_ERR_CAPACITY; }
  memcpy(cell_start, cs, sizeof(int32_t) * (ORBX_GRID_COLS * ORBX_GRID_ROWS + 1));
  memcpy(ids, f->h_mirror + ((uint8_t*)f->d_ids - f->d_block), sizeof(int32_t) * (size_t)total);
  return ORBX_OK;
}

int orbx_frame_destroy(orbx_frame_t f) {
  if (!f) return ORBX_OK;
  orbx_extractor* h = f->owner;
  if (!h) {                      // orbx_frame_upload: no extractor, no pool
    cudaSetDevice(f->device);
    free_frame_block(f);
    return ORBX_OK;
  }
  h->framesLive--;
  if (h->framePool.size() < 8) h->framePool.push_back(f);      // the tracking thread keeps 2-3 frames alive; key frames persist
  else {
    cudaSetDevice(f->device);
    h->framesAll.erase(std::remove(h->framesAll.begin(), h->framesAll.end(), f), h->framesAll.end());
    free_frame_block(f);
  }
  return ORBX_OK;
}

// ---- stage taps -------------------------------------------------------------------------------------------
int orbx_debug_level(orbx_handle h, int frame, int level, int blurred, uint8_t* out, int* w, int* height) {
  if (check_handle(h) || !h->haveGeom || level < 0 || level >= h->G.nlevels || frame < 0 || frame >= h->last_frames)
    return ORBX_ERR_ARG;
  const LevelGeom& L = h->G.L[level];
  if (w) *w = L.w;
  if (height) *height = L.h;
  if (!out) return ORBX_OK;
  ORBX_CUDA(cudaSetDevice(h->p.device));
  ORBX_CUDA(cudaDeviceSynchronize());
  const uint8_t* src; size_t pitch;
  if (blurred) { src = h->d_blur + L.blurOff + (size_t)frame * L.h * L.bpitch; pitch = L.bpitch; }
  else if (level == 0) { src = h->last_img0 + (size_t)frame * h->last_frameStride; pitch = h->last_rowStride; }
  else { src = h->d_pyr + L.pyrOff + (size_t)frame * L.h * L.pitch; pitch = L.pitch; }
  ORBX_CUDA(cudaMemcpy2D(out, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
  return ORBX_OK;
}

static int copy_keys(orbx_handle h, const uint32_t* d_src, int count, int32_t* xys, int cap, int* n) {
  std::vector<uint32_t> k(std::max(count, 1));
  if (count > 0) ORBX_CUDA(cudaMemcpy(k.data(), d_src, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost));
  for (int i = 0; i < count && i < cap; ++i) { xys[3 * i] = key_x(k[i]); xys[3 * i + 1] = key_y(k[i]); xys[3 * i + 2] = key_s(k[i]); }
  *n = count;
  return ORBX_OK;
}

int orbx_debug_candidates(orbx_handle h, int frame, int level, int32_t* xys, int cap, int* n) {
  if (check_handle(h) || !h->haveGeom || level < 0 || level >= h->G.nlevels || frame < 0 || frame >= h->last_frames || !n)
    return ORBX_ERR_ARG;
  ORBX_CUDA(cudaSetDevice(h->p.device));
  ORBX_CUDA(cudaDeviceSynchronize());
  int count = 0;
  ORBX_CUDA(cudaMemcpy(&count, h->d_candCount + (size_t)frame * h->G.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  return copy_keys(h, h->d_flatKeys + (size_t)frame * h->G.keysPerFrame + h->G.L[level].keyBase, count, xys, cap, n);
}

int orbx_debug_selected(orbx_handle h, int frame, int level, int32_t* xys, int cap, int* n) {
  if (check_handle(h) || !h->haveGeom || level < 0 || level >= h->G.nlevels || frame < 0 || frame >= h->last_frames || !n)
    return ORBX_ERR_ARG;
  ORBX_CUDA(cudaSetDevice(h->p.device));
  ORBX_CUDA(cudaDeviceSynchronize());
  int count = 0;
  ORBX_CUDA(cudaMemcpy(&count, h->d_selCount + (size_t)frame * h->G.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  return copy_keys(h, h->d_sel + (size_t)frame * h->G.selPerFrame + h->G.L[level].selBase, count, xys, cap, n);
}

}  // extern "C"
