// orb_extract.cuh -- geometry/handle types shared by the extractor kernels and the C ABI.
#pragma once
#include <vector>

#include <cuda.h>

#include "common.cuh"

namespace orbx {

constexpr int kMaxLevels = 16;
constexpr int kEdge = 19;            // EDGE_THRESHOLD        (ORBextractor.cpp:76)
constexpr int kMinBorder = 16;       // EDGE_THRESHOLD - 3    (ORBextractor.cpp:780)
#ifndef ORBX_CELLS_PER_CTA
#define ORBX_CELLS_PER_CTA 2           // -DORBX_CELLS_PER_CTA=1|2|4 builds A/B variants (tools/build_variants.sh)
#endif
constexpr int kCellsPerCta = ORBX_CELLS_PER_CTA;      // FAST cells handled by one CTA (one "slot")
constexpr int kMaxRoots = 16;
constexpr int kFastThreads = 128;
constexpr int kOctMaxThreads = 1024; // quadtree CTA size is chosen per launch (launch_octree): 128 ... 1024

struct LevelGeom {
  int w, h;                // level size
  int regW, regH;          // FAST region (w-32) x (h-32)              (ORBextractor.cpp:780-789)
  int nCols, nRows;        // cells                                    (:791-792)
  int wCell, hCell;        //                                          (:793-794)
  int groups;              // CTAs per cell row
  uint32_t groupsMagic;    // 0xFFFFFFFF / groups + 1: slot / groups == umulhi(slot, groupsMagic) for slot < 2^16
  int slot0, nSlots;       // slots of this level: nRows * groups, ordered (cell row, group)
  int nFeat;               // mnFeaturesPerLevel                       (:439-451)
  int nIni;                // quadtree roots                           (:549)
  float hX;                //                                          (:551)
  int rootX[kMaxRoots + 1];
  int selBase, selCap;     // per-frame offset / capacity of the selected-keypoint array
  int keyBase, keyCap;     // per-frame offset / capacity of the flat candidate array
  int pitch;               // row pitch of this level in the pyramid workspace (levels >= 1)
  int bpitch;              // row pitch in the blurred workspace
  size_t pyrOff;           // byte offset of the [chunk][h][pitch] array of this level
  size_t blurOff;
  int blurTile0;           // first blur tile of this level in the fused blur launch
  int blurTilesX, blurTilesY;
  int bwTile0, bwTilesX, bwTilesY;   // register-blocked blur: CTAs of 64 columns x 208 rows
  float scale;             // mvScaleFactor[level]
  float kpSize;            // (int)(31*scale)                          (:845)
};

struct Geom {
  int nlevels;
  int W, H;
  int iniTh, minTh;
  int totalSlots;
  int slotKeysPerFrame;    // sum of slot capacities
  int keysPerFrame;        // sum of keyCap
  int selPerFrame;         // sum of selCap
  int nodeCap;             // max nodes of any level's quadtree
  int maxSlotsPerLevel;
  int blurTiles;
  int bwTiles;
  int fastTileW, fastTileH;  // smem tile extents of the FAST kernel (pitch is a multiple of 16)
  int fastSurvCap;           // largest slot capacity (NMS survivors of one CTA)
  int fastScW, fastScH;      // score map of the warp-synchronous FAST kernel: cell interiors only (pitch multiple of 16)
  int fastCellPix, fastCellQuads, fastCellSurv;   // per-cell maxima (pixels, 4-pixel quads, NMS survivors), 16-byte rounded
  LevelGeom L[kMaxLevels];
};

// level-0 image set of the current chunk + workspaces
struct Bufs {
  const uint8_t* img0;     // level 0 = the caller's frames
  size_t rowStride0, frameStride0;
  uint8_t* pyr;            // levels >= 1
  uint8_t* blur;           // blurred levels (all)
  const int* slotKeyBase;  // [totalSlots+1]
  int* slotCount;          // [chunk][totalSlots]
  uint32_t* slotKeys;      // [chunk][slotKeysPerFrame]
  uint32_t* flatKeys;      // [chunk][keysPerFrame]
  uint16_t* nodeOf;        // [chunk][keysPerFrame]
  int* candCount;          // [chunk][nlevels]
  uint32_t* sel;           // [chunk][selPerFrame]
  int* selCount;           // [chunk][nlevels]
  const float2* pattern;   // 512 rBRIEF points
  const int* umax;         // 16
  int slotOff;             // first FAST slot / quadtree level of this launch (latency mode launches level 0 and levels >= 1
  int levelOff;            // separately, on two streams); 0 for whole-pyramid launches
  int octKeySmem;          // latency mode: the quadtree CTA keeps up to this many (key, node label) pairs in shared memory
};

// TMA descriptors (cuTensorMapEncodeTiled, rank 3: x bytes, y rows, frame) of every pyramid level, used by the FAST
// kernel to fetch its tile with one cp.async.bulk.tensor instead of a load/store loop.  use[l] == 0: the level cannot be
// described (base or pitch not 16-byte aligned) and the kernel falls back to vector/byte loads.
struct TmaSet {
  const CUtensorMap* map;  // 5*kMaxLevels descriptors in GLOBAL memory (64-byte aligned), written by the host before launch:
                           // [l] FAST tile of level l, [kMaxLevels + l] 48x31 orientation patch of level l,
                           // [2*kMaxLevels + l] 64x37 descriptor patch of BLURRED level l,
                           // [3*kMaxLevels + l] resize source box of destination level l (a box over level l-1),
                           // [4*kMaxLevels + l] 96x214 blur source box of level l
  int use[kMaxLevels];
  int usePatch;            // every level has both patch descriptors -> orient_desc_tma_kernel
  int useBlur;             // every level has a blur source descriptor -> blur_tma_kernel
  int frame0;              // z coordinate of the chunk's first frame in the level-0 maps (levels >= 1 are chunk-local)
};
constexpr int kOdUW = 48, kOdUH = 31;    // orientation patch box: 31 rows of (15 + 31 + pad) bytes, x origin 16-byte aligned
constexpr int kBtBoxW = 96, kBtBoxH = 214; // blur source box: 64 columns + 16-byte aligned halos, 208 rows + 3 + 3
constexpr int kOdBW = 64, kOdBH = 37;    // descriptor patch box: 37 rows of (15 + 37 + pad) bytes

struct ResizeTaps {        // device tables of one level
  int* xofs; short* xa0; short* xa1;
  int* yofs; short* yb0; short* yb1;
  // per destination quad (4 px): {aligned source byte offset, funnel shift, 4 byte-pair selectors, -} and the four
  // packed weight pairs a0 | a1 << 16; quadOk = every quad's taps fit an 8-byte window (scale factor <= 2)
  int4* quad; uint4* xw; int quadOk;
};

// TMA staging of the resize source (resize_tma_kernel): a CTA produces a 128 x 64 destination tile from ONE
// cp.async.bulk.tensor box of the source level; the box origin of every tile column / tile row comes from the tap tables.
#ifndef ORBX_RW_ROWS
#define ORBX_RW_ROWS 16                // destination rows per thread of the resize walks (A/B builds: 8, 32)
#endif
constexpr int kRwRows = ORBX_RW_ROWS;
constexpr int kRzTileW = 128, kRzTileH = 4 * kRwRows, kRzMaxTX = 32, kRzMaxTY = 3072 / kRzTileH;
struct ResizeTma {
  int use;                 // descriptor valid and the boxes fit (else: resize_walk_kernel)
  int rows;                // destination rows per warp: kRwRows, or half of it for small levels (more, shorter CTAs)
  int boxW, boxH;          // bytes x rows, boxW a multiple of 16, both <= 256
  short x0[kRzMaxTX];      // 16-byte aligned source byte offset of tile column tx
  short y0[kRzMaxTY];      // first source row of tile row ty
};

// Fused pyramid (pyramid_fused_kernel): ONE launch builds levels 1..n-1 of every frame of the chunk.  A CTA takes a ticket
// (level-major order), and a tile of level l waits on per-(frame, level, tile row) completion counters of level l-1 instead of
// on a kernel boundary.  PyrLevel: one destination level (device array, uploaded when the geometry or the input map changes).
struct PyrLevel {
  const CUtensorMap* map;  // source box descriptor (a box over level l-1; level 1: over the caller's frames)
  uint8_t* dst; size_t dframe;
  int dw, dh, dpitch;
  int sh;                  // rows of the source level
  int ntx, nty;            // destination tiles of this level (128 columns x 4 * R.rows rows)
  int srcTileH, srcNtx;    // tile height / tiles per tile row of the source level's own destination tiling (level 1: unused)
  ResizeTaps T; ResizeTma R;
};
struct PyrLaunch {
  int nlev, nframes, z0;   // destination levels, frames of this launch, z of the chunk's first frame in the level-0 map
  int start[kMaxLevels + 1];   // first ticket of destination level index i (level i + 1); start[nlev] = all tickets
};
constexpr int kPyrSyncStride = kRzMaxTY;   // completion counters per (frame, level)

}  // namespace orbx
