/* orb_b200.h -- C ABI of the B200-native ORB front-end (libvoslam_b200.so).
 *
 * Drop-in boundary for the data-parallel front-end hot path of guisongchen/vo_slam_test:
 *   - ORB feature extraction      reference: ORB_SLAM2::ORBextractor (include/myslam/ORBextractor.h:45-108,
 *                                 src/ORBextractor.cpp:1051-1112 operator())
 *   - binary-descriptor matching  reference: myslam::Matcher (include/myslam/matcher.h:16-37,
 *                                 src/matcher.cpp) and the Frame grid index (src/frame.cpp:72-97,199-247)
 *
 * Plain C: pointers and sizes only, no C++/torch/OpenCV types.  Every function returns an int status
 * (ORBX_OK == 0, negative == error) and never throws.  "host" entry points take host pointers and do
 * the H2D/D2H copies themselves; "_device" entry points take device pointers that are already resident
 * and run on the caller's CUDA stream (cudaStream_t passed as void*; NULL = default stream).
 * A handle is not re-entrant (like the reference extractor: ORBextractor.h:85): one handle per thread.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns ORBX_ERR_CUDA.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORBX_OK = 0,
  ORBX_ERR_ARG = -1,        /* null pointer / bad size / unsupported shape */
  ORBX_ERR_CUDA = -2,       /* CUDA runtime error or no device */
  ORBX_ERR_CAPACITY = -3,   /* caller buffer too small */
  ORBX_ERR_SHAPE = -4       /* image too small / aspect ratio the reference cannot handle either */
};

/* Same five parameters as the reference constructor (ORBextractor.h:51-52), plus the device. */
typedef struct orbx_params {
  int nfeatures;
  float scale_factor;
  int nlevels;
  int ini_th_fast;
  int min_th_fast;
  int device;               /* CUDA device ordinal */
} orbx_params;

/* Binary layout of cv::KeyPoint (28 bytes): what ORBextractor::operator() fills (ORBextractor.cpp:1081-1110). */
typedef struct orbx_keypoint {
  float x, y;               /* pt, already multiplied by the level scale factor (ORBextractor.cpp:1102-1108) */
  float size;               /* (int)(31 * scale[level])                        (ORBextractor.cpp:845,854) */
  float angle;              /* IC_Angle, degrees in [0,360)                    (ORBextractor.cpp:79-107)  */
  float response;           /* FAST score                                                                  */
  int32_t octave;           /* pyramid level                                                               */
  int32_t class_id;         /* always -1                                                                   */
} orbx_keypoint;

typedef struct orbx_extractor* orbx_handle;

const char* orbx_last_error(void);        /* thread-local message of the last failing call */
int orbx_device_count(int* n);
/* Pinned (page-locked) host staging memory for the batch entry points, visible to every CUDA context of the process
 * (cudaHostAllocPortable); write_combined != 0 adds cudaHostAllocWriteCombined: the host writes frames once and never reads
 * them back, the GPU's copy engine reads them without snooping the CPU caches.  The reference has no equivalent (its frames
 * are cv::Mat buffers read by the CPU, src/frame.cpp:22); a maintainer allocates the grabber's ring with this. */
int orbx_host_alloc(size_t bytes, int write_combined, void** out);
int orbx_host_free(void* p);

/* ---------------------------------------------------------------------------------------------------
 * Extractor  (replaces ORBextractor::ORBextractor / operator() / Get*ScaleFactors)
 * ------------------------------------------------------------------------------------------------- */
int orbx_create(const orbx_params* p, orbx_handle* out);
int orbx_destroy(orbx_handle h);
int orbx_get_levels(orbx_handle h, int* nlevels);                           /* ORBextractor.h:63 */
int orbx_scale_factors(orbx_handle h, float* out, int n);                   /* ORBextractor.h:69 */
int orbx_inv_scale_factors(orbx_handle h, float* out, int n);               /* ORBextractor.h:73 */
int orbx_features_per_level(orbx_handle h, int* out, int n);                /* mnFeaturesPerLevel */
/* Upper bound on keypoints per frame for this handle (sum over levels of N_l + 3). */
int orbx_max_keypoints(orbx_handle h, int* cap);

/* One frame, host buffers: ORBextractor::operator()(image, mask, keypoints, descriptors)
 * (ORBextractor.cpp:1051).  `stride` = bytes between rows.  kps[cap], desc[cap*32].  *n = count.
 * An empty image (w<=0 || h<=0 || !img) is a silent no-op with *n untouched, like the reference (:1054). */
int orbx_extract(orbx_handle h, const uint8_t* img, int w, int height, size_t stride, orbx_keypoint* kps,
                 uint8_t* desc, int cap, int* n);

/* A batch of equally sized frames, host buffers.  Frame f starts at imgs + f*frame_stride.
 * kps[nframes*cap], desc[nframes*cap*32], counts[nframes]. */
int orbx_extract_batch(orbx_handle h, const uint8_t* imgs, int nframes, int w, int height, size_t row_stride,
                       size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts);

/* Same, everything already resident in device memory; asynchronous on `stream`. */
int orbx_extract_batch_device(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height,
                              size_t row_stride, size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc,
                              int cap, int32_t* d_counts, void* stream);

/* Host-buffer batch: extraction of every frame + frame-to-frame Hamming top-2 (BASELINE config 2; pair p: queries
 * = descriptors of frame p, train = descriptors of frame p+1; the inner loop of matcher.cpp:481-507 with TH_LOW and
 * ratio as arguments).  Host->device copies, kernels and device->host copies of successive chunks overlap on three
 * streams; pass pinned memory for `imgs` and the outputs.  m_* are [(nframes-1) * cap]. */
int orbx_extract_match_batch(orbx_handle h, const uint8_t* imgs, int nframes, int w, int height, size_t row_stride,
                             size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts, int th,
                             float ratio, int32_t* m_idx, int32_t* m_d1, int32_t* m_d2, uint8_t* m_ok);

/* Per-stage device time of one resident batch, from CUDA events recorded on `stream` around every stage:
 * ms[0..4] = pyramid, FAST, quadtree, blur, orientation+descriptors (summed over chunks), ms[5] = whole batch. */
int orbx_profile_stages(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height, size_t row_stride,
                        size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts,
                        void* stream, float* ms);

/* Number of kernels this library launched on this handle since creation (bench bookkeeping). */
int orbx_launch_count(orbx_handle h, long long* n);

/* Stage taps for stage-level parity tests (valid after an extract call; frame index within the last
 * chunk processed).  Host output buffers. */
int orbx_debug_level(orbx_handle h, int frame, int level, int blurred, uint8_t* out, int* w, int* height);
int orbx_debug_candidates(orbx_handle h, int frame, int level, int32_t* xys /* n*3 */, int cap, int* n);
int orbx_debug_selected(orbx_handle h, int frame, int level, int32_t* xys /* n*3 */, int cap, int* n);

/* ---------------------------------------------------------------------------------------------------
 * Hamming matcher  (replaces Matcher::computeDistance matcher.cpp:1240-1256 and the best/second-best +
 * ratio loop matcher.cpp:481-507, generalised to all pairs)
 *
 * For each query i: train rows scanned in ascending index, strict '<', first index wins:
 *   idx[i], d1[i] = best; d2[i] = second best (256 if none); ok[i] = d1 <= th && (float)d1 < ratio*(float)d2.
 * ------------------------------------------------------------------------------------------------- */
int hamm_knn2(const uint8_t* q, int nq, const uint8_t* t, long long nt, int th, float ratio, int32_t* idx,
              int32_t* d1, int32_t* d2, uint8_t* ok, int device);
int hamm_knn2_device(const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt, int th, float ratio,
                     int32_t* d_idx, int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, void* d_workspace,
                     size_t workspace_bytes, void* stream);
size_t hamm_knn2_workspace_bytes(int nq, long long nt);

/* Frame-to-frame matching over a batch of extractor outputs (BASELINE config 2): pair p matches the
 * descriptors of frame qf[p] (queries) against those of frame tf[p] (train); desc is [nframes][cap][32],
 * counts[nframes]; outputs are [npairs][cap]. */
int hamm_knn2_pairs_device(const uint8_t* d_desc, const int32_t* d_counts, int cap, const int32_t* d_qf,
                           const int32_t* d_tf, int npairs, int th, float ratio, int32_t* d_idx, int32_t* d_d1,
                           int32_t* d_d2, uint8_t* d_ok, void* stream);

/* Merge of per-shard results of a train set split into `nshards` contiguous index ranges (BASELINE config 5):
 * in_* are [nshards][nq] with idx already global; lowest index wins ties, second = 2nd smallest of the
 * multiset {d1_s, d2_s}.  Used after the NCCL all-gather. */
int hamm_knn2_merge_device(const int32_t* d_idx_in, const int32_t* d_d1_in, const int32_t* d_d2_in, int nshards,
                           int nq, int th, float ratio, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2,
                           uint8_t* d_ok, void* stream);

/* Sharded top-2 with the exchange fused into the kernels (BASELINE config 5): instead of an NCCL all-gather, the
 * merge kernel of every rank STORES its per-shard records straight into exchange buffers in all peers' HBM (CUDA IPC
 * mappings over NVLink / NVSwitch) and releases a per-source flag; the final merge acquires the flags in its own buffer.
 *   hamm_exchange_alloc   this rank's buffer + its 64-byte IPC handle (exchange the handles with any host-side all-gather)
 *   hamm_exchange_open    map a peer's buffer from its handle;  hamm_exchange_close / hamm_exchange_free at shutdown
 *   hamm_knn2_sharded_device  bufs[world] = every rank's buffer as seen from this process (bufs[rank] = own);
 *                         epoch = 1, 2, 3, ... identical on all ranks per call; d_status (int32, zero-initialised) becomes
 *                         1 if a peer's flag did not arrive within ~1 s (the kernel never hangs).
 * Results are identical to hamm_knn2 over the concatenated train set. */
size_t hamm_exchange_bytes(int world, int max_queries);
int hamm_exchange_alloc(int device, int world, int max_queries, void** buf, unsigned char ipc_handle[64]);
int hamm_exchange_open(int device, const unsigned char ipc_handle[64], void** peer_buf);
int hamm_exchange_close(void* peer_buf);
int hamm_exchange_free(void* buf);
int hamm_knn2_sharded_device(const uint8_t* d_q, int nq, const uint8_t* d_t_local, long long nt_local, long long shard_lo, int th,
                             float ratio, int rank, int world, void* const* bufs, int max_queries, int epoch, int32_t* d_idx,
                             int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, int32_t* d_status, void* d_workspace,
                             size_t workspace_bytes, void* stream);

/* Same call split in two: phases bit 0 = local shard scan + scatter into the peers, bit 1 = flag wait + merge (3 = both).
 * Lets a single process that plays several ranks on one device enqueue every scatter before any merge (tests). */
int hamm_knn2_sharded_phases_device(const uint8_t* d_q, int nq, const uint8_t* d_t_local, long long nt_local, long long shard_lo,
                                    int th, float ratio, int rank, int world, void* const* bufs, int max_queries, int epoch,
                                    int32_t* d_idx, int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, int32_t* d_status,
                                    void* d_workspace, size_t workspace_bytes, void* stream, int phases);

long long hamm_launch_count(void);

/* The same job with everything resident in device memory (frames in, keypoints / descriptors / matches out), asynchronous on
 * `stream`: the frame-to-frame top-2 of a chunk is launched on the chunk's lane right behind its extraction, so the POPC-bound
 * matching of one lane runs under the extraction kernels of the other instead of alone at the end of the batch.
 * d_m_* are [(nframes-1) * cap].  Results are identical to orbx_extract_batch_device + hamm_knn2_pairs_device. */
int orbx_extract_match_batch_device(orbx_handle h, const uint8_t* d_imgs, int nframes, int w, int height, size_t row_stride,
                                    size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, int th,
                                    float ratio, int32_t* d_m_idx, int32_t* d_m_d1, int32_t* d_m_d2, uint8_t* d_m_ok, void* stream);

/* Which kernels serve the hamm_* entry points and the frame-to-frame matching of orbx_extract_match_batch:
 *   0 (default)  XOR + POPC on the integer pipe (carry-save tree, 5 POPC per pair) -- the design BASELINE's north_star states;
 *   1            descriptor bits as +-1 bytes through the legacy integer tensor pipe (mma.sync m16n8k32 s8 -> IMMA.16832),
 *                dot = 256 - 2 * hamming, exact integers: an A/B variant with bit-identical results (tests/test_gpu_hamming_mma.py).
 * Returns the previous value; any other argument only queries.  Initial value: environment ORBX_HAMM_MMA. */
int hamm_set_variant(int variant);
/* Map-scale scans hand the train rows out in blocks from a counter per query tile (all CTAs finish together whatever the warp
 * schedulers favour) instead of giving every CTA a fixed range: 0 = never, 1 = automatic (default: long scans only), 2 = whenever
 * the scan is split at all (tests).  Same results.  Returns the previous mode. */
int hamm_set_dynamic(int mode);

/* ---------------------------------------------------------------------------------------------------
 * Grid index + projection searches
 *   grid:  Frame::assignFeaturesToGrid / getFeaturesInArea   (frame.cpp:72-97,199-247; camera.h:8-9)
 *   sbp_frame: Matcher::searchByProjection(Frame*,Frame*,radius,checkRot)        (matcher.cpp:18-148)
 *   sbp_local: Matcher::searchByProjection(Frame*,vector<MapPoint*>&,thRadius)    (matcher.cpp:274-353)
 * Map points are passed already projected (u, v, invz ...), so no Sophus/Eigen arithmetic is involved.
 * ------------------------------------------------------------------------------------------------- */
#define ORBX_GRID_COLS 64
#define ORBX_GRID_ROWS 48

typedef struct orbx_frame_view {      /* the "current frame" side; host or device pointers per entry point */
  const orbx_keypoint* kps;           /* unKeypoints_ */
  const uint8_t* desc;                /* descriptors_, n x 32 */
  const float* uright;                /* uRight_ (<=0: no stereo check) */
  int n;
  float xmin, xmax, ymin, ymax;       /* image bounds (camera.cpp:42-45) */
  const float* scale_factors;         /* scaleFactors_ */
  int nlevels;
  const uint8_t* occupied0;           /* 1 if mappoints_[i] already holds a point with observations */
} orbx_frame_view;

typedef struct orbx_sbp_frame_points { /* map points of the last frame, in index order */
  int m;
  const uint8_t* valid;               /* 0 models `!mp || outlier` (matcher.cpp:46) */
  const float* u; const float* v; const float* invz;   /* projection into the current frame (:49-58) */
  const int32_t* octave;              /* frame_last->unKeypoints_[i].octave */
  const float* angle;                 /* frame_last->unKeypoints_[i].angle */
  const uint8_t* desc;                /* mp->getDescriptor(), m x 32 */
  const uint8_t* has_obs;             /* mp->observe_cnt_ > 0 */
} orbx_sbp_frame_points;

typedef struct orbx_sbp_local_points { /* local map points (matcher.cpp:278-298) */
  int m;
  const uint8_t* valid;               /* 0 models isBad() || !trackInLocalMap_ */
  const float* u; const float* v; const float* ur;   /* trackProj_u_/v_/uR_ */
  const int32_t* level;               /* trackScaleLevel_ */
  const float* view_cos;              /* viewCos_ */
  const uint8_t* desc;
  const uint8_t* has_obs;             /* getObsCnt() > 0 */
} orbx_sbp_local_points;

/* CSR grid: cell_start[64*48+1], ids[n]; cell index = ix*48 + iy; ids ascending inside a cell. Host buffers. */
int orbx_grid_build(const orbx_keypoint* kps, int n, float xmin, float xmax, float ymin, float ymax,
                    int32_t* cell_start, int32_t* ids, int device);

/* assign[i] (i < frame.n): index of the map point finally written to mappoints_[i]; -1 = never written;
 * -2 = written, then cleared by the rotation-histogram check.  *match_cnt = the function's return value. */
int orbx_search_by_projection_frame(const orbx_frame_view* frame, const orbx_sbp_frame_points* pts, float radius,
                                    float bf, int forward, int backward, int check_rot, int32_t* assign,
                                    int* match_cnt, int device);
int orbx_search_by_projection_local(const orbx_frame_view* frame, const orbx_sbp_local_points* pts,
                                    float th_radius, float ratio, int32_t* assign, int* match_cnt, int device);

/* SURVEY section 8f, rank 2 (same kernels, different gates).  Both take orbx_sbp_frame_points with
 *   valid[i]  = every host-side gate of the reference loop passed (null / bad / found set, depth sign, image bounds,
 *               distance range, viewing angle),  octave[i] = mp->predictScale(...),  invz / has_obs unused (pass zeros).
 * reloc: Matcher::searchByProjection(Frame*, KeyFrame*, radius, distThreshold, found, checkRot)   (matcher.cpp:150-272)
 *        frame->occupied0[i] = frame_curr->mappoints_[i] != nullptr;  angle[i] = keyframe->unKeypoints_[i].angle.
 * sim3 : Matcher::searchByProjection(KeyFrame*, Sim3&, loopMapPoints, matchMapPoints, th)         (matcher.cpp:356-447)
 *        `keyframe` describes the keyframe's features; occupied0[i] = matchMapPoints[i] != nullptr on entry.  The
 *        reference tests matchMapPoints[j] with the window position j instead of the feature index (:422); that is
 *        reproduced bit for bit. */
int orbx_search_by_projection_reloc(const orbx_frame_view* frame, const orbx_sbp_frame_points* pts, float radius,
                                    float dist_threshold, int check_rot, int32_t* assign, int* match_cnt, int device);
int orbx_search_by_projection_sim3(const orbx_frame_view* keyframe, const orbx_sbp_frame_points* pts, int th, int32_t* assign,
                                   int* match_cnt, int device);

/* Independent windowed best match per projected point -- the data-parallel core shared by Matcher::searchBySim3
 * (matcher.cpp:717-775,777-836), fuseMapPoints (:1026-1100) and fuseByPose (:1157-1224): window from
 * KeyFrame::getFeaturesInArea (keyframe.cpp:268-312), levels [level_predict-1, level_predict], strict '<' argmin in
 * window order.  best_idx[i] (i < pts->m) = keyframe feature, or -1 when the best distance exceeds dist_threshold.
 * chi2_gate != 0 adds fuseMapPoints' reprojection test (:1073-1095); pts->invz then carries ur = u - bf/z and the
 * keyframe's uright uses the reference's `>= 0` convention.  The pointer-graph surgery that follows in the fuse
 * functions (replaceMapPoint / addObservation) stays on the host and replays best_idx in order. */
int orbx_window_argmin(const orbx_frame_view* keyframe, const orbx_sbp_frame_points* pts, float th_radius, float dist_threshold,
                       int chi2_gate, int32_t* best_idx, int device);
/* Matcher::searchBySim3 (matcher.cpp:679-865): both directed searches (TH_HIGH) and the mutual-consistency check.
 * pts12 / pts21 hold one projected point per feature of kf1 / kf2 (valid = 0 for features without a usable map point). */
int orbx_search_by_sim3(const orbx_frame_view* kf1, const orbx_sbp_frame_points* pts12, const orbx_frame_view* kf2,
                        const orbx_sbp_frame_points* pts21, float th, int32_t* match12, int* found, int device);

/* ---------------------------------------------------------------------------------------------------
 * BoW-guided matching
 *   mode 0: Matcher::searchByBoW(KeyFrame*, Frame*, matches, checkRot)       (matcher.cpp:449-559)
 *           match[i] (i < b.n, frame feature) = index of the keyframe feature whose map point was assigned
 *   mode 1: Matcher::searchByBoW(KeyFrame*, KeyFrame*, matches, checkRot)    (matcher.cpp:561-677)
 *           match[i] (i < a.n, kf1 feature)   = index of the matched kf2 feature
 * -1 = no match, -2 = matched, then cleared by the rotation-histogram check.  Each side passes its DBoW3
 * FeatureVector as a CSR sorted by node id (std::map order): node_ids[ngroups], group_start[ngroups+1],
 * feat_idx[group_start[ngroups]] (feature indices in their vector order).  valid[i] = the feature's map point exists
 * and is not bad (side b in mode 0: all ones).  Host pointers.
 * ------------------------------------------------------------------------------------------------- */
typedef struct orbx_bow_side {
  int n;                       /* features */
  const uint8_t* desc;         /* n x 32 */
  const float* angle;          /* unKeypoints_[i].angle */
  const uint8_t* valid;
  int ngroups;
  const uint32_t* node_ids;
  const int32_t* group_start;
  const int32_t* feat_idx;
} orbx_bow_side;

int orbx_search_by_bow(const orbx_bow_side* a, const orbx_bow_side* b, int mode, float ratio, int th_low, int check_rot,
                       int32_t* match, int* match_cnt, int device);

/* Matcher::searchForTriangulation(KeyFrame*, KeyFrame*, matchIdxs, F12, checkRot)   (matcher.cpp:867-1010; rank 4).
 * side.valid[i] = the feature has NO map point (:902, :921).  F12: row-major 3x3 doubles; (ex, ey): epipole of camera 1 in
 * image 2 (:886-890); scale_factors2: keyframe2->scaleFactors_.  match[i] (i < a->side.n) = matched kf2 feature, -1 none,
 * -2 cleared by the rotation check.  The epipolar test (:1306-1324) is evaluated in double without FMA, left to right. */
typedef struct orbx_tri_side {
  orbx_bow_side side;
  const orbx_keypoint* kps;    /* unKeypoints_ */
  const float* uright;         /* uRight_ (>= 0: stereo) */
} orbx_tri_side;

int orbx_search_for_triangulation(const orbx_tri_side* a, const orbx_tri_side* b, const double* F12, float ex, float ey,
                                  const float* scale_factors2, int nlevels, int th_low, int check_rot, int32_t* match,
                                  int* match_cnt, int device);

/* MapPoint::computeDescriptor (mappoint.cpp:118-179) for `npoints` map points at once.
 * desc: all observed descriptors back to back (32 B each); start[npoints+1]: CSR (observations of point p are rows
 * start[p] .. start[p+1]-1, in the order the reference's std::map<KeyFrame*,size_t> iteration pushes them, bad key frames
 * already removed).  best[p] = row (relative to start[p]) whose median distance to the others is smallest, first row
 * wins ties; -1 for a point without observations (the reference keeps its old descriptor).  Host pointers. */
int orbx_medoid_descriptors(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best, int device);

/* ---------------------------------------------------------------------------------------------------
 * Frame post-processing between extraction and matching: what Frame::Frame does after
 * the extractor call (frame.cpp:22-32) -- undistortKeyPoints (frame.cpp:36-70, cv::undistortPoints(pts, pts, K, D,
 * noArray(), K): 5 iterations in double, bit-exact to OpenCV 4.13), findDepth (frame.cpp:108-133) and
 * assignFeaturesToGrid (frame.cpp:72-97) -- for a batch of frames laid out like the output of orbx_extract_batch.
 * ------------------------------------------------------------------------------------------------- */
typedef struct orbx_camera {
  float fx, fy, cx, cy;               /* camera.cpp:10-13 (K_ is CV_32F, camera.cpp:19-21) */
  float dist[8];                      /* k1 k2 p1 p2 [k3 [k4 k5 k6]]; camera.cpp:27-38 builds 4 or 5 of them */
  int ndist;                          /* 0..8; dist[0] == 0 copies the keypoints unchanged (frame.cpp:41-45) */
  float bf;                           /* camera_bf (frame.cpp:116,129) */
  float xmin, xmax, ymin, ymax;       /* image bounds of the grid (camera.cpp:42-45) */
} orbx_camera;

/* kps[nframes*cap] + counts[nframes]: extractor output.  depth: nframes float32 images (height x w, strides in BYTES),
 * sampled at the ORIGINAL keypoint with float->int truncation (frame.cpp:121-123); NULL = no depth (all -1).
 * Outputs per frame f (entries at or beyond counts[f] are unspecified):
 *   unkps[f*cap + i]     unKeypoints_[i]            uright[f*cap + i]  uRight_[i]  (-1 where depth <= 0)
 *   depth_out[f*cap + i] depth_[i]                  cell_start[f*(64*48+1) + c], ids[f*cap + ..]  CSR grid as orbx_grid_build
 * Device variant: everything resident, asynchronous on `stream`. */
int orbx_frame_finish(const orbx_camera* cam, const orbx_keypoint* kps, const int32_t* counts, int nframes, int cap,
                      const float* depth, int w, int height, size_t depth_row_stride, size_t depth_frame_stride,
                      orbx_keypoint* unkps, float* uright, float* depth_out, int32_t* cell_start, int32_t* ids,
                      int device);
int orbx_frame_finish_device(const orbx_camera* cam, const orbx_keypoint* d_kps, const int32_t* d_counts, int nframes,
                             int cap, const float* d_depth, int w, int height, size_t depth_row_stride,
                             size_t depth_frame_stride, orbx_keypoint* d_unkps, float* d_uright, float* d_depth_out,
                             int32_t* d_cell_start, int32_t* d_ids, int device, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Device-resident frame: the reference's per-frame call chain without host round trips.
 *   Frame::Frame (frame.cpp:22-32) = extractor call + undistortKeyPoints + findDepth + assignFeaturesToGrid, then the tracking
 *   thread runs 1-4 searches against that frame (visualOdometry.cpp:240 searchByProjection(Frame*,Frame*), :265 / :329
 *   searchByBoW(KeyFrame*,Frame*), :329 searchByProjection(Frame*,KeyFrame*), :354 searchByProjection(Frame*,local map)).
 * orbx_frame_create does the whole constructor on the device in one call (image up, kernels, ONE packed copy of keypoints /
 * descriptors / undistorted keypoints / uRight / depth down) and keeps everything the searches read -- undistorted keypoints,
 * descriptors, uRight, the 64x48 CSR grid -- in HBM; the *_h searches take the handle, so per search only the projected map
 * points and the frame's `occupied0` flags go up and the assignment comes down.  Results are bit-identical to
 * orbx_extract + orbx_frame_finish + the host-array searches.
 *   depth: float32 image (height x w, row stride in BYTES, host memory) or NULL.  findDepth needs it only at the ~1000 keypoint
 *   positions, so the image is never uploaded: the host samples it after the packed copy down and sends uRight_ (4 KB) back.
 *   A frame belongs to the extractor handle that created it (its stream, its
 *   block pool): destroy frames before their extractor; orbx_destroy frees whatever is left.  One thread per extractor.
 * ------------------------------------------------------------------------------------------------- */
typedef struct orbx_frame* orbx_frame_t;
int orbx_frame_create(orbx_handle h, const orbx_camera* cam, const uint8_t* img, int w, int height, size_t stride,
                      const float* depth, size_t depth_row_stride, orbx_frame_t* out, int* n);
/* Make an existing host-side Frame / KeyFrame resident (e.g. a key frame that relocalisation or loop closing searches into
 * repeatedly): unKeypoints_ / descriptors_ / uRight_ / scaleFactors_ of `view` go up once, the grid is built on the device;
 * view->occupied0 is ignored.  The handle works with every *_h search; orbx_frame_get is not available for it. */
int orbx_frame_upload(const orbx_frame_view* view, int device, orbx_frame_t* out);
int orbx_frame_size(orbx_frame_t f, int* n);
/* host copies of keypoints_ / descriptors_ / unKeypoints_ / uRight_ / depth_ (any pointer may be NULL; cap >= n) */
int orbx_frame_get(orbx_frame_t f, orbx_keypoint* kps, uint8_t* desc, orbx_keypoint* unkps, float* uright, float* depth, int cap);
/* the frame's 64x48 grid as the CSR of orbx_grid_build (Frame::gridKeypoints_, frame.cpp:72-89); cap >= number of gridded ids */
int orbx_frame_grid(orbx_frame_t f, int32_t* cell_start, int32_t* ids, int cap);
int orbx_frame_destroy(orbx_frame_t f);
/* occupied0[i] (i < n): as orbx_frame_view.occupied0 of the corresponding host-array search */
int orbx_search_by_projection_frame_h(orbx_frame_t frame, const uint8_t* occupied0, const orbx_sbp_frame_points* pts, float radius,
                                      float bf, int forward, int backward, int check_rot, int32_t* assign, int* match_cnt);
int orbx_search_by_projection_local_h(orbx_frame_t frame, const uint8_t* occupied0, const orbx_sbp_local_points* pts,
                                      float th_radius, float ratio, int32_t* assign, int* match_cnt);
int orbx_search_by_projection_reloc_h(orbx_frame_t frame, const uint8_t* occupied0, const orbx_sbp_frame_points* pts, float radius,
                                      float dist_threshold, int check_rot, int32_t* assign, int* match_cnt);
/* searchByBoW(KeyFrame*, Frame*) (mode 0 of orbx_search_by_bow): frame_groups carries the frame's n, valid flags and
 * FeatureVector CSR; its desc / angle may be NULL (read from the handle). */
int orbx_search_by_bow_h(const orbx_bow_side* keyframe, orbx_frame_t frame, const orbx_bow_side* frame_groups, float ratio,
                         int th_low, int check_rot, int32_t* match, int* match_cnt);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
