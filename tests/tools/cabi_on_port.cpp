// tests/tools/cabi_on_port.cpp -- TEST INFRASTRUCTURE.  The handful of C-ABI entry points include/orb_b200_matcher.hpp calls,
// forwarded to the CPU oracle port (oracle/liborbport.so).  Linking tests/tools/matcher_adapter_check.cpp with this file
// INSTEAD of libvoslam_b200.so checks the adapter's host logic (object walk, flattening, replay of the writes) on a machine
// without a GPU: RefMatcher (loops over objects) must agree with adapter + port (flat arrays).  Never part of the product.
#include <cstdint>
#include <cstring>

#include "orb_b200.h"

struct port_sbp_frame_in {
  const void* kps; const uint8_t* desc; const float* uright; int n;
  float xmin, xmax, ymin, ymax;
  const float* scale_factors; int nlevels;
  const uint8_t* occupied0;
  int m; const uint8_t* valid; const float* u; const float* v; const float* invz; const int32_t* octave;
  const float* angle; const uint8_t* mp_desc; const uint8_t* has_obs;
  float radius; float bf; int forward; int backward; int check_rot;
};
struct port_sbp_local_in {
  const void* kps; const uint8_t* desc; const float* uright; int n;
  float xmin, xmax, ymin, ymax;
  const float* scale_factors; int nlevels;
  const uint8_t* occupied0;
  int m; const uint8_t* valid; const float* u; const float* v; const float* ur; const int32_t* level;
  const float* view_cos; const uint8_t* mp_desc; const uint8_t* has_obs;
  float th_radius; float ratio;
};
struct port_bow_side {
  int n; const uint8_t* desc; const float* angle; const uint8_t* valid;
  int ngroups; const uint32_t* node_ids; const int32_t* group_start; const int32_t* feat_idx;
};
struct port_tri_side { port_bow_side side; const void* kps; const float* uright; };
struct port_camera { float fx, fy, cx, cy; float dist[8]; int ndist; float bf; float xmin, xmax, ymin, ymax; };
extern "C" {
int port_frame_finish(const port_camera* cam, const void* kps_in, int n, const float* depth, int W, int H, size_t depth_step,
                      void* unkps_out, float* uright, float* depth_out, int* cell_start, int* ids);
int port_search_for_triangulation(const port_tri_side* a, const port_tri_side* b, const double* F12, float ex, float ey,
                                  const float* scale2, int th_low, int check_rot, int32_t* match);
int port_sbp_frame(const port_sbp_frame_in* in, int32_t* assign);
int port_sbp_local(const port_sbp_local_in* in, int32_t* assign);
int port_sbp_reloc(const port_sbp_frame_in* in, float dist_threshold, int32_t* assign);
int port_sbp_sim3(const port_sbp_frame_in* in, int th, int32_t* assign);
void port_medoid(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best);
void port_window_argmin(const port_sbp_frame_in* in, float th_radius, float dist_threshold, int chi2, int32_t* best);
int port_search_by_sim3(const port_sbp_frame_in* in12, const port_sbp_frame_in* in21, float th, int32_t* match12);
int port_search_by_bow(const port_bow_side* a, const port_bow_side* b, int mode, float ratio, int th_low, int check_rot,
                       int32_t* match);
void port_knn2(const uint8_t* q, int Q, const uint8_t* t, long long M, int th, float ratio, int32_t* idx, int32_t* d1,
               int32_t* d2, uint8_t* ok, int nthreads);

int orbx_device_count(int* n) { *n = 1; return ORBX_OK; }
const char* orbx_last_error(void) { return "cabi_on_port"; }

int orbx_search_by_projection_frame(const orbx_frame_view* f, const orbx_sbp_frame_points* p, float radius, float bf,
                                    int forward, int backward, int check_rot, int32_t* assign, int* match_cnt, int) {
  port_sbp_frame_in in = {f->kps, f->desc, f->uright, f->n, f->xmin, f->xmax, f->ymin, f->ymax, f->scale_factors, f->nlevels,
                          f->occupied0, p->m, p->valid, p->u, p->v, p->invz, p->octave, p->angle, p->desc, p->has_obs,
                          radius, bf, forward, backward, check_rot};
  *match_cnt = port_sbp_frame(&in, assign);
  return ORBX_OK;
}
int orbx_search_by_projection_reloc(const orbx_frame_view* f, const orbx_sbp_frame_points* p, float radius, float dist_threshold,
                                    int check_rot, int32_t* assign, int* match_cnt, int) {
  port_sbp_frame_in in = {f->kps, f->desc, f->uright, f->n, f->xmin, f->xmax, f->ymin, f->ymax, f->scale_factors, f->nlevels,
                          f->occupied0, p->m, p->valid, p->u, p->v, p->invz, p->octave, p->angle, p->desc, p->has_obs,
                          radius, 0.f, 0, 0, check_rot};
  *match_cnt = port_sbp_reloc(&in, dist_threshold, assign);
  return ORBX_OK;
}
int orbx_search_by_projection_sim3(const orbx_frame_view* f, const orbx_sbp_frame_points* p, int th, int32_t* assign,
                                   int* match_cnt, int) {
  port_sbp_frame_in in = {f->kps, f->desc, f->uright, f->n, f->xmin, f->xmax, f->ymin, f->ymax, f->scale_factors, f->nlevels,
                          f->occupied0, p->m, p->valid, p->u, p->v, p->invz, p->octave, p->angle, p->desc, p->has_obs,
                          0.f, 0.f, 0, 0, 0};
  *match_cnt = port_sbp_sim3(&in, th, assign);
  return ORBX_OK;
}
int orbx_search_by_sim3(const orbx_frame_view* kf1, const orbx_sbp_frame_points* p12, const orbx_frame_view* kf2,
                        const orbx_sbp_frame_points* p21, float th, int32_t* match12, int* found, int) {
  // the port pairs "searched frame" with "projected points": kf1's points are searched in kf2 and vice versa
  port_sbp_frame_in in12 = {kf2->kps, kf2->desc, kf2->uright, kf2->n, kf2->xmin, kf2->xmax, kf2->ymin, kf2->ymax, kf2->scale_factors,
                            kf2->nlevels, kf2->occupied0, p12->m, p12->valid, p12->u, p12->v, p12->invz, p12->octave, p12->angle,
                            p12->desc, p12->has_obs, 0.f, 0.f, 0, 0, 0};
  port_sbp_frame_in in21 = {kf1->kps, kf1->desc, kf1->uright, kf1->n, kf1->xmin, kf1->xmax, kf1->ymin, kf1->ymax, kf1->scale_factors,
                            kf1->nlevels, kf1->occupied0, p21->m, p21->valid, p21->u, p21->v, p21->invz, p21->octave, p21->angle,
                            p21->desc, p21->has_obs, 0.f, 0.f, 0, 0, 0};
  *found = port_search_by_sim3(&in12, &in21, th, match12);
  return ORBX_OK;
}
int orbx_window_argmin(const orbx_frame_view* f, const orbx_sbp_frame_points* p, float th_radius, float dist_threshold, int chi2,
                       int32_t* best, int) {
  port_sbp_frame_in in = {f->kps, f->desc, f->uright, f->n, f->xmin, f->xmax, f->ymin, f->ymax, f->scale_factors, f->nlevels,
                          f->occupied0, p->m, p->valid, p->u, p->v, p->invz, p->octave, p->angle, p->desc, p->has_obs,
                          0.f, 0.f, 0, 0, 0};
  port_window_argmin(&in, th_radius, dist_threshold, chi2, best);
  return ORBX_OK;
}
int orbx_search_for_triangulation(const orbx_tri_side* a, const orbx_tri_side* b, const double* F12, float ex, float ey,
                                  const float* scale_factors2, int, int th_low, int check_rot, int32_t* match, int* match_cnt, int) {
  static_assert(sizeof(port_tri_side) == sizeof(orbx_tri_side), "same layout");
  *match_cnt = port_search_for_triangulation((const port_tri_side*)a, (const port_tri_side*)b, F12, ex, ey, scale_factors2, th_low,
                                             check_rot, match);
  return ORBX_OK;
}
int orbx_frame_finish(const orbx_camera* cam, const orbx_keypoint* kps, const int32_t* counts, int nframes, int cap, const float* depth,
                      int w, int height, size_t depth_row_stride, size_t depth_frame_stride, orbx_keypoint* unkps, float* uright,
                      float* depth_out, int32_t* cell_start, int32_t* ids, int) {
  static_assert(sizeof(port_camera) == sizeof(orbx_camera), "same layout");
  for (int f = 0; f < nframes; ++f)
    port_frame_finish((const port_camera*)cam, kps + (size_t)f * cap, counts[f],
                      depth ? (const float*)((const char*)depth + f * depth_frame_stride) : nullptr, w, height, depth_row_stride,
                      unkps + (size_t)f * cap, uright + (size_t)f * cap, depth_out + (size_t)f * cap,
                      cell_start + (size_t)f * (ORBX_GRID_COLS * ORBX_GRID_ROWS + 1), ids + (size_t)f * cap);
  return ORBX_OK;
}
int orbx_medoid_descriptors(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best, int) {
  port_medoid(desc, start, npoints, best);
  return ORBX_OK;
}
int orbx_search_by_projection_local(const orbx_frame_view* f, const orbx_sbp_local_points* p, float th_radius, float ratio,
                                    int32_t* assign, int* match_cnt, int) {
  port_sbp_local_in in = {f->kps, f->desc, f->uright, f->n, f->xmin, f->xmax, f->ymin, f->ymax, f->scale_factors, f->nlevels,
                          f->occupied0, p->m, p->valid, p->u, p->v, p->ur, p->level, p->view_cos, p->desc, p->has_obs,
                          th_radius, ratio};
  *match_cnt = port_sbp_local(&in, assign);
  return ORBX_OK;
}
int orbx_search_by_bow(const orbx_bow_side* a, const orbx_bow_side* b, int mode, float ratio, int th_low, int check_rot,
                       int32_t* match, int* match_cnt, int) {
  static_assert(sizeof(port_bow_side) == sizeof(orbx_bow_side), "same layout");
  *match_cnt = port_search_by_bow((const port_bow_side*)a, (const port_bow_side*)b, mode, ratio, th_low, check_rot, match);
  return ORBX_OK;
}
// ---- device-resident frames, port flavour: the "resident" copy is a host copy of the arrays; the *_h searches rebuild the
// view from it, so the adapter's handle path (registry look-up, occupied0 only) runs on a machine without a GPU too
}  // extern "C"
#include <vector>
struct orbx_frame {
  std::vector<orbx_keypoint> kps; std::vector<uint8_t> desc; std::vector<float> uright, scale;
  float xmin, xmax, ymin, ymax;
  orbx_frame_view view(const uint8_t* occupied0) const {
    orbx_frame_view v;
    v.kps = kps.data(); v.desc = desc.data(); v.uright = uright.data(); v.n = (int)kps.size();
    v.xmin = xmin; v.xmax = xmax; v.ymin = ymin; v.ymax = ymax; v.scale_factors = scale.data(); v.nlevels = (int)scale.size();
    v.occupied0 = occupied0;
    return v;
  }
};
extern "C" {
int orbx_frame_upload(const orbx_frame_view* v, int, orbx_frame_t* out) {
  orbx_frame* f = new orbx_frame();
  f->kps.assign(v->kps, v->kps + v->n); f->desc.assign(v->desc, v->desc + (size_t)v->n * 32);
  f->uright.assign(v->uright, v->uright + v->n); f->scale.assign(v->scale_factors, v->scale_factors + v->nlevels);
  f->xmin = v->xmin; f->xmax = v->xmax; f->ymin = v->ymin; f->ymax = v->ymax;
  *out = f;
  return ORBX_OK;
}
int orbx_frame_size(orbx_frame_t f, int* n) { *n = (int)f->kps.size(); return ORBX_OK; }
int orbx_frame_destroy(orbx_frame_t f) { delete f; return ORBX_OK; }
int orbx_search_by_projection_frame_h(orbx_frame_t f, const uint8_t* occ, const orbx_sbp_frame_points* p, float radius, float bf,
                                      int forward, int backward, int check_rot, int32_t* assign, int* match_cnt) {
  const orbx_frame_view v = f->view(occ);
  return orbx_search_by_projection_frame(&v, p, radius, bf, forward, backward, check_rot, assign, match_cnt, 0);
}
int orbx_search_by_projection_local_h(orbx_frame_t f, const uint8_t* occ, const orbx_sbp_local_points* p, float th_radius, float ratio,
                                      int32_t* assign, int* match_cnt) {
  const orbx_frame_view v = f->view(occ);
  return orbx_search_by_projection_local(&v, p, th_radius, ratio, assign, match_cnt, 0);
}
int orbx_search_by_projection_reloc_h(orbx_frame_t f, const uint8_t* occ, const orbx_sbp_frame_points* p, float radius,
                                      float dist_threshold, int check_rot, int32_t* assign, int* match_cnt) {
  const orbx_frame_view v = f->view(occ);
  return orbx_search_by_projection_reloc(&v, p, radius, dist_threshold, check_rot, assign, match_cnt, 0);
}
int orbx_search_by_bow_h(const orbx_bow_side* a, orbx_frame_t f, const orbx_bow_side* groups, float ratio, int th_low, int check_rot,
                         int32_t* match, int* match_cnt) {
  orbx_bow_side b = *groups;
  std::vector<float> angle(f->kps.size() ? f->kps.size() : 1);
  for (size_t i = 0; i < f->kps.size(); ++i) angle[i] = f->kps[i].angle;
  b.desc = f->desc.data(); b.angle = angle.data();            // read from the "resident" copy, like the library does
  return orbx_search_by_bow(a, &b, 0, ratio, th_low, check_rot, match, match_cnt, 0);
}
// The extractor itself is not served by the port here (tests/tools/adapter_check.cpp and the Python tests cover it on the GPU):
// orbx_create fails, and matcher_adapter_check.cpp skips its constructFrame block.
int orbx_create(const orbx_params*, orbx_handle* out) { *out = nullptr; return ORBX_ERR_CUDA; }
int orbx_destroy(orbx_handle) { return ORBX_OK; }
int orbx_max_keypoints(orbx_handle, int* cap) { *cap = 0; return ORBX_ERR_CUDA; }
int orbx_extract(orbx_handle, const uint8_t*, int, int, size_t, orbx_keypoint*, uint8_t*, int, int*) { return ORBX_ERR_CUDA; }
int orbx_frame_create(orbx_handle, const orbx_camera*, const uint8_t*, int, int, size_t, const float*, size_t, orbx_frame_t* out, int*) {
  *out = nullptr;
  return ORBX_ERR_CUDA;
}
int orbx_frame_get(orbx_frame_t, orbx_keypoint*, uint8_t*, orbx_keypoint*, float*, float*, int) { return ORBX_ERR_CUDA; }
int orbx_frame_grid(orbx_frame_t, int32_t*, int32_t*, int) { return ORBX_ERR_CUDA; }
int hamm_knn2(const uint8_t* q, int nq, const uint8_t* t, long long nt, int th, float ratio, int32_t* idx, int32_t* d1,
              int32_t* d2, uint8_t* ok, int) {
  port_knn2(q, nq, t, nt, th, ratio, idx, d1, d2, ok, 1);
  return ORBX_OK;
}
}
