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
extern "C" {
int port_sbp_frame(const port_sbp_frame_in* in, int32_t* assign);
int port_sbp_local(const port_sbp_local_in* in, int32_t* assign);
int port_sbp_reloc(const port_sbp_frame_in* in, float dist_threshold, int32_t* assign);
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
int hamm_knn2(const uint8_t* q, int nq, const uint8_t* t, long long nt, int th, float ratio, int32_t* idx, int32_t* d1,
              int32_t* d2, uint8_t* ok, int) {
  port_knn2(q, nq, t, nt, th, ratio, idx, d1, d2, ok, 1);
  return ORBX_OK;
}
}
