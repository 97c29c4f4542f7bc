// tests/tools/helpers_check.cpp -- TEST INFRASTRUCTURE.  The stand-in object types (oracle/compat_myslam/myslam_stub.hpp) restate a
// few helpers of frame.cpp / keyframe.cpp / mappoint.cpp / camera.cpp; the reference's own matcher.cpp, compiled in place, runs
// on top of them.  This program holds each restated helper against the reference's own code for it (line ranges compiled in
// place, oracle/_ref/librefhelpers.so) on random inputs.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "myslam_stub.hpp"

extern "C" {
int refh_grid_build(const void* kps, int n, float xmin, float xmax, float ymin, float ymax, int* cell_start, int* ids);
int refh_features_in_area(const void* kps, int n, float xmin, float xmax, float ymin, float ymax, const float* u, const float* v,
                          const float* r, const int* min_level, const int* max_level, int nq, int* out, int* out_start, int cap);
int refh_is_in_img(float xmin, float xmax, float ymin, float ymax, float u, float v);
int refh_predict_scale(float maxDistance, float currDist, const float* scaleFactors, int nlevels, int keyframe);
void refh_camera2pixel(float fx, float fy, float cx, float cy, double x, double y, double z, double* u, double* v);
}

using namespace myslam;
static unsigned g_seed = 7;
static unsigned rnd() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 8; }
static float urand() { return (float)(rnd() & 0xFFFFFF) / 16777216.0f; }

int main() {
  int bad = 0;
  const int W = 640, H = 480, N = 2000, NQ = 2000;
  Frame f; KeyFrame kf;
  f.N_ = kf.N_ = N;
  f.unKeypoints_.resize(N);
  for (int i = 0; i < N; ++i)
    f.unKeypoints_[i] = cv::KeyPoint(urand() * (W + 20) - 10, urand() * (H + 20) - 10, 31.f, 0.f, 0.f, (int)(rnd() % 8), -1);
  kf.unKeypoints_ = f.unKeypoints_;
  f.xMin_ = kf.xMin_ = 0; f.yMin_ = kf.yMin_ = 0; f.xMax_ = kf.xMax_ = W; f.yMax_ = kf.yMax_ = H;
  f.gridPerPixelWidth_ = kf.gridPerPixelWidth_ = 64.f / W; f.gridPerPixelHeight_ = kf.gridPerPixelHeight_ = 48.f / H;
  f.assignFeaturesToGrid(); kf.assignFeaturesToGrid();

  // grid buckets
  std::vector<int> cs(64 * 48 + 1), ids(N);
  refh_grid_build(f.unKeypoints_.data(), N, 0, W, 0, H, cs.data(), ids.data());
  for (int ix = 0; ix < 64; ++ix)
    for (int iy = 0; iy < 48; ++iy) {
      const int c = ix * 48 + iy;
      std::vector<int> want(ids.begin() + cs[c], ids.begin() + cs[c + 1]);
      if (want != f.gridKeypoints_[ix][iy] || want != kf.gridKeypoints_[ix][iy]) ++bad;
    }
  std::printf("assignFeaturesToGrid: %d differing buckets\n", bad);

  // window queries, Frame (level window) and KeyFrame (no level filter)
  std::vector<float> u(NQ), v(NQ), r(NQ);
  std::vector<int> lo(NQ), hi(NQ), out((size_t)N * 64), os(NQ + 1);
  for (int q = 0; q < NQ; ++q) {
    u[q] = urand() * (W + 60) - 30; v[q] = urand() * (H + 60) - 30; r[q] = 2.f + urand() * (q % 7 == 0 ? 300.f : 40.f);
    lo[q] = (q % 3 == 0) ? -1 : (int)(rnd() % 7); hi[q] = lo[q] + (int)(rnd() % 3);
  }
  // the query list is long: run it in slices so that `out` always has room
  int badq = 0; long hits = 0;
  for (int q0 = 0; q0 < NQ; q0 += 50) {
    refh_features_in_area(f.unKeypoints_.data(), N, 0, W, 0, H, &u[q0], &v[q0], &r[q0], &lo[q0], &hi[q0], 50, out.data(), os.data(), (int)out.size());
    for (int q = q0; q < q0 + 50; ++q) {
      std::vector<int> want(out.begin() + os[q - q0], out.begin() + os[q - q0 + 1]);
      std::vector<int> got = lo[q] < 0 ? kf.getFeaturesInArea(u[q], v[q], r[q]) : f.getFeaturesInArea(u[q], v[q], r[q], lo[q], hi[q]);
      if (want != got) ++badq;
      hits += (long)want.size();
    }
  }
  std::printf("getFeaturesInArea: %d of %d queries differ (%ld features returned)\n", badq, NQ, hits);
  bad += badq + (hits < NQ);

  int badi = 0, badp = 0, badc = 0;
  float sf[8]; sf[0] = 1.f;
  for (int l = 1; l < 8; ++l) sf[l] = (float)(sf[l - 1] * (double)1.2f);
  f.scaleFactors_.assign(sf, sf + 8); kf.scaleFactors_.assign(sf, sf + 8);
  Camera cam; cam.fx_ = 517.3f; cam.fy_ = 516.5f; cam.cx_ = 318.6f; cam.cy_ = 255.3f;
  for (int t = 0; t < 20000; ++t) {
    const float uu = (t % 5 == 0) ? (float)(rnd() % 3) * 320.f : urand() * 700.f - 30.f, vv = (t % 7 == 0) ? (float)(rnd() % 3) * 240.f : urand() * 540.f - 30.f;
    if ((int)kf.isInImg(uu, vv) != refh_is_in_img(0, W, 0, H, uu, vv)) ++badi;
    MapPoint mp; mp.maxDistance_ = 0.5f + urand() * 20.f;
    const float dist = 0.05f + urand() * 30.f;
    if (mp.predictScale(dist, &f) != refh_predict_scale(mp.maxDistance_, dist, sf, 8, 0)) ++badp;
    if (mp.predictScale(dist, &kf) != refh_predict_scale(mp.maxDistance_, dist, sf, 8, 1)) ++badp;
    const Vector3d p(urand() * 4 - 2, urand() * 4 - 2, 0.2 + urand() * 8);
    double ru, rv;
    refh_camera2pixel(cam.fx_, cam.fy_, cam.cx_, cam.cy_, p[0], p[1], p[2], &ru, &rv);
    const Vector2d px = cam.camera2pixel(p);
    if (px[0] != ru || px[1] != rv) ++badc;
  }
  std::printf("isInImg: %d, predictScale: %d, camera2pixel: %d differences in 20000 trials each\n", badi, badp, badc);
  bad += badi + badp + badc;
  std::printf(bad ? "helpers: %d DIFFERENCES\n" : "helpers: stand-in types agree with the reference's own code\n", bad);
  return bad ? 1 : 0;
}
