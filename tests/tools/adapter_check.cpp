// tests/tools/adapter_check.cpp -- TEST INFRASTRUCTURE.  Compile/link check of include/orb_b200_adapter.hpp against the
// OpenCV-compat shim and, with a GPU, a byte-for-byte comparison of what the C++ adapter hands back (std::vector<cv::KeyPoint>,
// descriptor Mat) with the CPU oracle: the port (oracle/liborbport.so, argv[1]) and -- when built -- the reference's own
// ORB_SLAM2::ORBextractor class compiled in place (oracle/_ref/liborbref_parity.so, argv[2]).  Covers what the ctypes tests
// bypass: Mat::step != cols, the KeyPoint / descriptor copy-out, the getters.  Exit code != 0 on ANY differing byte.
//   g++ -std=c++11 -Ioracle/compat -Iinclude tests/tools/adapter_check.cpp -Lvo_slam_test_b200/lib -lvoslam_b200 -ldl -o /tmp/adapter_check
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "opencv/cv.h"
#include "orb_b200_adapter.hpp"

struct port_params { int nfeatures; float scale_factor; int nlevels; int ini_th; int min_th; };
typedef int (*port_extract_fn)(const port_params*, const uint8_t*, int, int, size_t, void*, uint8_t*, int, uint8_t*);
typedef void* (*ref_create_fn)(int, float, int, int, int);
typedef void (*ref_destroy_fn)(void*);
typedef int (*ref_extract_fn)(void*, const uint8_t*, int, int, size_t, void*, uint8_t*, int);

static void fill(uint8_t* p, int W, int H, size_t step, unsigned seed) {
  unsigned s = seed;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      s = s * 1664525u + 1013904223u;
      p[(size_t)y * step + x] = (uint8_t)((((x / 24) + (y / 24)) & 1) * 120 + 60 + (s >> 29) + ((x * 7 + y * 3) % 23));
    }
}

static int compare(const char* what, const std::vector<cv::KeyPoint>& kps, const cv::Mat& desc, const std::vector<uint8_t>& rk,
                   const std::vector<uint8_t>& rd, int rn) {
  if ((int)kps.size() != rn || desc.rows != rn) { std::printf("FAIL %s: %zu keypoints, oracle %d\n", what, kps.size(), rn); return 1; }
  for (int i = 0; i < rn; ++i) {
    if (std::memcmp(&kps[i], &rk[(size_t)i * 28], 28) != 0) { std::printf("FAIL %s: keypoint %d differs\n", what, i); return 1; }
    if (std::memcmp(desc.ptr(i), &rd[(size_t)i * 32], 32) != 0) { std::printf("FAIL %s: descriptor %d differs\n", what, i); return 1; }
  }
  std::printf("adapter parity ok vs %s: %d keypoints, %d descriptor bytes identical\n", what, rn, rn * 32);
  return 0;
}

int main(int argc, char** argv) {
  int ndev = 0;
  orbx_device_count(&ndev);
  if (ndev == 0) { std::printf("adapter links; no CUDA device -> compute skipped\n"); return 0; }
  if (argc < 2) { std::printf("usage: adapter_check liborbport.so [liborbref_parity.so]\n"); return 2; }
  void* hp = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!hp) { std::printf("FAIL: cannot load the oracle port %s: %s\n", argv[1], dlerror()); return 2; }
  port_extract_fn port_extract = (port_extract_fn)dlsym(hp, "port_extract");
  if (!port_extract) { std::printf("FAIL: port_extract missing\n"); return 2; }
  ref_create_fn rcreate = nullptr; ref_destroy_fn rdestroy = nullptr; ref_extract_fn rextract = nullptr;
  if (argc > 2) {
    void* hr = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    if (hr) {
      rcreate = (ref_create_fn)dlsym(hr, "orbref_create"); rdestroy = (ref_destroy_fn)dlsym(hr, "orbref_destroy");
      rextract = (ref_extract_fn)dlsym(hr, "orbref_extract");
    }
  }
  struct Case { int W, H; size_t step; int off; int nf; float sf; int nl; unsigned seed; };
  const Case cases[] = {{640, 480, 640, 0, 1000, 1.2f, 8, 12345u},      // contiguous
                        {640, 480, 701, 3, 1000, 1.2f, 8, 777u},        // a view into a wider buffer: step != cols, odd base address
                        {320, 240, 336, 16, 300, 1.2f, 8, 9u},
                        {752, 480, 752, 0, 1500, 1.1f, 6, 31u}};
  int bad = 0, refChecked = 0;
  for (const Case& c : cases) {
    std::vector<uint8_t> buf(c.step * c.H + 64);
    fill(buf.data() + c.off, c.W, c.H, c.step, c.seed);
    cv::Mat img(c.H, c.W, CV_8UC1, buf.data() + c.off, c.step);
    ORB_SLAM2::ORBextractor ex(c.nf, c.sf, c.nl, 20, 7);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    ex(img, cv::Mat(), kps, desc);
    const int cap = 1 << 15;
    std::vector<uint8_t> rk((size_t)cap * 28), rd((size_t)cap * 32);
    port_params pp = {c.nf, c.sf, c.nl, 20, 7};
    int rn = port_extract(&pp, buf.data() + c.off, c.W, c.H, c.step, rk.data(), rd.data(), cap, nullptr);
    bad += compare("oracle port", kps, desc, rk, rd, rn);
    if (rcreate && rextract) {
      void* r = rcreate(c.nf, c.sf, c.nl, 20, 7);
      rn = rextract(r, buf.data() + c.off, c.W, c.H, c.step, rk.data(), rd.data(), cap);
      bad += compare("the reference's own ORBextractor class (oracle/_ref)", kps, desc, rk, rd, rn);
      rdestroy(r);
      ++refChecked;
    }
    if (ex.GetLevels() != c.nl || (int)ex.GetScaleFactors().size() != c.nl || ex.GetScaleFactor() != c.sf) { std::printf("FAIL: getters\n"); ++bad; }
    if (desc.rows >= 2) {
      int d = myslam_b200::computeDistance(desc.row(0), desc.row(1)), w = 0;
      for (int b = 0; b < 32; ++b) w += __builtin_popcount(desc.ptr(0)[b] ^ desc.ptr(1)[b]);
      if (d != w) { std::printf("FAIL: computeDistance %d != %d\n", d, w); ++bad; }
    }
  }
  // empty image: silent return, outputs untouched (ORBextractor.cpp:1054-1055)
  {
    ORB_SLAM2::ORBextractor ex(500, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kps(3);
    cv::Mat desc;
    ex(cv::Mat(), cv::Mat(), kps, desc);
    if (kps.size() != 3) { std::printf("FAIL: empty image must leave the outputs untouched\n"); ++bad; }
  }
  std::printf("adapter_check: %s (%d cases, reference class compared in %d)\n", bad ? "FAILED" : "all equal", (int)(sizeof(cases) / sizeof(cases[0])), refChecked);
  return bad ? 1 : 0;
}
