// tests/tools/matcher_adapter_check.cpp -- compile/link check of include/orb_b200_matcher.hpp against stand-in Frame /
// KeyFrame / MapPoint types, and (with a GPU) a run-time comparison with a CPU statement of the reference's own loops
// (oracle/compat_myslam/myslam_stub.hpp, RefMatcher) on identical object graphs: the mappoints_ / mappointMatches vectors written by
// the two must be the same pointers, and the returned match counts equal.
//   g++ -std=c++11 -Ioracle/compat_myslam -Ioracle/compat -Iinclude tests/tools/matcher_adapter_check.cpp -Lvo_slam_test_b200/lib -lvoslam_b200 -o /tmp/mcheck
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>

#include "myslam_stub.hpp"
#include "orb_b200_frame.hpp"
#include "orb_b200_matcher.hpp"

using namespace myslam;

#ifdef REF_MATCHER
// The reference's OWN src/matcher.cpp, compiled in place against these stand-in types (oracle/Makefile ->
// oracle/_ref/libmatcherref.so, entry points in oracle/ref_matcher_wrap.cpp), plays the checker instead of RefMatcher.
int refm_searchByProjection_FF(float ratio, Frame* c, Frame* l, float radius, bool checkRot);
int refm_searchByProjection_FK(float ratio, Frame* c, KeyFrame* kf, float radius, float distTh, const std::set<MapPoint*>& found, bool checkRot);
int refm_searchByProjection_local(float ratio, Frame* f, const std::vector<MapPoint*>& mps, float th);
int refm_searchByProjection_sim3(float ratio, KeyFrame* kf, Sim3& Scw, std::vector<MapPoint*>& loopPts, std::vector<MapPoint*>& matchPts, int th);
int refm_searchByBoW_KF(float ratio, KeyFrame* kf, Frame* f, std::vector<MapPoint*>& out, bool checkRot);
int refm_searchByBoW_KK(float ratio, KeyFrame* k1, KeyFrame* k2, std::vector<MapPoint*>& out, bool checkRot);
int refm_searchBySim3(float ratio, KeyFrame* k1, KeyFrame* k2, std::vector<MapPoint*>& m12, Sim3& S12, float th);
int refm_searchForTriangulation(float ratio, KeyFrame* k1, KeyFrame* k2, std::vector<std::pair<int, int> >& idxs, Matrix3d& F12, bool checkRot);
int refm_fuseMapPoints(float ratio, KeyFrame* kf, std::vector<MapPoint*>& mps, const float& th);
int refm_fuseByPose(float ratio, KeyFrame* kf, Sim3& Scw, std::vector<MapPoint*>& loopPts, std::vector<MapPoint*>& replacePts, float th);
int refm_computeDistance(const cv::Mat& a, const cv::Mat& b);
struct Checker {
  float r;
  explicit Checker(float ratio) : r(ratio) {}
  static const char* name() { return "the reference's own src/matcher.cpp (compiled in place)"; }
  int searchByProjection(Frame* c, Frame* l, float radius, bool rot) { return refm_searchByProjection_FF(r, c, l, radius, rot); }
  int searchByProjection(Frame* c, KeyFrame* kf, float radius, float dt, const std::set<MapPoint*>& f, bool rot) { return refm_searchByProjection_FK(r, c, kf, radius, dt, f, rot); }
  int searchByProjection(Frame* f, const std::vector<MapPoint*>& mps, float th) { return refm_searchByProjection_local(r, f, mps, th); }
  int searchByProjection(KeyFrame* kf, Sim3& S, std::vector<MapPoint*>& lp, std::vector<MapPoint*>& mp, int th) { return refm_searchByProjection_sim3(r, kf, S, lp, mp, th); }
  int searchByBoW(KeyFrame* kf, Frame* f, KeyFrame* kf2, std::vector<MapPoint*>& out, bool rot) { return kf2 ? refm_searchByBoW_KK(r, kf, kf2, out, rot) : refm_searchByBoW_KF(r, kf, f, out, rot); }
  int searchBySim3(KeyFrame* k1, KeyFrame* k2, std::vector<MapPoint*>& m12, Sim3& S, float th) { return refm_searchBySim3(r, k1, k2, m12, S, th); }
  int searchForTriangulation(KeyFrame* k1, KeyFrame* k2, std::vector<std::pair<int, int> >& idxs, Matrix3d& F, bool rot) { return refm_searchForTriangulation(r, k1, k2, idxs, F, rot); }
  int fuseMapPoints(KeyFrame* kf, std::vector<MapPoint*>& mps, const float& th) { return refm_fuseMapPoints(r, kf, mps, th); }
  int fuseByPose(KeyFrame* kf, Sim3& S, std::vector<MapPoint*>& lp, std::vector<MapPoint*>& rp, float th) { return refm_fuseByPose(r, kf, S, lp, rp, th); }
  static int computeDistance(const cv::Mat& a, const cv::Mat& b) { return refm_computeDistance(a, b); }
};
#else
struct Checker : RefMatcher {
  explicit Checker(float ratio) : RefMatcher(ratio) {}
  static const char* name() { return "the object-walking restatement (RefMatcher)"; }
};
#endif

static unsigned g_seed = 1;
static unsigned rnd() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 8; }
static float urand() { return (float)(rnd() & 0xFFFFFF) / 16777216.0f; }

static void flipBits(uchar* d, int nflip) {
  for (int k = 0; k < nflip; ++k) { const unsigned b = rnd() % 256; d[b >> 3] ^= (uchar)(1u << (b & 7)); }
}

struct Scene {
  Camera cam;
  std::vector<MapPoint> points;         // storage; pointers into it are shared by every copy of the frames
  Frame cur, last;
  KeyFrame kf1, kf2;
  std::vector<MapPoint*> local;
};

static void fillFeatures(std::vector<cv::KeyPoint>& kps, cv::Mat& desc, std::vector<float>* uright, int n, int W, int H) {
  kps.resize(n);
  desc = cv::Mat(n, 32, CV_8UC1);
  if (uright) uright->resize(n);
  for (int i = 0; i < n; ++i) {
    kps[i] = cv::KeyPoint(urand() * W, urand() * H, 31.f, urand() * 360.f, 20.f + (float)(rnd() % 60), (int)(rnd() % 8), -1);
    for (int b = 0; b < 32; ++b) desc.data[(size_t)i * 32 + b] = (uchar)rnd();
    if (uright) (*uright)[i] = (rnd() % 3 == 0) ? -1.f : kps[i].pt.x - 30.f * urand();
  }
}

static void buildScene(Scene& s, unsigned seed, bool moveForward) {
  g_seed = seed;
  const int W = 640, H = 480, N = 1000, M = 900;
  s.cam.fx_ = 517.3f; s.cam.fy_ = 516.5f; s.cam.cx_ = 318.6f; s.cam.cy_ = 255.3f; s.cam.bf_ = 40.f; s.cam.b_ = 40.f / 517.3f;
  Frame& c = s.cur;
  c.camera_ = &s.cam;
  c.N_ = N;
  fillFeatures(c.unKeypoints_, c.descriptors_, &c.uRight_, N, W, H);
  c.xMin_ = 0; c.yMin_ = 0; c.xMax_ = W; c.yMax_ = H;
  c.gridPerPixelWidth_ = 64.f / (c.xMax_ - c.xMin_); c.gridPerPixelHeight_ = 48.f / (c.yMax_ - c.yMin_);
  c.scaleFactors_.resize(8); c.scaleFactors_[0] = 1.f;
  for (int l = 1; l < 8; ++l) c.scaleFactors_[l] = (float)(c.scaleFactors_[l - 1] * (double)1.2f);
  c.assignFeaturesToGrid();
  c.Tcw_ = SE3::rotY(0.02, 0.05, -0.01, 0.1);
  c.mappoints_.assign(N, nullptr);

  s.points.resize(M + 200);
  Frame& l = s.last;
  l.camera_ = &s.cam;
  l.Tcw_ = moveForward ? SE3::rotY(0.02, 0.05, -0.01, 0.5) : SE3::rotY(0.021, 0.06, -0.012, 0.11);
  l.N_ = M;
  l.unKeypoints_.resize(M);
  l.mappoints_.assign(M, nullptr);
  l.outliers_.assign(M, false);
  const SE3 Twc = c.Tcw_.inverse();
  for (int i = 0; i < M; ++i) {
    const int j = (int)(rnd() % N);                 // the current-frame feature this point should land near
    MapPoint& mp = s.points[i];
    const double z = (rnd() % 40 == 0) ? -1.0 : 1.0 + 6.0 * urand();
    const double u = c.unKeypoints_[j].pt.x + 8.0 * (urand() - 0.5) + ((rnd() % 25 == 0) ? 700.0 : 0.0);
    const double v = c.unKeypoints_[j].pt.y + 8.0 * (urand() - 0.5);
    mp.pos_ = Twc * Vector3d((u - s.cam.cx_) / s.cam.fx_ * z, (v - s.cam.cy_) / s.cam.fy_ * z, z);
    mp.descriptor_ = c.descriptors_.row(j).clone();
    flipBits(mp.descriptor_.data, (int)(rnd() % 70));
    mp.observe_cnt_ = (int)(rnd() % 3);             // 0: a temporal point that does not block later ones
    l.unKeypoints_[i] = c.unKeypoints_[j];
    l.unKeypoints_[i].octave = std::max(0, std::min(7, c.unKeypoints_[j].octave + (int)(rnd() % 3) - 1));
    l.unKeypoints_[i].angle = std::fmod(c.unKeypoints_[j].angle + ((rnd() % 6 == 0) ? 360.f * urand() : 3.f * urand()), 360.f);
    l.mappoints_[i] = (rnd() % 10 == 0) ? nullptr : &mp;
    l.outliers_[i] = rnd() % 12 == 0;
    // local-map view of the same point
    mp.badFlag_ = rnd() % 15 == 0;
    mp.trackInLocalMap_ = rnd() % 8 != 0;
    mp.trackScaleLevel_ = l.unKeypoints_[i].octave;
    mp.trackProj_u_ = (float)u; mp.trackProj_v_ = (float)v; mp.trackProj_uR_ = (float)u - 40.f / (float)std::fabs(z);
    mp.viewCos_ = 0.99f + 0.01f * urand();
    {                                                // scale-invariance range around the true distance (reloc search)
      const float d3 = (float)(mp.pos_ - Twc.translation()).norm();
      mp.maxDistance_ = d3 * (0.7f + 1.6f * urand());
      mp.minDistance_ = mp.maxDistance_ / 3.5831808f;
      const Vector3d ray = mp.pos_ - Twc.translation();      // mean viewing direction: mostly towards the camera, sometimes not
      const double nr = ray.norm() > 0 ? ray.norm() : 1.0;
      mp.normalVector_ = (rnd() % 9 == 0) ? Vector3d(1, 0, 0) : Vector3d(ray[0] / nr + 0.3 * (urand() - 0.5), ray[1] / nr, ray[2] / nr);
    }
    s.local.push_back(&mp);
  }
  for (int i = 0; i < N; ++i)                        // some features of the current frame already hold points
    if (rnd() % 9 == 0) { MapPoint& mp = s.points[M + (rnd() % 200)]; mp.observe_cnt_ = (int)(rnd() % 2); c.mappoints_[i] = &mp; }

  // two key frames + feature vectors (node = a coarse hash of the feature so that corresponding features share a node)
  KeyFrame* kfs[2] = {&s.kf1, &s.kf2};
  for (int k = 0; k < 2; ++k) {
    KeyFrame& kf = *kfs[k];
    kf.N_ = N;
    kf.unKeypoints_ = c.unKeypoints_;
    kf.descriptors_ = c.descriptors_.clone();
    kf.mappoints_.assign(N, nullptr);
    kf.scaleFactors_ = c.scaleFactors_;
    kf.camera_ = &s.cam;
    kf.uRight_ = c.uRight_;
    kf.xMin_ = c.xMin_; kf.xMax_ = c.xMax_; kf.yMin_ = c.yMin_; kf.yMax_ = c.yMax_;
    kf.gridPerPixelWidth_ = c.gridPerPixelWidth_; kf.gridPerPixelHeight_ = c.gridPerPixelHeight_;
    kf.assignFeaturesToGrid();
    for (int i = 0; i < N; ++i) {
      flipBits(kf.descriptors_.data + (size_t)i * 32, (int)(rnd() % 50));
      kf.unKeypoints_[i].angle = std::fmod(c.unKeypoints_[i].angle + ((rnd() % 5 == 0) ? 360.f * urand() : 4.f * urand()), 360.f);
      if (rnd() % 5 != 0) kf.mappoints_[i] = &s.points[rnd() % M];
      if (rnd() % 7 != 0) kf.featVec_[(unsigned)((i * 2654435761u) >> 26) + (rnd() % 11 == 0 ? 100u : 0u)].push_back((unsigned)i);
    }
  }
  for (int i = 0; i < N; ++i)
    if (rnd() % 7 != 0) c.featVec_[(unsigned)((i * 2654435761u) >> 26) + (rnd() % 13 == 0 ? 200u : 0u)].push_back((unsigned)i);
}

// Fingerprint of what the CHECKER wrote (indices into the scene's point storage): goes into the printed report so that
// tests/golden/matcher_reference_counts.txt carries the reference's assignments, not only its counts.
static unsigned long long fingerprint(const std::vector<MapPoint*>& v, const std::vector<MapPoint>& store) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < v.size(); ++i) { h ^= (unsigned long long)(v[i] ? (v[i] - &store[0]) + 1 : 0); h *= 1099511628211ull; }
  return h;
}
static int fails = 0;
static void expect(bool ok, const char* what) {
  std::printf("  %-58s %s\n", what, ok ? "same" : "DIFFERENT");
  if (!ok) ++fails;
}

int main(int argc, char** argv) {
  const unsigned seedBase = argc > 1 ? (unsigned)std::strtoul(argv[1], nullptr, 10) : 77u;   // other scenes for the CPU sweep
  int ndev = 0;
  orbx_device_count(&ndev);
  if (ndev == 0) { std::printf("matcher adapter links; no CUDA device -> compute skipped\n"); return 0; }
  typedef myslam_b200::MatcherT<Frame, KeyFrame, MapPoint> Matcher;
  std::printf("checker: %s\n", Checker::name());
  {  // the scenes are built with libm calls: a machine whose libm rounds these differently builds (slightly) different scenes
    volatile double x = 0.02, y = 0.021, z = 1.7;
    volatile float w = 2.5f;
    const double sig[4] = {std::cos(x), std::sin(y), std::log(z), (double)std::pow(1.2f, (float)w)};
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < 4; ++i) { unsigned long long b; std::memcpy(&b, &sig[i], 8); h ^= b; h *= 1099511628211ull; }
    std::printf("libm signature %016llx\n", h);
  }
  for (int round = 0; round < 4; ++round) {
    const bool fwd = round == 1, checkRot = round != 2;
    const float ratio = round == 3 ? 0.9f : 0.7f;
    Scene a, b;                                   // identical object graphs: a for the reference loops, b for the adapter
    buildScene(a, seedBase + round, fwd);
    buildScene(b, seedBase + round, fwd);
    Checker ref(ratio);
    Matcher gpu(ratio);
    std::printf("round %d (forward %d, checkRot %d, ratio %.2f)\n", round, (int)fwd, (int)checkRot, ratio);

    // 1. frame-to-frame projection search; compare by index into the scene's own point storage
    Frame ca = a.cur, cb = b.cur;
    int na = ref.searchByProjection(&ca, &a.last, 15.f, checkRot);
    int nb = gpu.searchByProjection(&cb, &b.last, 15.f, checkRot);
    bool same = na == nb;
    for (size_t i = 0; i < ca.mappoints_.size() && same; ++i) {
      const long ia = ca.mappoints_[i] ? (long)(ca.mappoints_[i] - &a.points[0]) : -1;
      const long ib = cb.mappoints_[i] ? (long)(cb.mappoints_[i] - &b.points[0]) : -1;
      same = ia == ib;
    }
    std::printf("  searchByProjection(Frame*,Frame*): %d matches [%016llx]\n", na, fingerprint(ca.mappoints_, a.points));
    expect(same && na > 50, "searchByProjection(Frame*, Frame*, radius, checkRot)");

    // 2. local-map projection search
    ca = a.cur; cb = b.cur;
    na = ref.searchByProjection(&ca, a.local, 1.0f + round);
    nb = gpu.searchByProjection(&cb, b.local, 1.0f + round);
    same = na == nb;
    for (size_t i = 0; i < ca.mappoints_.size() && same; ++i) {
      const long ia = ca.mappoints_[i] ? (long)(ca.mappoints_[i] - &a.points[0]) : -1;
      const long ib = cb.mappoints_[i] ? (long)(cb.mappoints_[i] - &b.points[0]) : -1;
      same = ia == ib;
    }
    std::printf("  searchByProjection(Frame*,vector<MapPoint*>): %d matches [%016llx]\n", na, fingerprint(ca.mappoints_, a.points));
    expect(same && na > 50, "searchByProjection(Frame*, vector<MapPoint*>&, thRadius)");

    // 2b. relocalisation projection search against a key frame (some of its points already "found")
    {
      KeyFrame ka = a.kf1, kb = b.kf1;
      for (size_t i = 0; i < ka.mappoints_.size(); ++i) {          // give the key frame the points that project near feature i
        ka.mappoints_[i] = i < a.last.mappoints_.size() ? a.last.mappoints_[i] : nullptr;
        kb.mappoints_[i] = i < b.last.mappoints_.size() ? b.last.mappoints_[i] : nullptr;
        if (i < a.last.unKeypoints_.size()) { ka.unKeypoints_[i].angle = a.last.unKeypoints_[i].angle; kb.unKeypoints_[i].angle = b.last.unKeypoints_[i].angle; }
      }
      std::set<MapPoint*> fa, fb;
      for (size_t i = 0; i < 900; i += 17) { fa.insert(&a.points[i]); fb.insert(&b.points[i]); }
      ca = a.cur; cb = b.cur;
      na = ref.searchByProjection(&ca, &ka, 10.f + 5.f * round, round == 1 ? 64.f : 100.f, fa, checkRot);
      nb = gpu.searchByProjection(&cb, &kb, 10.f + 5.f * round, round == 1 ? 64.f : 100.f, fb, checkRot);
      same = na == nb;
      for (size_t i = 0; i < ca.mappoints_.size() && same; ++i) {
        const long ia = ca.mappoints_[i] ? (long)(ca.mappoints_[i] - &a.points[0]) : -1;
        const long ib = cb.mappoints_[i] ? (long)(cb.mappoints_[i] - &b.points[0]) : -1;
        same = ia == ib;
      }
      std::printf("  searchByProjection(Frame*,KeyFrame*): %d matches [%016llx]\n", na, fingerprint(ca.mappoints_, a.points));
      expect(same && na > 50, "searchByProjection(Frame*, KeyFrame*, radius, distTh, found, rot)");
    }

    // 2c. loop-closure projection search through a Sim3 (key frame pose = the current frame's pose, scale != 1)
    {
      const double sc = 1.0 + 0.1 * round;
      SE3 T = a.cur.Tcw_;
      for (int k = 0; k < 3; ++k) T.t[k] *= sc;             // Sim3(s, R, s*t): x -> s (R x + t) projects like the SE3
      Sim3 Sa(T, sc), Sb(T, sc);
      std::vector<MapPoint*> la = a.local, lb = b.local, ma2(a.kf1.N_, nullptr), mb2(b.kf1.N_, nullptr);
      for (size_t i = 0; i < la.size(); i += 13) { la[i] = nullptr; lb[i] = nullptr; }
      for (size_t i = 0; i < ma2.size(); i += 7) { ma2[i] = &a.points[(i * 31) % 900]; mb2[i] = &b.points[(i * 31) % 900]; }
      na = ref.searchByProjection(&a.kf1, Sa, la, ma2, 10);
      nb = gpu.searchByProjection(&b.kf1, Sb, lb, mb2, 10);
      same = na == nb;
      for (size_t i = 0; i < ma2.size() && same; ++i)
        same = (ma2[i] ? (long)(ma2[i] - &a.points[0]) : -1) == (mb2[i] ? (long)(mb2[i] - &b.points[0]) : -1);
      std::printf("  searchByProjection(KeyFrame*,Sim3&): %d matches [%016llx]\n", na, fingerprint(ma2, a.points));
      expect(same && na > 10, "searchByProjection(KeyFrame*, Sim3&, loopPts, matchPts, th)");
    }

    // 2d. searchBySim3 between two key frames at the same pose (S12 close to identity, scale != 1 folded into the points)
    {
      KeyFrame ka1 = a.kf1, ka2 = a.kf2, kb1 = b.kf1, kb2 = b.kf2;
      KeyFrame* ks[4] = {&ka1, &ka2, &kb1, &kb2};
      Scene* sc[4] = {&a, &a, &b, &b};
      for (int k4 = 0; k4 < 4; ++k4) {
        const int k = k4 & 1;                                      // key frame 1 / 2 of either scene
        ks[k4]->Tcw_ = sc[k4]->cur.Tcw_;
        ks[k4]->descriptors_ = ks[k4]->descriptors_.clone();       // cv::Mat copies are shallow
        for (size_t i = 0; i < ks[k4]->mappoints_.size(); ++i)     // points that really project near feature i of both key frames
          ks[k4]->mappoints_[i] = (i < 900 && (i * 7 + k) % 11 != 0) ? &sc[k4]->points[i] : nullptr;
        for (size_t i = 0; i < 900; ++i) {                         // ... which means feature i must sit where point i projects
          const MapPoint& mp = sc[k4]->points[i];
          ks[k4]->unKeypoints_[i].pt.x = mp.trackProj_u_ + (float)((i * 13 + k) % 5) - 2.f;
          ks[k4]->unKeypoints_[i].pt.y = mp.trackProj_v_ + (float)((i * 17 + k) % 5) - 2.f;
          ks[k4]->unKeypoints_[i].octave = mp.trackScaleLevel_;
          std::memcpy(ks[k4]->descriptors_.data + i * 32, mp.descriptor_.data, 32);
          ks[k4]->descriptors_.data[i * 32 + (i % 32)] ^= (uchar)(1u << k);
        }
        ks[k4]->assignFeaturesToGrid();
      }
      for (int i = 0; i < 900; ++i) {                              // predictScale must land on the feature's level
        for (int k = 0; k < 2; ++k) {
          MapPoint& mp = (k ? b : a).points[i];
          const float d3 = (float)(( (k ? b : a).cur.Tcw_ * mp.pos_)).norm();
          mp.maxDistance_ = d3 * std::pow(1.2f, (float)mp.trackScaleLevel_ - 0.5f);
          mp.minDistance_ = mp.maxDistance_ / 3.5831808f;
          if (i % 50 == 0) mp.observedKFs_[k ? &kb2 : &ka2] = (i * 3) % 1000;
        }
      }
      SE3 I; I.t[0] = 0.002; I.t[2] = -0.001;
      Sim3 Sa(I, 1.0 + 0.002 * round), Sb(I, 1.0 + 0.002 * round);
      std::vector<MapPoint*> m12a(ka1.N_, nullptr), m12b(kb1.N_, nullptr);
      for (size_t i = 0; i < 900; i += 50) { m12a[i] = &a.points[i]; m12b[i] = &b.points[i]; }
      na = ref.searchBySim3(&ka1, &ka2, m12a, Sa, 7.5f);
      nb = gpu.searchBySim3(&kb1, &kb2, m12b, Sb, 7.5f);
      same = na == nb;
      for (size_t i = 0; i < m12a.size() && same; ++i)
        same = (m12a[i] ? (long)(m12a[i] - &a.points[0]) : -1) == (m12b[i] ? (long)(m12b[i] - &b.points[0]) : -1);
      std::printf("  searchBySim3(KeyFrame*,KeyFrame*): %d matches [%016llx]\n", na, fingerprint(m12a, a.points));
      expect(same && na > 50, "searchBySim3(KeyFrame*, KeyFrame*, matches12, S12, th)");
      for (int i = 0; i < 900; ++i) { a.points[i].observedKFs_.clear(); b.points[i].observedKFs_.clear(); }   // ka2 / kb2 die here
    }

    // 3. BoW search key frame -> frame
    std::vector<MapPoint*> ma, mb;
    na = ref.searchByBoW(&a.kf1, &a.cur, nullptr, ma, checkRot);
    nb = gpu.searchByBoW(&b.kf1, &b.cur, mb, checkRot);
    same = na == nb && ma.size() == mb.size();
    for (size_t i = 0; i < ma.size() && same; ++i)
      same = (ma[i] ? (long)(ma[i] - &a.points[0]) : -1) == (mb[i] ? (long)(mb[i] - &b.points[0]) : -1);
    std::printf("  searchByBoW(KeyFrame*,Frame*): %d matches [%016llx]\n", na, fingerprint(ma, a.points));
    expect(same && na > 50, "searchByBoW(KeyFrame*, Frame*, matches, checkRot)");

    // 4. BoW search key frame -> key frame
    na = ref.searchByBoW(&a.kf1, nullptr, &a.kf2, ma, checkRot);
    nb = gpu.searchByBoW(&b.kf1, &b.kf2, mb, checkRot);
    same = na == nb && ma.size() == mb.size();
    for (size_t i = 0; i < ma.size() && same; ++i)
      same = (ma[i] ? (long)(ma[i] - &a.points[0]) : -1) == (mb[i] ? (long)(mb[i] - &b.points[0]) : -1);
    std::printf("  searchByBoW(KeyFrame*,KeyFrame*): %d matches [%016llx]\n", na, fingerprint(ma, a.points));
    expect(same && na > 20, "searchByBoW(KeyFrame*, KeyFrame*, matches, checkRot)");

    // 4b. searchForTriangulation: key frame 2 a little to the right of key frame 1, F12 of that stereo-like pair
    {
      KeyFrame ka1 = a.kf1, ka2 = a.kf2, kb1 = b.kf1, kb2 = b.kf2;
      KeyFrame* ks[4] = {&ka1, &ka2, &kb1, &kb2};
      for (int k4 = 0; k4 < 4; ++k4) {
        KeyFrame& kf = *ks[k4];
        kf.Tcw_ = SE3();
        if (k4 & 1) {
          kf.Tcw_.t[0] = -0.3;                              // camera 2 sits at x = +0.3
          for (int i = 0; i < 1000; ++i) kf.unKeypoints_[i].pt.x -= 25.f + (float)(i % 7);   // disparity; same row => on the epipolar line
        }
        for (int i = 0; i < 1000; ++i) {
          if (i % 9 == 0) kf.unKeypoints_[i].pt.y += 9.f;  // off the epipolar line
          if ((i + (k4 & 1)) % 4 != 0) kf.mappoints_[i] = nullptr;   // most features have no map point yet
          if (i % 2 == (k4 & 1)) kf.uRight_[i] = -1.f;
        }
      }
      // F12 with p1^T F12 p2 = 0 for pure x-translation and equal intrinsics: rows of image 1 map to rows of image 2
      Matrix3d F; for (int i = 0; i < 9; ++i) F.m[i] = 0;
      F.m[1 * 3 + 2] = 1.0; F.m[2 * 3 + 1] = -1.0;
      std::vector<std::pair<int, int> > pa, pb;
      na = ref.searchForTriangulation(&ka1, &ka2, pa, F, checkRot);
      nb = gpu.searchForTriangulation(&kb1, &kb2, pb, F, checkRot);
      same = na == nb && pa == pb;
      unsigned long long ht = 1469598103934665603ull;
      for (size_t i = 0; i < pa.size(); ++i) { ht ^= (unsigned long long)pa[i].first * 4096u + (unsigned long long)pa[i].second; ht *= 1099511628211ull; }
      std::printf("  searchForTriangulation: %d matches [%016llx]\n", na, ht);
      expect(same && na > 60, "searchForTriangulation(KeyFrame*, KeyFrame*, idxs, F12, checkRot)");
    }

    // 6. the two fuse functions (last: they rewire the object graph).  Key-frame features sit where the points project.
    for (int pass = 0; pass < 2; ++pass) {
      KeyFrame ka = a.kf1, kb = b.kf1;
      KeyFrame* ks[2] = {&ka, &kb};
      Scene* sc[2] = {&a, &b};
      for (int k = 0; k < 2; ++k) {
        KeyFrame& kf = *ks[k];
        Scene& s = *sc[k];
        kf.Tcw_ = s.cur.Tcw_;
        kf.descriptors_ = kf.descriptors_.clone();
        for (int i = 0; i < 1000; ++i) {
          kf.mappoints_[i] = nullptr;
          if (i >= 900) continue;
          MapPoint& mp = s.points[i];
          mp.badFlag_ = i % 15 == 3; mp.observedKFs_.clear(); mp.observe_cnt_ = i % 4;
          kf.unKeypoints_[i].pt.x = mp.trackProj_u_ + (float)((i * 13) % 5) * 0.5f - 1.f;
          kf.unKeypoints_[i].pt.y = mp.trackProj_v_ + (float)((i * 17) % 5) * 0.5f - 1.f;
          kf.unKeypoints_[i].octave = mp.trackScaleLevel_;
          kf.uRight_[i] = (i % 3 == 0) ? -1.f : mp.trackProj_uR_ + 0.3f;
          std::memcpy(kf.descriptors_.data + i * 32, mp.descriptor_.data, 32);
          kf.descriptors_.data[i * 32 + (i % 32)] ^= 1;
          const float d3 = (float)((s.cur.Tcw_ * mp.pos_)).norm();
          mp.maxDistance_ = d3 * std::pow(1.2f, (float)mp.trackScaleLevel_ - 0.5f);
          mp.minDistance_ = mp.maxDistance_ / 3.5831808f;
          if (i % 5 == 1) {                                  // the feature already holds another (spare) point
            MapPoint& org = s.points[900 + (i % 200)];
            if (!org.beObserved(&kf)) { org.badFlag_ = i % 35 == 1; org.observe_cnt_ = i % 3; org.addObservation(&kf, i); kf.mappoints_[i] = &org; }
          }
          if (i % 19 == 2) mp.addObservation(&kf, (i + 500) % 1000);   // already observed by the key frame elsewhere
        }
        kf.assignFeaturesToGrid();
      }
      std::vector<MapPoint*> la = a.local, lb = b.local, ra(la.size(), nullptr), rb(lb.size(), nullptr);
      for (size_t i = 0; i < la.size(); i += 23) { la[i] = nullptr; lb[i] = nullptr; }
      for (size_t i = 5; i < la.size(); i += 40) { la[i] = la[i - 1]; lb[i] = lb[i - 1]; }   // duplicates in the list
      if (pass == 0) {
        na = ref.fuseMapPoints(&ka, la, 3.0f);
        nb = gpu.fuseMapPoints(&kb, lb, 3.0f);
      } else {
        Sim3 Sa(a.cur.Tcw_, 1.0), Sb(b.cur.Tcw_, 1.0);
        na = ref.fuseByPose(&ka, Sa, la, ra, 4.0f);
        nb = gpu.fuseByPose(&kb, Sb, lb, rb, 4.0f);
      }
      same = na == nb;
      for (size_t i = 0; i < ka.mappoints_.size() && same; ++i)
        same = (ka.mappoints_[i] ? (long)(ka.mappoints_[i] - &a.points[0]) : -1) == (kb.mappoints_[i] ? (long)(kb.mappoints_[i] - &b.points[0]) : -1);
      for (size_t i = 0; i < ra.size() && same; ++i)
        same = (ra[i] ? (long)(ra[i] - &a.points[0]) : -1) == (rb[i] ? (long)(rb[i] - &b.points[0]) : -1);
      int nbad = 0;
      for (size_t i = 0; i < a.points.size() && same; ++i) {
        same = a.points[i].badFlag_ == b.points[i].badFlag_ && a.points[i].observe_cnt_ == b.points[i].observe_cnt_ &&
               a.points[i].observedKFs_.size() == b.points[i].observedKFs_.size();
        nbad += a.points[i].badFlag_;
      }
      std::printf("  %s: %d fused (%d bad points afterwards) [%016llx %016llx]\n", pass ? "fuseByPose" : "fuseMapPoints", na, nbad, fingerprint(ka.mappoints_, a.points), fingerprint(ra, a.points));
      expect(same && na > 50, pass ? "fuseByPose(KeyFrame*, Sim3&, loopPts, replacePts, th)" : "fuseMapPoints(KeyFrame*, mappoints, threshold)");
      for (size_t i = 0; i < a.points.size(); ++i) { a.points[i].observedKFs_.clear(); b.points[i].observedKFs_.clear(); }
    }

    // 5. computeDistance
    const int d0 = Checker::computeDistance(a.cur.descriptors_.row(3), a.kf1.descriptors_.row(3));
    const int d1 = Matcher::computeDistance(b.cur.descriptors_.row(3), b.kf1.descriptors_.row(3));
    expect(d0 == d1, "computeDistance(Mat, Mat)");
  }
  // The same tracking-thread searches with the current frame RESIDENT on the device (myslam_b200::makeResident registers an
  // orbx_frame_t for the Frame*; the adapter then sends only the projected points and the occupancy flags per search).
  {
    Scene a, b;
    buildScene(a, seedBase + 40, false);
    buildScene(b, seedBase + 40, false);
    Checker ref(0.7f);
    Matcher gpu(0.7f);
    std::printf("resident frame (orbx_frame_t)\n");
    auto samePtrs = [&](const Frame& x, const Frame& y) {
      for (size_t i = 0; i < x.mappoints_.size(); ++i)
        if ((x.mappoints_[i] ? (long)(x.mappoints_[i] - &a.points[0]) : -1) != (y.mappoints_[i] ? (long)(y.mappoints_[i] - &b.points[0]) : -1)) return false;
      return true;
    };
    Frame ca = a.cur, cb = b.cur;
    myslam_b200::makeResident(&cb);
    int na = ref.searchByProjection(&ca, &a.last, 15.f, true), nb = gpu.searchByProjection(&cb, &b.last, 15.f, true);
    expect(na == nb && na > 50 && samePtrs(ca, cb), "resident: searchByProjection(Frame*, Frame*)");
    ca = a.cur; cb = b.cur;                                        // same object, same features: the handle stays valid
    na = ref.searchByProjection(&ca, a.local, 2.0f); nb = gpu.searchByProjection(&cb, b.local, 2.0f);
    expect(na == nb && na > 50 && samePtrs(ca, cb), "resident: searchByProjection(Frame*, local map)");
    {
      KeyFrame ka = a.kf1, kb = b.kf1;
      for (size_t i = 0; i < ka.mappoints_.size(); ++i) {
        ka.mappoints_[i] = i < a.last.mappoints_.size() ? a.last.mappoints_[i] : nullptr;
        kb.mappoints_[i] = i < b.last.mappoints_.size() ? b.last.mappoints_[i] : nullptr;
        if (i < a.last.unKeypoints_.size()) { ka.unKeypoints_[i].angle = a.last.unKeypoints_[i].angle; kb.unKeypoints_[i].angle = b.last.unKeypoints_[i].angle; }
      }
      std::set<MapPoint*> fa, fb;
      for (size_t i = 0; i < 900; i += 17) { fa.insert(&a.points[i]); fb.insert(&b.points[i]); }
      ca = a.cur; cb = b.cur;
      na = ref.searchByProjection(&ca, &ka, 15.f, 100.f, fa, true); nb = gpu.searchByProjection(&cb, &kb, 15.f, 100.f, fb, true);
      expect(na == nb && na > 50 && samePtrs(ca, cb), "resident: searchByProjection(Frame*, KeyFrame*)");
    }
    std::vector<MapPoint*> ma, mb;
    cb = b.cur;
    na = ref.searchByBoW(&a.kf1, &a.cur, nullptr, ma, true); nb = gpu.searchByBoW(&b.kf1, &cb, mb, true);
    bool same = na == nb && ma.size() == mb.size();
    for (size_t i = 0; i < ma.size() && same; ++i)
      same = (ma[i] ? (long)(ma[i] - &a.points[0]) : -1) == (mb[i] ? (long)(mb[i] - &b.points[0]) : -1);
    expect(same && na > 50, "resident: searchByBoW(KeyFrame*, Frame*)");
    myslam_b200::dropResident(&cb);
  }

  // Frame::Frame after the extractor call (frame.cpp:29-31): undistortKeyPoints + findDepth + assignFeaturesToGrid
  for (int variant = 0; variant < 3; ++variant) {
    Scene a, b;
    buildScene(a, 900u + variant, false);
    buildScene(b, 900u + variant, false);
    const float dist[5] = {variant == 2 ? 0.f : 0.2624f, -0.9531f, -0.0054f, 0.0026f, 1.1633f};   // TUM fr1-like; variant 2: k1 == 0 copies
    a.cam.setIntrinsics(dist, variant == 1 ? 5 : 4);
    b.cam.setIntrinsics(dist, variant == 1 ? 5 : 4);
    std::vector<float> depthStore(480 * 640);
    cv::Mat depth(480, 640, CV_32F, depthStore.data(), 640 * sizeof(float));   // CV_32F 480 x 640 over external memory
    g_seed = 31u + variant;
    for (int y = 0; y < 480; ++y)
      for (int x = 0; x < 640; ++x) depth.at<float>(y, x) = (rnd() % 6 == 0) ? 0.f : 0.5f + 7.f * urand();
    Frame* fr[2] = {&a.cur, &b.cur};
    for (int k = 0; k < 2; ++k) {
      fr[k]->keypoints_ = fr[k]->unKeypoints_;
      for (size_t i = 0; i < fr[k]->keypoints_.size(); ++i) {             // keep the depth lookup inside the image
        fr[k]->keypoints_[i].pt.x = std::min(fr[k]->keypoints_[i].pt.x, 639.f);
        fr[k]->keypoints_[i].pt.y = std::min(fr[k]->keypoints_[i].pt.y, 479.f);
      }
      fr[k]->unKeypoints_.clear(); fr[k]->uRight_.clear(); fr[k]->depth_.clear();
    }
    a.cur.undistortKeyPoints(); a.cur.findDepth(depth); a.cur.assignFeaturesToGrid();
    myslam_b200::finishFrame(&b.cur, depth);
    bool same = a.cur.unKeypoints_.size() == b.cur.unKeypoints_.size() &&
                std::memcmp(a.cur.unKeypoints_.data(), b.cur.unKeypoints_.data(), a.cur.unKeypoints_.size() * sizeof(cv::KeyPoint)) == 0 &&
                a.cur.uRight_ == b.cur.uRight_ && a.cur.depth_ == b.cur.depth_;
    int moved = 0;
    for (size_t i = 0; i < a.cur.keypoints_.size(); ++i) moved += a.cur.keypoints_[i].pt.x != a.cur.unKeypoints_[i].pt.x;
    for (int ix = 0; ix < 64 && same; ++ix)
      for (int iy = 0; iy < 48 && same; ++iy) same = a.cur.gridKeypoints_[ix][iy] == b.cur.gridKeypoints_[ix][iy];
    std::printf("finishFrame variant %d: %d of %zu keypoints moved by the undistortion\n", variant, moved, a.cur.keypoints_.size());
    expect(same && (variant == 2 ? moved == 0 : moved > 900), "Frame::Frame post-processing (undistort, depth, grid)");
  }

  // Frame::Frame as ONE call (myslam_b200::constructFrame: extractor + undistort + depth + grid on the device, frame left
  // resident) against the step-by-step path: orbx_extract, then the reference's own three member functions on the host.
  {
    orbx_params prm = {1000, 1.2f, 8, 20, 7, 0};
    orbx_handle ex = nullptr;
    if (orbx_create(&prm, &ex) == ORBX_OK) {
      Scene a, b;
      buildScene(a, 4711u, false);
      buildScene(b, 4711u, false);
      const float dist[5] = {0.2624f, -0.9531f, -0.0054f, 0.0026f, 1.1633f};
      a.cam.setIntrinsics(dist, 5); b.cam.setIntrinsics(dist, 5);
      std::vector<uchar> pix(480 * 640);
      g_seed = 99u;
      for (int y = 0; y < 480; ++y)
        for (int x = 0; x < 640; ++x) pix[(size_t)y * 640 + x] = (uchar)((((x / 24) + (y / 24)) & 1) * 120 + 60 + (rnd() % 8) + ((x * 7 + y * 3) % 23));
      cv::Mat gray(480, 640, CV_8UC1, pix.data(), 640);
      std::vector<float> depthStore(480 * 640);
      cv::Mat depth(480, 640, CV_32F, depthStore.data(), 640 * sizeof(float));
      for (int y = 0; y < 480; ++y)
        for (int x = 0; x < 640; ++x) depth.at<float>(y, x) = (rnd() % 6 == 0) ? 0.f : 0.5f + 7.f * urand();
      // step by step (reference members on the host)
      int cap = 0, n = 0;
      orbx_max_keypoints(ex, &cap);
      Frame& fa = a.cur;
      fa.keypoints_.resize(cap);
      std::vector<uchar> dsc((size_t)cap * 32);
      const int rc = orbx_extract(ex, pix.data(), 640, 480, 640, reinterpret_cast<orbx_keypoint*>(fa.keypoints_.data()), dsc.data(), cap, &n);
      fa.keypoints_.resize(n); fa.N_ = n;
      fa.descriptors_ = cv::Mat(n, 32, CV_8UC1);
      std::memcpy(fa.descriptors_.data, dsc.data(), (size_t)n * 32);
      fa.unKeypoints_.clear(); fa.uRight_.clear(); fa.depth_.clear();
      fa.undistortKeyPoints(); fa.findDepth(depth); fa.assignFeaturesToGrid();
      // one call
      Frame& fb = b.cur;
      myslam_b200::constructFrame(&fb, ex, gray, depth);
      bool same = rc == ORBX_OK && n > 900 && fb.N_ == fa.N_ && fa.keypoints_.size() == fb.keypoints_.size() &&
                  std::memcmp(fa.keypoints_.data(), fb.keypoints_.data(), (size_t)n * sizeof(cv::KeyPoint)) == 0 &&
                  std::memcmp(fa.unKeypoints_.data(), fb.unKeypoints_.data(), (size_t)n * sizeof(cv::KeyPoint)) == 0 &&
                  std::memcmp(fa.descriptors_.data, fb.descriptors_.data, (size_t)n * 32) == 0 && fa.uRight_ == fb.uRight_ && fa.depth_ == fb.depth_;
      for (int ix = 0; ix < 64 && same; ++ix)
        for (int iy = 0; iy < 48 && same; ++iy) same = fa.gridKeypoints_[ix][iy] == fb.gridKeypoints_[ix][iy];
      int withDepth = 0;
      for (int i = 0; i < n; ++i) withDepth += fb.uRight_[i] > 0;
      std::printf("constructFrame: %d keypoints, %d with depth\n", n, withDepth);
      expect(same && withDepth > 500, "Frame::Frame in one call (constructFrame) == extractor + reference members");
      // the frame is resident: searches against it take the handle path and must equal the reference loops on the host copy
      fa.mappoints_.assign(n, nullptr); fb.mappoints_.assign(n, nullptr);
      Scene* sc[2] = {&a, &b};
      for (int k = 0; k < 2; ++k) {                                // re-aim the scene's map points at the extracted features
        Scene& s = *sc[k];
        Frame& f = k ? fb : fa;
        g_seed = 515u;
        const SE3 Twc = f.Tcw_.inverse();
        for (size_t i = 0; i < s.last.mappoints_.size(); ++i) {
          const int j = (int)(rnd() % n);
          MapPoint& mp = s.points[i];
          const double z = 1.0 + 6.0 * urand();
          const double u = f.unKeypoints_[j].pt.x + 6.0 * (urand() - 0.5), v = f.unKeypoints_[j].pt.y + 6.0 * (urand() - 0.5);
          mp.pos_ = Twc * Vector3d((u - s.cam.cx_) / s.cam.fx_ * z, (v - s.cam.cy_) / s.cam.fy_ * z, z);
          mp.descriptor_ = f.descriptors_.row(j).clone();
          flipBits(mp.descriptor_.data, (int)(rnd() % 60));
          s.last.unKeypoints_[i] = f.unKeypoints_[j];
          mp.trackProj_u_ = (float)u; mp.trackProj_v_ = (float)v; mp.trackProj_uR_ = (float)u - 40.f / (float)z;
          mp.trackScaleLevel_ = f.unKeypoints_[j].octave;
        }
      }
      Checker ref(0.7f);
      Matcher gpu(0.7f);
      int na = ref.searchByProjection(&fa, &a.last, 15.f, true), nb = gpu.searchByProjection(&fb, &b.last, 15.f, true);
      same = na == nb;
      for (int i = 0; i < n && same; ++i)
        same = (fa.mappoints_[i] ? (long)(fa.mappoints_[i] - &a.points[0]) : -1) == (fb.mappoints_[i] ? (long)(fb.mappoints_[i] - &b.points[0]) : -1);
      std::printf("constructFrame + searchByProjection(Frame*,Frame*) on the resident frame: %d accepted\n", na);
      expect(same && na > 100, "resident frame from constructFrame: searchByProjection(Frame*, Frame*)");
      fa.mappoints_.assign(n, nullptr); fb.mappoints_.assign(n, nullptr);
      na = ref.searchByProjection(&fa, a.local, 2.0f); nb = gpu.searchByProjection(&fb, b.local, 2.0f);
      same = na == nb;
      for (int i = 0; i < n && same; ++i)
        same = (fa.mappoints_[i] ? (long)(fa.mappoints_[i] - &a.points[0]) : -1) == (fb.mappoints_[i] ? (long)(fb.mappoints_[i] - &b.points[0]) : -1);
      expect(same && na > 100, "resident frame from constructFrame: searchByProjection(Frame*, local map)");
      myslam_b200::dropResident(&fb);
      orbx_destroy(ex);
    } else {
      std::printf("constructFrame: no extractor behind this C ABI (CPU port build) -> skipped\n");
    }
  }

  // MapPoint::computeDescriptor for many points at once
  {
    Scene a, b;
    buildScene(a, 1234u, false);
    buildScene(b, 1234u, false);
    std::vector<KeyFrame> kfa(12), kfb(12);
    g_seed = 99u;
    for (int k = 0; k < 12; ++k) {
      kfa[k] = a.kf1; kfa[k].descriptors_ = a.kf1.descriptors_.clone();
      for (int i = 0; i < 1000; ++i) flipBits(kfa[k].descriptors_.data + (size_t)i * 32, (int)(rnd() % 40));
      kfa[k].bad_kf_ = k == 5;
      kfb[k] = b.kf1; kfb[k].descriptors_ = kfa[k].descriptors_.clone(); kfb[k].bad_kf_ = kfa[k].bad_kf_;
    }
    std::vector<MapPoint*> pa, pb;
    for (int i = 0; i < 900; ++i) {
      MapPoint& ma = a.points[i]; MapPoint& mb = b.points[i];
      ma.observedKFs_.clear(); mb.observedKFs_.clear();
      const int nobs = (i % 17 == 0) ? 0 : 1 + (int)(rnd() % 12);          // some points without observations
      for (int o = 0; o < nobs; ++o) {
        const int k = (int)(rnd() % 12);
        const size_t idx = (i % 3 == 0) ? (size_t)i : (size_t)(rnd() % 1000);
        ma.observedKFs_[&kfa[k]] = idx; mb.observedKFs_[&kfb[k]] = idx;
      }
      if (i % 29 == 0) { ma.observedKFs_.clear(); mb.observedKFs_.clear(); ma.observedKFs_[&kfa[5]] = 7; mb.observedKFs_[&kfb[5]] = 7; }   // only a bad key frame
      pa.push_back(&ma); pb.push_back(&mb);
    }
    pa.push_back(nullptr); pb.push_back(nullptr);
    for (size_t i = 0; i + 1 < pa.size(); ++i) refComputeDescriptor(pa[i]);
    myslam_b200::computeDescriptors(pb);
    bool same = true;
    int changed = 0;
    Scene c; buildScene(c, 1234u, false);                                  // untouched copy: how many descriptors moved at all
    for (int i = 0; i < 900 && same; ++i) {
      same = std::memcmp(a.points[i].descriptor_.data, b.points[i].descriptor_.data, 32) == 0;
      changed += std::memcmp(a.points[i].descriptor_.data, c.points[i].descriptor_.data, 32) != 0;
    }
    std::printf("computeDescriptors: %d of 900 map points got a new representative descriptor\n", changed);
    expect(same && changed > 500, "MapPoint::computeDescriptor, batched (medoid by median Hamming distance)");
    for (int i = 0; i < 900; ++i) { a.points[i].observedKFs_.clear(); b.points[i].observedKFs_.clear(); }
  }

  // Empty inputs: the reference's loops are no-ops; the adapter must return 0 without touching the device or the objects.
  {
    Scene a;
    buildScene(a, 5u, false);
    typedef myslam_b200::MatcherT<Frame, KeyFrame, MapPoint> M;
    M gpu(0.7f);
    Checker ref(0.7f);
    Frame none; none.camera_ = &a.cam; none.scaleFactors_ = a.cur.scaleFactors_; none.xMax_ = 640; none.yMax_ = 480;
    none.gridPerPixelWidth_ = 0.1f; none.gridPerPixelHeight_ = 0.1f;
    KeyFrame knone; knone.camera_ = &a.cam; knone.scaleFactors_ = a.cur.scaleFactors_; knone.assignFeaturesToGrid();
    std::vector<MapPoint*> empty, out1, out2;
    std::set<MapPoint*> found;
    Sim3 S(a.cur.Tcw_, 1.0);
    Matrix3d F; for (int i = 0; i < 9; ++i) F.m[i] = 0;
    std::vector<std::pair<int, int> > idxs(3, std::make_pair(1, 2));
    Frame c1 = a.cur, c2 = a.cur;
    bool ok = gpu.searchByProjection(&c1, &none, 15.f, true) == 0 && ref.searchByProjection(&c2, &none, 15.f, true) == 0;   // no map points
    ok = ok && gpu.searchByProjection(&none, &a.last, 15.f, true) == 0;                                                       // no features
    ok = ok && gpu.searchByProjection(&c1, empty, 1.f) == 0 && gpu.searchByProjection(&none, a.local, 1.f) == 0;
    ok = ok && gpu.searchByProjection(&c1, &knone, 10.f, 100.f, found, true) == 0 && gpu.searchByProjection(&none, &a.kf1, 10.f, 100.f, found, true) == 0;
    ok = ok && gpu.searchByProjection(&knone, S, a.local, empty, 10) == 0 && gpu.searchByProjection(&a.kf1, S, empty, out1, 10) == 0;
    ok = ok && gpu.searchByBoW(&knone, &c1, out1, true) == 0 && out1.size() == c1.N_;
    ok = ok && gpu.searchByBoW(&a.kf1, &none, out1, true) == 0 && out1.empty();
    ok = ok && gpu.searchByBoW(&a.kf1, &knone, out2, true) == 0 && out2.size() == a.kf1.N_;
    ok = ok && gpu.searchBySim3(&knone, &a.kf1, empty, S, 7.5f) == 0;
    ok = ok && gpu.searchForTriangulation(&knone, &a.kf1, idxs, F, true) == 0 && idxs.empty();
    ok = ok && gpu.fuseMapPoints(&a.kf1, empty, 3.f) == 0 && gpu.fuseMapPoints(&knone, a.local, 3.f) == 0;
    ok = ok && gpu.fuseByPose(&a.kf1, S, empty, out1, 4.f) == 0 && gpu.fuseByPose(&knone, S, a.local, out2, 4.f) == 0;
    for (size_t i = 0; i < c1.mappoints_.size() && ok; ++i) ok = c1.mappoints_[i] == a.cur.mappoints_[i];
    std::printf("empty inputs\n");
    expect(ok, "every entry point with empty point lists / featureless frames");
  }
  // Timing of the tracking-thread calls through the adapter (host objects in, host objects out, synchronous), next to the
  // object-walking CPU statement on this machine's host (one core).  Informative only: nothing is asserted on it.
  {
    Scene a;
    buildScene(a, 4242u, false);
    typedef myslam_b200::MatcherT<Frame, KeyFrame, MapPoint> M;
    M gpu(0.7f);
    Checker ref(0.7f);
    const int reps = 20;
    double tr[3] = {0, 0, 0}, tg[3] = {0, 0, 0};
    std::vector<MapPoint*> out;
    for (int r = -2; r < reps; ++r) {                      // two warm-up rounds
      for (int which = 0; which < 2; ++which) {
        for (int k = 0; k < 3; ++k) {
          Frame c = a.cur;                                 // fresh copy: the searches write mappoints_
          const auto t0 = std::chrono::steady_clock::now();
          if (k == 0) { if (which) gpu.searchByProjection(&c, &a.last, 15.f, true); else ref.searchByProjection(&c, &a.last, 15.f, true); }
          if (k == 1) { if (which) gpu.searchByProjection(&c, a.local, 3.f); else ref.searchByProjection(&c, a.local, 3.f); }
          if (k == 2) { if (which) gpu.searchByBoW(&a.kf1, &c, out, true); else ref.searchByBoW(&a.kf1, &c, nullptr, out, true); }
          const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
          if (r >= 0) (which ? tg : tr)[k] += ms / reps;
        }
      }
    }
    const char* names[3] = {"searchByProjection(Frame*,Frame*)", "searchByProjection(Frame*,local map)", "searchByBoW(KeyFrame*,Frame*)"};
    for (int k = 0; k < 3; ++k)
      std::printf("timing  %-40s CPU statement %.3f ms   adapter %.3f ms   (1000 features, 900 points)\n", names[k], tr[k], tg[k]);
  }
  std::printf(fails ? "matcher adapter: %d comparisons DIFFER\n" : "matcher adapter: all comparisons identical\n", fails);
  return fails ? 1 : 0;
}
