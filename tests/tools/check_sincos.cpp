// tests/tools/check_sincos.cpp -- exhaustive pinning of the fp64 restatement of glibc's float sin/cos
// (the device function glibc_sincosf in vo_slam_test_b200/csrc/orb_kernels.cu) against the box's libm,
// for EVERY float32 in [0, 6.2832] (the only inputs ORBextractor.cpp:114-115 can produce).
//   g++ -O2 -ffp-contract=off -mfma tests/tools/check_sincos.cpp -o /tmp/check_sincos -lpthread && /tmp/check_sincos
// Prints mismatch counts for the plain (no-FMA) and the FMA-contracted evaluation order.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

static const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
                    c4 = 0x1.99343027bf8c3p-16;
static const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
static const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
static const double sign4[4] = {1.0, -1.0, -1.0, 1.0};

template <bool FMA> static inline double mad(double a, double b, double c) { return FMA ? fma(a, b, c) : a * b + c; }

template <bool FMA>
static inline void restated(float y, float* sp, float* cp) {
  double x = y;
  int n = 0;
  double k = 1.0;
  if (fabsf(y) < 0x1.921FB6p-1f) {
    if (fabsf(y) < 0x1p-12f) { *sp = y; *cp = 1.0f; return; }
  } else {
    double r = x * hpi_inv;
    n = ((int32_t)r + 0x800000) >> 24;
    x = mad<FMA>(-(double)n, hpi, x);
    double s = sign4[n & 3];
    if (n & 2) k = -1.0;
    double x2 = x * x;
    x = x * s;
    // fallthrough to poly with x2
    double x4 = x2 * x2, x3 = x2 * x;
    double cc2 = mad<FMA>(x2, k * c4, k * c3);
    double ss1 = mad<FMA>(x2, s3, s2);
    double cc1 = mad<FMA>(x2, k * c1, k * c0);
    double x5 = x3 * x2, x6 = x4 * x2;
    double sv = mad<FMA>(x3, s1, x);
    double cv = mad<FMA>(x4, k * c2, cc1);
    float sinv = (float)mad<FMA>(x5, ss1, sv);
    float cosv = (float)mad<FMA>(x6, cc2, cv);
    if (n & 1) { *sp = cosv; *cp = sinv; } else { *sp = sinv; *cp = cosv; }
    return;
  }
  double x2 = x * x;
  double x4 = x2 * x2, x3 = x2 * x;
  double cc2 = mad<FMA>(x2, c4, c3);
  double ss1 = mad<FMA>(x2, s3, s2);
  double cc1 = mad<FMA>(x2, c1, c0);
  double x5 = x3 * x2, x6 = x4 * x2;
  double sv = mad<FMA>(x3, s1, x);
  double cv = mad<FMA>(x4, c2, cc1);
  *sp = (float)mad<FMA>(x5, ss1, sv);
  *cp = (float)mad<FMA>(x6, cc2, cv);
}

int main() {
  float hi = 6.2832f;
  uint32_t hib; memcpy(&hib, &hi, 4);
  const int NT = std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 8;
  std::atomic<long long> badPlain{0}, badFma{0}, badPair{0};
  std::vector<std::thread> th;
  for (int t = 0; t < NT; ++t)
    th.emplace_back([&, t]() {
      long long bp = 0, bf = 0, bq = 0;
      for (uint64_t b = t; b <= hib; b += NT) {
        uint32_t bb = (uint32_t)b; float y; memcpy(&y, &bb, 4);
        float s, c; sincosf(y, &s, &c);
        float s0 = sinf(y), c0v = cosf(y);
        float sa, ca, sb, cb;
        restated<false>(y, &sa, &ca);
        restated<true>(y, &sb, &cb);
        if (memcmp(&s, &sa, 4) || memcmp(&c, &ca, 4)) ++bp;
        if (memcmp(&s, &sb, 4) || memcmp(&c, &cb, 4)) ++bf;
        if (memcmp(&s, &s0, 4) || memcmp(&c, &c0v, 4)) ++bq;
      }
      badPlain += bp; badFma += bf; badPair += bq;
    });
  for (auto& x : th) x.join();
  printf("floats checked: %u\n", hib + 1);
  printf("mismatch vs libm sincosf: plain=%lld fma=%lld ; sincosf vs (sinf,cosf): %lld\n", (long long)badPlain,
         (long long)badFma, (long long)badPair);
  return 0;
}
