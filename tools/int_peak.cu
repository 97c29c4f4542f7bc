// tools/int_peak.cu -- measures the integer-pipe peaks the Hamming roofline is quoted against (POPC, LOP3, and the
// LOP3+POPC mix of the carry-save distance), the same way MEASURED_PEAKS.json was produced for HBM: a saturating
// micro-kernel timed with CUDA events.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/int_peak.cu -o /tmp/int_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, unsigned seed) {
  unsigned a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u + blockIdx.x;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i] = __popc(a[i]) + a[i]; }                                   // POPC + IADD (dependent chain x8 ILP)
      else if (MODE == 1) { a[i] = (a[i] ^ a[(i + 1) & 7]) & (a[(i + 2) & 7] | it); }  // LOP3
      else { acc += __popc(a[i] ^ (acc + i)); }                                        // XOR + POPC + ADD
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
double run(const char* name, double ops_per_iter) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, iters = 4096;
  unsigned* out;
  cudaMalloc(&out, blocks * 256 * sizeof(unsigned));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, iters, 1);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, r + 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double ops = (double)blocks * 256 * iters * ops_per_iter;
  const double rate = ops / (best * 1e-3);
  printf("\"%s\": %.4e,\n", name, rate);
  cudaFree(out);
  return rate;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\n\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", p.name, p.multiProcessorCount, p.clockRate);
  double popc = run<0>("popc_per_s", 8);
  run<1>("lop3_per_s", 8);
  run<2>("xor_popc_add_per_s", 8);
  printf("\"popc_per_clk_per_sm\": %.2f\n}\n", popc / ((double)p.multiProcessorCount * p.clockRate * 1e3));
  return 0;
}
