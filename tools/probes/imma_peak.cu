// Peak rate of the legacy warp-level integer MMA (mma.sync.m16n8k32 u8 -> IMMA.16832) on this GPU.  One MMA = 16*8*32 MACs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_peak imma_peak.cu && ./imma_peak
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void __launch_bounds__(256) k(int iters, int* out) {
  uint32_t A[4] = {threadIdx.x, 1, 2, 3}, B[2] = {blockIdx.x, 5};
  int C[8][4];
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) C[j][i] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(C[j][0]), "+r"(C[j][1]), "+r"(C[j][2]), "+r"(C[j][3]) : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]));
  }
  int s = 0;
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += C[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int* out; cudaMalloc(&out, sizeof(int) * sms * 8 * 256);
  for (int warps = 1; warps <= 8; warps *= 2) {
    const int iters = 20000;
    k<<<sms * 8 / (8 / warps > 0 ? 1 : 1), 32 * warps>>>(100, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms * 4, 32 * warps>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * 4 * warps * iters * 8;
    printf("warps/CTA %d (CTAs %d): %.3f ms, %.1f G MMA/s, %.1f TMAC/s = %.0f TOPS, %.2f MMA/clk/SM @1.9GHz\n", warps, sms * 4, ms, mmas / ms / 1e6,
           mmas * 4096 / ms / 1e9, 2 * mmas * 4096 / ms / 1e9, mmas / ms / 1e6 / sms / 1.9);
  }
  return 0;
}
