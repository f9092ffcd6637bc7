// Issue / throughput microbenchmark for the packed fp32x2 instructions of sm_100 (FFMA2) against
// scalar FFMA, alone and mixed with integer ALU work -- the question behind k_pair_fx2: does a
// packed operation cost one issue slot for two results?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_f32x2 scripts/ubench_f32x2.cu && /tmp/ubench_f32x2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, UNROLL = 16;

template <int MODE>  // 0: FFMA  1: FFMA2  2: FFMA + IADD3/LOP3 mix  3: FFMA2 + IADD3/LOP3 mix
__global__ void bench(float* out, int n) {
  float2 a[4] = {{1.f, 2.f}, {3.f, 4.f}, {5.f, 6.f}, {7.f, 8.f}};
  const float2 m = {1.0000001f, 0.9999999f}, c = {1e-9f, -1e-9f};
  unsigned k0 = threadIdx.x, k1 = blockIdx.x, k2 = 12345u, k3 = 777u;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (MODE == 0 || MODE == 2) {
          a[q].x = fmaf(a[q].x, m.x, c.x);
          a[q].y = fmaf(a[q].y, m.y, c.y);
        } else {
          a[q] = __ffma2_rn(a[q], m, c);
        }
      }
      if (MODE >= 2) {  // four independent integer ops per eight (MODE 2) / four (MODE 3) FP issue slots
        k0 = (k0 + k3) ^ 0x9e3779b9u;
        k1 = (k1 + k0) ^ 0x85ebca6bu;
        k2 = (k2 + k1) ^ 0xc2b2ae35u;
        k3 = (k3 + k2) ^ 0x27d4eb2fu;
      }
    }
  }
  float s = 0.f;
  for (int q = 0; q < 4; ++q) s += a[q].x + a[q].y;
  if (s == 12345.678f || (k0 ^ k1 ^ k2 ^ k3) == 0xdeadbeefu) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256;
  bench<MODE><<<blocks, threads>>>(d, 0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<blocks, threads>>>(d, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * threads * ITERS * UNROLL * 8;  // fp32 FMAs
  printf("%-28s %8.3f ms  %7.2f TFMA/s  (%6.2f fp32 TFLOP/s)\n", name, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d;
  cudaMalloc(&d, 1 << 24);
  run<0>("FFMA", d, sms);
  run<1>("FFMA2", d, sms);
  run<2>("FFMA + 4 int ops / 8 FMA", d, sms);
  run<3>("FFMA2 + 4 int ops / 8 FMA", d, sms);
  printf("equal FMA/s for FFMA and FFMA2 alone but FFMA2 faster in the mix => packed ops save issue slots, not datapath\n");
  return 0;
}
