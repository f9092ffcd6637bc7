// Throughput of fire-and-forget global reductions (RED.ADD.F32) on sm_100 -- the number behind "why a full list and
// not Newton's third law" (DESIGN.md 4): a half list needs one float3 reduction per in-cutoff pair (1.5e7 per step at
// 99,999 atoms).  Patterns: the partner indices of a real row are ascending but sparse, so (a) scattered addresses,
// (b) neighbouring lanes on neighbouring atoms (AoS float4 records, 3 scalar reds), (c) one vector red.v4.f32 per lane.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_red scripts/ubench_red.cu && /tmp/ubench_red
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 100000;        // atoms (float4 force records: 1.6 MB, L2 resident like the real array)
constexpr int PER_THREAD = 256;  // reductions per thread

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int MODE>  // 0 scattered scalar x3, 1 row-like (lane l -> atom base + 2*l) scalar x3, 2 row-like vector v4
__global__ void bench(float4* f) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = t >> 5;
  for (int k = 0; k < PER_THREAD; ++k) {
    unsigned j;
    if (MODE == 0) j = mix(t * 977u + k) % N;
    else j = (mix(warp * 131u + k) % (N - 64)) + 2 * lane;  // a warp's 32 partners within 64 consecutive atoms
    float* p = reinterpret_cast<float*>(f + j);
    const float v = 1e-6f * (float)(k + 1);
    if (MODE == 2) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v), "f"(-v), "f"(0.5f * v), "f"(0.f) : "memory");
    } else {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 1), "f"(-v) : "memory");
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 2), "f"(0.5f * v) : "memory");
    }
  }
}

template <int MODE>
void run(const char* name, float4* d, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256;
  bench<MODE><<<blocks, threads>>>(d);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<blocks, threads>>>(d);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double pairs = (double)blocks * threads * PER_THREAD;  // float3 reductions
  printf("%-44s %8.3f ms  %8.2f G float3-reductions/s  => 1.5e7 pairs in %7.1f us\n", name, ms, pairs / ms * 1e-6, 1.5e7 / (pairs / ms) * 1e3);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float4* d = nullptr;
  cudaMalloc(&d, sizeof(float4) * N);
  cudaMemset(d, 0, sizeof(float4) * N);
  run<0>("scattered atoms, 3 scalar red.add.f32", d, sms);
  run<1>("row-like partners, 3 scalar red.add.f32", d, sms);
  run<2>("row-like partners, 1 red.add.v4.f32", d, sms);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
