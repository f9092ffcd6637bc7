// selftest.cpp -- tests of the interpreter itself (tests/test_simt_selftest.py): two toy kernels, one with a
// shared-memory race and one with the missing __syncwarp in place, run under whatever SIMT_SCHEDULE says.
#include "cuda_runtime.h"

#include "simt.h"

namespace {
__global__ void neighbour_read(int* out, int with_sync) {
  __shared__ int cell[64];
  const int t = threadIdx.x;
  cell[t] = 0;
  __syncthreads();
  cell[t] = t + 1;
  if (with_sync) __syncthreads();
  out[blockIdx.x * blockDim.x + t] = cell[(t + 1) % blockDim.x];  // reads what the NEXT thread wrote
}
}  // namespace

// Returns how many threads saw their neighbour's value (all of them when the barrier is there).
extern "C" int simt_selftest_neighbours(int with_sync, int blocks) {
  std::vector<int> out(64 * blocks, -1);
  int* p = out.data();
  simt::run_grid(dim3(blocks), dim3(64), [=] { neighbour_read(p, with_sync); });
  int ok = 0;
  for (int b = 0; b < blocks; ++b)
    for (int t = 0; t < 64; ++t) ok += out[b * 64 + t] == (t + 1) % 64 + 1;
  return ok;
}

// The order in which the threads of one block get their first turn, to see the schedule that is in force.
extern "C" void simt_selftest_order(int* first_turn, int n) {
  static int counter;
  counter = 0;
  simt::run_grid(dim3(1), dim3(n), [=] { first_turn[threadIdx.x] = counter++; });
}
