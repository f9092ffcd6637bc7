// integrate.cuh -- velocity Verlet + Langevin kick (K5, K6) and kinetic energy.
//
// Replaces integrator.py:61-74 (_first_VV, langevin, _second_VV) and
// kinetic_energy (integrator.py:8-30).  Pure streaming kernels, one thread per
// atom; the arithmetic keeps the reference's operation order with single rounded
// ops, so with injected noise the trajectory matches the reference fp32 path to
// the last bit as long as the forces do.
#pragma once
#include "context.cuh"
#include "pair.cuh"

namespace tmd {

constexpr int INTEG_THREADS = 256;

// pos += vel*dt + ((0.5*a)*dt)*dt ;  vel += (0.5*dt)*a ;  a = F/m     (integrator.py:61-64)
__global__ void __launch_bounds__(INTEG_THREADS)
k_vv_first(int natoms, int lo, int cnt, unsigned long long* counters, float* __restrict__ pos,
           float* __restrict__ vel, const float* __restrict__ forces, const float* __restrict__ masses,
           float dt, float hdt) {
  const int r = blockIdx.y;
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;  // atoms [lo, lo+cnt) only
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) counters[1] += 1;  // Philox position of this step
  if (i >= lo + cnt) return;
  const float m = masses[i];
  const size_t a = ((size_t)r * natoms + i) * 3;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float acc = div_rn(forces[a + d], m);
    const float v = vel[a + d];
    const float drift = add_rn(mul_rn(v, dt), mul_rn(mul_rn(mul_rn(0.5f, acc), dt), dt));
    pos[a + d] = add_rn(pos[a + d], drift);
    vel[a + d] = add_rn(v, mul_rn(hdt, acc));
  }
}

// Philox4x32-10 counter-based generator (Salmon et al., SC'11).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// three N(0,1) draws for (atom slot, step): Box-Muller on Philox output
__device__ __forceinline__ void normal3(uint64_t seed, uint64_t step, uint64_t slot, float out[3]) {
  const uint4 u = philox4x32_10(make_uint4((unsigned)slot, (unsigned)(slot >> 32), (unsigned)step, (unsigned)(step >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float two_pi = 6.283185307179586f;
  const float u0 = ((float)(u.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = ((float)(u.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(u.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = ((float)(u.w >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s, c;
  sincosf(two_pi * u1, &s, &c);
  out[0] = r0 * c;
  out[1] = r0 * s;
  out[2] = r1 * cosf(two_pi * u3);
}

// [vel += ((-gamma*vel)*dt + xi*vcoeff)]  then  vel += (0.5*dt)*(F/m)
// (integrator.py:72-74 then 67-69), optionally followed by the kinetic energy.
template <bool THERMOSTAT, bool KINETIC>
__global__ void __launch_bounds__(INTEG_THREADS)
k_vv_second(int natoms, int lo, int cnt, const unsigned long long* __restrict__ counters,
            float* __restrict__ vel, const float* __restrict__ forces,
            const float* __restrict__ masses, float dt, float hdt, float neg_gamma,
            const float* __restrict__ vcoeff, const float* __restrict__ noise, uint64_t seed,
            uint64_t step_offset, double* __restrict__ ke) {
  const int r = blockIdx.y;
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t step = step_offset + counters[1];  // device-resident step count: graph replayable
  double ek = 0.0;
  if (i < lo + cnt) {
    const float m = masses[i];
    const size_t slot = (size_t)r * natoms + i;
    const size_t a = slot * 3;
    float xi[3] = {0.f, 0.f, 0.f};
    float vc = 0.f;
    if (THERMOSTAT) {
      vc = vcoeff[i];
      if (noise) {
        xi[0] = noise[a]; xi[1] = noise[a + 1]; xi[2] = noise[a + 2];
      } else {
        normal3(seed, step, slot, xi);
      }
    }
    float v2 = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = vel[a + d];
      if (THERMOSTAT) v = add_rn(v, add_rn(mul_rn(mul_rn(neg_gamma, v), dt), mul_rn(xi[d], vc)));
      v = add_rn(v, mul_rn(hdt, div_rn(forces[a + d], m)));
      vel[a + d] = v;
      v2 += v * v;
    }
    if (KINETIC) ek = 0.5 * (double)m * (double)v2;
  }
  if (KINETIC) {
    __shared__ double red[INTEG_THREADS / 32];
    block_accumulate<INTEG_THREADS / 32>(ek, ke + r, red);
  }
}

__global__ void __launch_bounds__(INTEG_THREADS)
k_kinetic(int natoms, int lo, int cnt, const float* __restrict__ vel, const float* __restrict__ masses,
          double* __restrict__ ke) {
  const int r = blockIdx.y;
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  double ek = 0.0;
  if (i < lo + cnt) {
    const size_t a = ((size_t)r * natoms + i) * 3;
    const float vx = vel[a], vy = vel[a + 1], vz = vel[a + 2];
    ek = 0.5 * (double)masses[i] * (double)(vx * vx + vy * vy + vz * vz);
  }
  __shared__ double red[INTEG_THREADS / 32];
  block_accumulate<INTEG_THREADS / 32>(ek, ke + r, red);
}

}  // namespace tmd
