// pair.cuh -- the non-bonded pair kernel (K3): LJ (+switch), Coulomb / reaction
// field, repulsion, repulsionCG over the full Verlet list.
//
// Replaces, per step and replica, the reference's all-pairs distance pass, cutoff
// mask, per-term evaluation and index_add_ scatter (forces.py:264-319, 381-491).
//
// Mapping: one warp per atom (sorted order), lanes stride over the atom's
// neighbour row.  Row reads are coalesced 128-byte lines; partner records are
// 16-byte gathers that hit L1/L2 because atoms are sorted by cell.  Every pair is
// seen from both sides (full list), so forces need no atomics and no scatter:
// each warp reduces its atom's force with shuffles and writes it once.  The
// cutoff decision uses the reference's exact fp32 predicate (physics.cuh).
#pragma once
#include "context.cuh"

namespace tmd {

constexpr int PAIR_WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-level sum of one double per thread -> atomicAdd into *dst by thread 0.
template <int NWARPS>
__device__ __forceinline__ void block_accumulate(double v, double* dst, double* smem /*NWARPS*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) t += smem[w];
    if (t != 0.0) atomicAdd(dst, t);
  }
}

template <bool ENERGY, bool PERIODIC>
__global__ void __launch_bounds__(PAIR_WARPS * 32)
k_pair(DeviceState S, float* __restrict__ forces, double* __restrict__ energies) {
  const int r = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
  const int N = S.natoms;
  const size_t base = (size_t)r * N;
  const PairParams pp = S.pp;
  float e_el = 0.f, e_lj = 0.f, e_rep = 0.f, e_cg = 0.f;

  if (k < N) {
    const float4* __restrict__ xq = S.xq_s + base;
    const int* __restrict__ types = S.type_s + base;
    const int* __restrict__ row = S.nbr + (base + k) * (size_t)S.row_cap;
    const int n = S.nnbr[base + k];
    const float4 pi = xq[k];
    const int ti = types[k] * S.ntypes;
    const bool need_ab = (pp.terms & (T_LJ | T_REP | T_REPCG)) != 0;
    float Lx = 0.f, Ly = 0.f, Lz = 0.f, iLx = 0.f, iLy = 0.f, iLz = 0.f;
    if (PERIODIC) {
      const Grid* g = S.grid + r;
      Lx = g->L[0]; Ly = g->L[1]; Lz = g->L[2];
      iLx = g->invL[0]; iLy = g->invL[1]; iLz = g->invL[2];
    }
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll 2
    for (int e = lane; e < n; e += 32) {
      const int j = __ldcs(row + e);  // streamed once per step: do not keep in L1
      const float4 pj = xq[j];
      float dx = sub_rn(pi.x, pj.x), dy = sub_rn(pi.y, pj.y), dz = sub_rn(pi.z, pj.z);
      if (PERIODIC) {
        dx = min_image(dx, Lx, iLx);
        dy = min_image(dy, Ly, iLy);
        dz = min_image(dz, Lz, iLz);
      }
      const float s = norm2_ref(dx, dy, dz);
      if (s <= pp.s_max) {
        float2 ab = make_float2(0.f, 0.f);
        if (need_ab) ab = __ldg(S.AB + ti + types[j]);
        float rinv;
        const float dedr = pair_terms(pp, s, pi.w * pj.w, ab.x, ab.y, e_el, e_lj, e_rep, e_cg, rinv);
        const float c = dedr * rinv;  // force on i is -unit*dE/dr = -(w/r) dE/dr
        fx -= dx * c;
        fy -= dy * c;
        fz -= dz * c;
      }
    }
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (lane == 0) {
      float* f = forces + (base + S.perm[base + k]) * 3;
      f[0] = fx;
      f[1] = fy;
      f[2] = fz;
    }
  }
  if (ENERGY) {
    __shared__ double red[PAIR_WARPS];
    double* E = energies + (size_t)r * TMD_NUM_ENERGIES;
    // every pair is visited from both of its atoms
    if (pp.terms & T_ELEC) block_accumulate<PAIR_WARPS>(0.5 * (double)e_el, E + TMD_E_ELECTROSTATICS, red);
    if (pp.terms & T_LJ) block_accumulate<PAIR_WARPS>(0.5 * (double)e_lj, E + TMD_E_LJ, red);
    if (pp.terms & T_REP) block_accumulate<PAIR_WARPS>(0.5 * (double)e_rep, E + TMD_E_REPULSION, red);
    if (pp.terms & T_REPCG) block_accumulate<PAIR_WARPS>(0.5 * (double)e_cg, E + TMD_E_REPULSIONCG, red);
  }
}

}  // namespace tmd
