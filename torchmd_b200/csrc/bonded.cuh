// bonded.cuh -- bonded terms (K4): bonds, angles, proper/improper torsions, 1-4.
//
// Replaces forces.py:122-258 (+ evaluate_bonds/angles/torsion, forces.py:494-605).
// One thread per term instance, original atom order, forces accumulated with
// fp32 atomics AFTER the pair kernel has written the non-bonded force (so the
// pair kernel's plain store doubles as the zeroing of Forces.compute,
// forces.py:113-114).  These are O(N) and a few percent of the pair work.
#pragma once
#include "context.cuh"
#include "pair.cuh"

namespace tmd {

constexpr int BONDED_THREADS = 128;

struct BoxView {
  int periodic;
  Vec3 L, invL;
};
__device__ __forceinline__ BoxView box_of(const DeviceState& S, int r) {
  const Grid* g = S.grid + r;
  BoxView b;
  b.periodic = g->periodic;
  b.L = {g->L[0], g->L[1], g->L[2]};
  b.invL = {g->invL[0], g->invL[1], g->invL[2]};
  return b;
}
__device__ __forceinline__ Vec3 load3(const float* p, size_t atom) {
  return {p[atom * 3 + 0], p[atom * 3 + 1], p[atom * 3 + 2]};
}
__device__ __forceinline__ void add3(float* f, size_t atom, Vec3 v) {
  atomicAdd(f + atom * 3 + 0, v.x);
  atomicAdd(f + atom * 3 + 1, v.y);
  atomicAdd(f + atom * 3 + 2, v.z);
}

// bonds: E = k (r-r0)^2; bonds longer than the cutoff are skipped like the
// reference does (forces.py:128-136).
__global__ void __launch_bounds__(BONDED_THREADS)
k_bonds(DeviceState S, BondedSet B, const float* __restrict__ pos, float* __restrict__ forces,
        double* __restrict__ energies) {
  const int r = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)r * S.natoms;
  float e = 0.f;
  if (t < B.n) {
    const BoxView bx = box_of(S, r);
    const int i = B.idx[2 * t], j = B.idx[2 * t + 1];
    const Vec3 d = delta_ref(load3(pos, base + i), load3(pos, base + j), bx.periodic, bx.L, bx.invL);
    const float dist = sqrt_rn(norm2_ref(d.x, d.y, d.z));
    if (!S.pp.has_cutoff || dist <= S.pp.cutoff) {
      float dedr;
      bond_term(dist, B.prm[2 * t], B.prm[2 * t + 1], e, dedr);
      const Vec3 fv = (dedr / dist) * d;
      add3(forces, base + i, -1.0f * fv);
      add3(forces, base + j, fv);
    }
  }
  if (energies) {
    __shared__ double red[BONDED_THREADS / 32];
    block_accumulate<BONDED_THREADS / 32>((double)e, energies + (size_t)r * TMD_NUM_ENERGIES + TMD_E_BONDS, red);
  }
}

__global__ void __launch_bounds__(BONDED_THREADS)
k_angles(DeviceState S, BondedSet B, const float* __restrict__ pos, float* __restrict__ forces,
         double* __restrict__ energies) {
  const int r = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)r * S.natoms;
  float e = 0.f;
  if (t < B.n) {
    const BoxView bx = box_of(S, r);
    const int a0 = B.idx[3 * t], a1 = B.idx[3 * t + 1], a2 = B.idx[3 * t + 2];
    const Vec3 p1 = load3(pos, base + a1);
    const Vec3 r21 = delta_ref(load3(pos, base + a0), p1, bx.periodic, bx.L, bx.invL);
    const Vec3 r23 = delta_ref(load3(pos, base + a2), p1, bx.periodic, bx.L, bx.invL);
    Vec3 f0, f1, f2;
    e = angle_term(r21, r23, B.prm[2 * t], B.prm[2 * t + 1], f0, f1, f2);
    add3(forces, base + a0, f0);
    add3(forces, base + a1, f1);
    add3(forces, base + a2, f2);
  }
  if (energies) {
    __shared__ double red[BONDED_THREADS / 32];
    block_accumulate<BONDED_THREADS / 32>((double)e, energies + (size_t)r * TMD_NUM_ENERGIES + TMD_E_ANGLES, red);
  }
}

// proper dihedrals (slot TMD_E_DIHEDRALS) and impropers (slot TMD_E_IMPROPERS)
__global__ void __launch_bounds__(BONDED_THREADS)
k_torsions(DeviceState S, BondedSet B, int slot, const float* __restrict__ pos,
           float* __restrict__ forces, double* __restrict__ energies) {
  const int r = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)r * S.natoms;
  float e = 0.f;
  if (t < B.n) {
    const BoxView bx = box_of(S, r);
    const int a0 = B.idx[4 * t], a1 = B.idx[4 * t + 1], a2 = B.idx[4 * t + 2], a3 = B.idx[4 * t + 3];
    const Vec3 p0 = load3(pos, base + a0), p1 = load3(pos, base + a1);
    const Vec3 p2 = load3(pos, base + a2), p3 = load3(pos, base + a3);
    const Vec3 r12 = delta_ref(p0, p1, bx.periodic, bx.L, bx.invL);
    const Vec3 r23 = delta_ref(p1, p2, bx.periodic, bx.L, bx.invL);
    const Vec3 r34 = delta_ref(p2, p3, bx.periodic, bx.L, bx.invL);
    const TorsionGeom g = torsion_geom(r12, r23, r34);
    float coef = 0.f;
    for (int m = B.term_ptr[t]; m < B.term_ptr[t + 1]; ++m)
      torsion_term(g.phi, B.terms[3 * m], B.terms[3 * m + 1], B.terms[3 * m + 2], B.amber, e, coef);
    Vec3 f0, f1, f2, f3;
    torsion_forces(g, coef, f0, f1, f2, f3);
    add3(forces, base + a0, f0);
    add3(forces, base + a1, f1);
    add3(forces, base + a2, f2);
    add3(forces, base + a3, f3);
  }
  if (energies) {
    __shared__ double red[BONDED_THREADS / 32];
    block_accumulate<BONDED_THREADS / 32>((double)e, energies + (size_t)r * TMD_NUM_ENERGIES + slot, red);
  }
}

// 1-4 pairs (forces.py:185-236): LJ scaled by 1/scnb with no cutoff and no switch,
// Coulomb scaled by 1/scee and never reaction-field; energies are booked under the
// lj / electrostatics slots, the "1-4" slot stays zero.
__global__ void __launch_bounds__(BONDED_THREADS)
k_pairs14(DeviceState S, BondedSet B, const float* __restrict__ q_scaled,
          const float* __restrict__ pos, float* __restrict__ forces, double* __restrict__ energies) {
  const int r = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)r * S.natoms;
  float e_lj = 0.f, e_el = 0.f;
  if (t < B.n) {
    const BoxView bx = box_of(S, r);
    const int i = B.idx[2 * t], j = B.idx[2 * t + 1];
    const Vec3 d = delta_ref(load3(pos, base + i), load3(pos, base + j), bx.periodic, bx.L, bx.invL);
    const float dist = sqrt_rn(norm2_ref(d.x, d.y, d.z));
    const float rinv = 1.0f / dist;
    const float A = B.prm[4 * t], Bc = B.prm[4 * t + 1], scnb = B.prm[4 * t + 2], scee = B.prm[4 * t + 3];
    float dedr = 0.f;
    if (S.pp.terms & T_LJ) {
      const float r6 = rinv * rinv * rinv * rinv * rinv * rinv;
      const float a12 = A * r6 * r6, b6 = Bc * r6;
      e_lj = (a12 - b6) / scnb;
      dedr += (6.0f * b6 - 12.0f * a12) * rinv / scnb;
    }
    if (S.pp.terms & T_ELEC) {
      e_el = q_scaled[i] * q_scaled[j] * rinv / scee;
      dedr -= e_el * rinv;
    }
    const Vec3 fv = (dedr * rinv) * d;
    add3(forces, base + i, -1.0f * fv);
    add3(forces, base + j, fv);
  }
  if (energies) {
    __shared__ double red[BONDED_THREADS / 32];
    double* E = energies + (size_t)r * TMD_NUM_ENERGIES;
    if (S.pp.terms & T_LJ) block_accumulate<BONDED_THREADS / 32>((double)e_lj, E + TMD_E_LJ, red);
    if (S.pp.terms & T_ELEC) block_accumulate<BONDED_THREADS / 32>((double)e_el, E + TMD_E_ELECTROSTATICS, red);
  }
}

}  // namespace tmd
