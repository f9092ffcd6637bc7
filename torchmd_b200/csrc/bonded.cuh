// bonded.cuh -- bonded terms (K4): bonds, angles, proper/improper torsions, 1-4.
//
// Replaces forces.py:122-258 (+ evaluate_bonds/angles/torsion, forces.py:494-605).
//
// Atom-centric and deterministic: one thread per ATOM walks the list of bonded term
// instances the atom takes part in (CSR built once on the host from the topology),
// re-evaluates each term, keeps the force on its own atom, and adds the sum to the
// force array with a plain read-modify-write -- no atomics, fixed summation order,
// bitwise reproducible.  Every term is evaluated once per participating atom (2-4x
// redundant arithmetic on an O(N) workload that is a few percent of the pair kernel);
// its energy is booked by the atom in slot 0 only.  Runs AFTER the pair kernel, whose
// plain store of the non-bonded force doubles as the zeroing of Forces.compute
// (forces.py:113-114).
#pragma once
#include "context.cuh"
#include "pair.cuh"

namespace tmd {

constexpr int BONDED_THREADS = 128;

// kinds of term instance an atom entry can point at
enum { BK_BOND = 0, BK_ANGLE = 1, BK_DIHEDRAL = 2, BK_IMPROPER = 3, BK_PAIR14 = 4 };

struct BondedTables {
  const int* atom_ptr;   // (natoms+1) CSR over atoms
  const int* entries;    // packed: kind (3 bits) | slot in the term (2 bits) | term index (27 bits)
  BondedSet bonds, angles, torsions[2], pairs14;
};

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

struct BondedEnergies {
  double bond = 0., angle = 0., dih = 0., imp = 0., lj = 0., el = 0.;
};

// Sum of the bonded forces on atom a of replica r (fp64), energies booked by the atom in slot 0 of each term.
__device__ __forceinline__ Vec3d bonded_force_on_atom(const DeviceState& S, const BondedTables& T, const float* __restrict__ q_scaled,
                                                      const float* __restrict__ pos, int r, int a, BondedEnergies& E) {
  const size_t base = (size_t)r * S.natoms;
  const BoxView bx = box_of(S, r);
  Vec3d f = {0., 0., 0.};
  double &e_bond = E.bond, &e_angle = E.angle, &e_dih = E.dih, &e_imp = E.imp, &e_lj = E.lj, &e_el = E.el;
  for (int p = T.atom_ptr[a]; p < T.atom_ptr[a + 1]; ++p) {
    const unsigned ent = (unsigned)T.entries[p];
    const int kind = ent >> 29, slot = (ent >> 27) & 3, t = ent & 0x7ffffff;
    if (kind == BK_BOND) {
      // E = k (r-r0)^2; bonds longer than the cutoff are skipped like the reference (forces.py:128-136)
      const int i = T.bonds.idx[2 * t], j = T.bonds.idx[2 * t + 1];
      const Vec3 pi = load3(pos, base + i), pj = load3(pos, base + j);
      const Vec3 dref = delta_ref(pi, pj, bx.periodic, bx.L, bx.invL);
      if (!S.pp.has_cutoff || sqrt_rn(norm2_ref(dref.x, dref.y, dref.z)) <= S.pp.cutoff) {  // reference decision
        const Vec3d d = delta_f64(pi, pj, bx.periodic, bx.L);
        const double dist = norm(d);
        double e, dedr;
        bond_term<double>(dist, T.bonds.prm[2 * t], T.bonds.prm[2 * t + 1], e, dedr);
        const Vec3d fv = (dedr / dist) * d;  // force on j; i gets the opposite
        f = slot == 0 ? f - fv : f + fv;
        if (slot == 0) e_bond += e;
      }
    } else if (kind == BK_ANGLE) {
      const int a0 = T.angles.idx[3 * t], a1 = T.angles.idx[3 * t + 1], a2 = T.angles.idx[3 * t + 2];
      const Vec3 p1 = load3(pos, base + a1);
      const Vec3d r21 = delta_f64(load3(pos, base + a0), p1, bx.periodic, bx.L);
      const Vec3d r23 = delta_f64(load3(pos, base + a2), p1, bx.periodic, bx.L);
      Vec3d f0, f1, f2;
      const double e = angle_term<double>(r21, r23, T.angles.prm[2 * t], T.angles.prm[2 * t + 1], f0, f1, f2);
      f = f + (slot == 0 ? f0 : (slot == 1 ? f1 : f2));
      if (slot == 0) e_angle += e;
    } else if (kind == BK_DIHEDRAL || kind == BK_IMPROPER) {
      const BondedSet& B = T.torsions[kind == BK_IMPROPER];
      const Vec3 p0 = load3(pos, base + B.idx[4 * t]), p1 = load3(pos, base + B.idx[4 * t + 1]);
      const Vec3 p2 = load3(pos, base + B.idx[4 * t + 2]), p3 = load3(pos, base + B.idx[4 * t + 3]);
      const Vec3d r12 = delta_f64(p0, p1, bx.periodic, bx.L);
      const Vec3d r23 = delta_f64(p1, p2, bx.periodic, bx.L);
      const Vec3d r34 = delta_f64(p2, p3, bx.periodic, bx.L);
      const TorsionGeom<double> g = torsion_geom(r12, r23, r34);
      double e = 0., coef = 0.;
      for (int m = B.term_ptr[t]; m < B.term_ptr[t + 1]; ++m)
        torsion_term<double>(g.phi, B.terms[3 * m], B.terms[3 * m + 1], B.terms[3 * m + 2], B.amber, e, coef);
      Vec3d f0, f1, f2, f3;
      torsion_forces(g, coef, f0, f1, f2, f3);
      f = f + (slot == 0 ? f0 : (slot == 1 ? f1 : (slot == 2 ? f2 : f3)));
      if (slot == 0) {
        if (kind == BK_IMPROPER) e_imp += e;
        else e_dih += e;
      }
    } else {  // BK_PAIR14 (forces.py:185-236): LJ/scnb with no cutoff or switch, Coulomb/scee, never RF
      const int i = T.pairs14.idx[2 * t], j = T.pairs14.idx[2 * t + 1];
      const Vec3d d = delta_f64(load3(pos, base + i), load3(pos, base + j), bx.periodic, bx.L);
      const double dist = norm(d);
      const double rinv = 1.0 / dist;
      const float* prm = T.pairs14.prm + 4 * t;  // A, B, scnb, scee
      double dedr = 0.;
      if (S.pp.terms & T_LJ) {
        const double r6 = rinv * rinv * rinv * rinv * rinv * rinv;
        const double a12 = prm[0] * r6 * r6, b6 = prm[1] * r6;
        if (slot == 0) e_lj += (a12 - b6) / prm[2];
        dedr += (6.0 * b6 - 12.0 * a12) * rinv / prm[2];
      }
      if (S.pp.terms & T_ELEC) {
        const double e = (double)q_scaled[i] * (double)q_scaled[j] * rinv / prm[3];
        if (slot == 0) e_el += e;
        dedr -= e * rinv;
      }
      const Vec3d fv = (dedr * rinv) * d;
      f = slot == 0 ? f - fv : f + fv;
    }
  }
  return f;
}

// block-level reduction of the bonded energies into the per-replica slots
__device__ __forceinline__ void bonded_energy_reduce(const DeviceState& S, const BondedTables& T, int r, const BondedEnergies& E,
                                                     double* __restrict__ energies, double* red) {
  double* Eo = energies + (size_t)r * TMD_NUM_ENERGIES;
  if (T.bonds.n) block_accumulate<BONDED_THREADS / 32>(E.bond, Eo + TMD_E_BONDS, red);
  if (T.angles.n) block_accumulate<BONDED_THREADS / 32>(E.angle, Eo + TMD_E_ANGLES, red);
  if (T.torsions[0].n) block_accumulate<BONDED_THREADS / 32>(E.dih, Eo + TMD_E_DIHEDRALS, red);
  if (T.torsions[1].n) block_accumulate<BONDED_THREADS / 32>(E.imp, Eo + TMD_E_IMPROPERS, red);
  if (T.pairs14.n) {
    if (S.pp.terms & T_LJ) block_accumulate<BONDED_THREADS / 32>(E.lj, Eo + TMD_E_LJ, red);
    if (S.pp.terms & T_ELEC) block_accumulate<BONDED_THREADS / 32>(E.el, Eo + TMD_E_ELECTROSTATICS, red);
  }
}

__global__ void __launch_bounds__(BONDED_THREADS)
k_bonded(DeviceState S, BondedTables T, const float* __restrict__ q_scaled, const float* __restrict__ pos,
         float* __restrict__ forces, double* __restrict__ energies, double* __restrict__ scratch) {
  const int r = blockIdx.y;
  const int a = S.own_lo + blockIdx.x * blockDim.x + threadIdx.x;  // owned atoms only
  const size_t base = (size_t)r * S.natoms;
  BondedEnergies E;
  if (a < S.own_lo + S.own_n) {
    const Vec3d f = bonded_force_on_atom(S, T, q_scaled, pos, r, a, E);
    if (scratch) {
      // overlapped with the pair kernel on another stream: the fp64 sums go to a scratch
      // buffer, k_add_bonded folds them into the forces afterwards (same single rounding)
      double* o = scratch + (base + a) * 3;
      o[0] = f.x;
      o[1] = f.y;
      o[2] = f.z;
    } else if (S.cl.on) {
      // cluster path: the pair forces sit in slot order; this kernel brings them home and adds its own sums
      const float4 pf = S.cl.f[(size_t)r * (S.cl.slots + 1) + S.cl.inv[base + a]];
      float* out = forces + (base + a) * 3;
      out[0] = (float)((double)pf.x + f.x);
      out[1] = (float)((double)pf.y + f.y);
      out[2] = (float)((double)pf.z + f.z);
    } else if (T.atom_ptr[a + 1] > T.atom_ptr[a]) {
      float* out = forces + (base + a) * 3;
      out[0] = (float)((double)out[0] + f.x);
      out[1] = (float)((double)out[1] + f.y);
      out[2] = (float)((double)out[2] + f.z);
    }
  }
  if (energies) {
    __shared__ double red[BONDED_THREADS / 32];
    bonded_energy_reduce(S, T, r, E, energies, red);
  }
}

// forces += bonded sums of k_bonded's scratch mode (owned atoms), rounded once like the in-place path
__global__ void __launch_bounds__(BONDED_THREADS)
k_add_bonded(int natoms, int lo, int cnt, float* __restrict__ forces, const double* __restrict__ scratch) {
  const int a = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= lo + cnt) return;
  const size_t e = ((size_t)blockIdx.y * natoms + a) * 3;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double s = scratch[e + d];
    if (s != 0.0) forces[e + d] = (float)((double)forces[e + d] + s);
  }
}

}  // namespace tmd
