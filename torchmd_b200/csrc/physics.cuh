// physics.cuh -- per-pair / per-term arithmetic of the torchmd force field,
// shared by every kernel.  Pure functions on scalars; compiled for the device
// by nvcc and (for unit tests only, tests/hostcheck) for the host by g++.
//
// Two kinds of arithmetic live here:
//  * the DECISION arithmetic (minimum image, squared distance, cutoff test) is
//    written with explicitly rounded single operations so that a pair is inside
//    the cutoff here exactly when the reference's fp32 torch path says so
//    (forces.py:360-372, 76-81): sub, div, round-half-even, mul, sub, then
//    sqrt_rn(fma(z,z,fma(y,y,x*x))) <= fl32(cutoff)   [measured, see
//    tests/golden/PROVENANCE.txt];
//  * the VALUE arithmetic (energies, force coefficients) is ordinary fp32 with
//    FMA contraction allowed; it only has to meet the 1e-4 kcal/mol/A tolerance.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TMD_HD __host__ __device__ __forceinline__
#else
#define TMD_HD inline
#endif

namespace tmd {

// ---- exactly rounded primitives ------------------------------------------------
TMD_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
TMD_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
TMD_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
TMD_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
TMD_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
TMD_HD float sqrt_rn(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  volatile float r = sqrtf(a);
  return r;
#endif
}
TMD_HD float rsqrt_fast(float a) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return 1.0f / sqrtf(a);
#endif
}

TMD_HD float rcp_rn(float a) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(a);
#else
  return 1.0f / a;
#endif
}
// 1/sqrt(a) to ~1 ulp: hardware approximation (2 ulp) + one Newton step.  The r^-12
// wall amplifies the relative error of 1/r thirteen-fold, so the raw approximation
// alone would cost ~5e-5 kcal/mol/A on a close O-O pair.
TMD_HD float rsqrt_refined(float a) {
  float y = rsqrt_fast(a);
  float t = a * y;
  return fmaf(y, fmaf(-0.5f * t, y, 0.5f), y);
}

// ---- minimum image, one component -----------------------------------------------
// Reference: w = d - L * round(d / L), four separately rounded ops, round half to
// even (forces.py:364).  The quotient is first estimated with a multiply and
// rounded with the 1.5*2^23 trick; whenever that estimate is further than 0.05
// from a rounding boundary (and small enough for the trick) it provably equals
// rint(div_rn(d, L)), because |d*fl(1/L) - div_rn(d,L)| < 2^-22 |q| < 1e-3 for
// |q| < 4096.  Otherwise the true IEEE division is taken.  L*r and the final
// subtraction are single rounded ops like torch's.
TMD_HD float min_image(float d, float L, float invL) {
  const float magic = 12582912.0f;  // 1.5 * 2^23
  float q = d * invL;
  float r = sub_rn(add_rn(q, magic), magic);
  if (!(fabsf(q - r) < 0.45f) || !(fabsf(q) < 4096.0f)) r = rintf(div_rn(d, L));
  return sub_rn(d, mul_rn(L, r));
}

// Largest fp32 s with sqrt_rn(s) <= rc: turns the reference's `dist <= cutoff` into a
// comparison on the squared distance without changing a single decision (host only).
inline float squared_threshold(float rc) {
  float s = rc * rc;
  while (sqrtf(s) > rc) s = nextafterf(s, 0.0f);
  for (;;) {
    float up = nextafterf(s, INFINITY);
    if (sqrtf(up) <= rc) s = up;
    else break;
  }
  return s;
}

// Squared length in the rounding order of torch.norm(dim=1) on (P,3) fp32.
TMD_HD float norm2_ref(float x, float y, float z) {
  return fma_rn(z, z, fma_rn(y, y, mul_rn(x, x)));
}

// ---- pair parameters (uniform per launch) -----------------------------------------
struct PairParams {
  uint32_t terms;     // TMD_TERM(...) mask of pair terms
  int periodic;       // 0: no wrapping (box all zero)
  int has_cutoff;     // 0: every listed pair interacts
  int has_switch;     // LJ switching on (needs cutoff)
  int rfa;            // reaction-field electrostatics (needs cutoff)
  float s_max;        // largest fp32 s with sqrt_rn(s) <= fl32(cutoff); +inf if no cutoff
  float cutoff;       // fl32(cutoff)
  float switch_dist;  // fl32(switch_dist)
  float inv_sw_width; // 1 / (cutoff - switch_dist)
  float krf, crf;     // reaction-field constants (forces.py:466-468)
  float two_krf;
};

enum : uint32_t {
  T_ELEC = 1u << 5,
  T_LJ = 1u << 6,
  T_REP = 1u << 7,
  T_REPCG = 1u << 8,
};

// Energy and dE/dr of one in-cutoff pair at squared distance s.
//   qq   : coulomb_constant * q_i * q_j
//   A, B : LJ table entries of the type pair
// Returns dE/dr summed over the enabled terms; energies are ADDED to e_lj / e_el /
// e_rep / e_repcg.  Follows forces.py:381-491 including the reference's switched
// LJ force  s*dE/dr + E*s'/r  (the extra 1/r is the reference's, forces.py:410-412).
TMD_HD float pair_terms(const PairParams& pp, float s, float qq, float A, float B,
                        float& e_el, float& e_lj, float& e_rep, float& e_repcg,
                        float& rinv_out) {
  float rinv = rsqrt_refined(s);
  float r = s * rinv;
  float rinv2 = rcp_rn(s);  // correctly rounded: the high powers below inherit 3x, not 6x, its error
  float rinv6 = rinv2 * rinv2 * rinv2;
  float dedr = 0.0f;
  rinv_out = rinv;
  if (pp.terms & T_LJ) {
    float a12 = A * rinv6 * rinv6;
    float b6 = B * rinv6;
    float e = a12 - b6;
    float f = (6.0f * b6 - 12.0f * a12) * rinv;
    if (pp.has_switch) {
      // branch-free: t = 0 below the switch distance gives sw = 1, dsw = 0
      float t = fmaxf((r - pp.switch_dist) * pp.inv_sw_width, 0.0f);
      float sw = 1.0f + t * t * t * (-10.0f + t * (15.0f - t * 6.0f));
      float dsw = t * t * (-30.0f + t * (60.0f - t * 30.0f)) * pp.inv_sw_width;
      f = sw * f + e * dsw * rinv;
      e = e * sw;
    }
    e_lj += e;
    dedr += f;
  }
  if (pp.terms & T_ELEC) {
    if (pp.rfa) {
      e_el += qq * (rinv + pp.krf * s - pp.crf);
      dedr += qq * (pp.two_krf * r - rinv2);
    } else {
      float e = qq * rinv;
      e_el += e;
      dedr -= e * rinv;
    }
  }
  if (pp.terms & T_REP) {
    float a12 = A * rinv6 * rinv6;
    e_rep += a12;
    dedr -= 12.0f * a12 * rinv;
  }
  if (pp.terms & T_REPCG) {
    float b6 = B * rinv6;
    e_repcg += b6;
    dedr -= 6.0f * b6 * rinv;
  }
  return dedr;
}

// ---- bonded terms --------------------------------------------------------------------
struct Vec3 {
  float x, y, z;
};
TMD_HD Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
TMD_HD Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
TMD_HD Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
TMD_HD float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
TMD_HD Vec3 cross(Vec3 a, Vec3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
TMD_HD float norm(Vec3 a) { return sqrtf(dot(a, a)); }

// Minimum-image difference of two positions, reference rounding.
TMD_HD Vec3 delta_ref(Vec3 a, Vec3 b, int periodic, Vec3 L, Vec3 invL) {
  Vec3 d = {sub_rn(a.x, b.x), sub_rn(a.y, b.y), sub_rn(a.z, b.z)};
  if (periodic) {
    d.x = min_image(d.x, L.x, invL.x);
    d.y = min_image(d.y, L.y, invL.y);
    d.z = min_image(d.z, L.z, invL.z);
  }
  return d;
}

// Harmonic bond (forces.py:494-503): E = k (r-r0)^2, dE/dr = 2k (r-r0).
TMD_HD void bond_term(float r, float k, float r0, float& e, float& dedr) {
  float x = r - r0;
  e = k * x * x;
  dedr = 2.0f * k * x;
}

// Harmonic angle (forces.py:506-539).  r21 = p0-p1, r23 = p2-p1.
TMD_HD float angle_term(Vec3 r21, Vec3 r23, float k, float theta0, Vec3& f0, Vec3& f1, Vec3& f2) {
  float inv21 = 1.0f / norm(r21);
  float inv23 = 1.0f / norm(r23);
  float c = dot(r23, r21) * inv21 * inv23;
  c = fminf(fmaxf(c, -1.0f), 1.0f);
  float dth = acosf(c) - theta0;
  float sn = sqrtf(1.0f - c * c);
  float coef = (sn != 0.0f) ? (-2.0f * k * dth / sn) : 0.0f;  // zero force at sin==0
  f0 = (coef * inv21) * ((c * inv21) * r21 - inv23 * r23);
  f2 = (coef * inv23) * ((c * inv23) * r23 - inv21 * r21);
  f1 = -1.0f * (f0 + f2);
  return k * dth * dth;
}

// Torsion angle phi = -atan2(sin, cos) (forces.py:544-553) and the geometric
// factors the force projection needs.
struct TorsionGeom {
  Vec3 cA, cB;
  float nA2, nB2, n23, g1, g2, phi;
};
TMD_HD TorsionGeom torsion_geom(Vec3 r12, Vec3 r23, Vec3 r34) {
  TorsionGeom g;
  g.cA = cross(r12, r23);
  g.cB = cross(r23, r34);
  Vec3 cC = cross(r23, g.cA);
  float nA = norm(g.cA), nB = norm(g.cB), nC = norm(cC);
  Vec3 uB = (1.0f / nB) * g.cB;
  float cosphi = dot(g.cA, uB) / nA;
  float sinphi = dot(cC, uB) / nC;
  g.phi = -atan2f(sinphi, cosphi);
  g.nA2 = nA * nA;
  g.nB2 = nB * nB;
  float n23sq = dot(r23, r23);
  g.n23 = sqrtf(n23sq);
  g.g1 = dot(r12, r23) / n23sq;
  g.g2 = dot(r34, r23) / n23sq;
  return g;
}
// One torsion term: energy and dE/dphi-like coefficient (forces.py:566-579).
TMD_HD void torsion_term(float phi, float k, float phi0, float per, int amber_form, float& e,
                         float& coef) {
  if (amber_form) {
    float a = per * phi - phi0;
    e += k * (1.0f + cosf(a));
    coef += -per * k * sinf(a);
  } else {
    const float pi = 3.14159265358979323846f;
    float a = phi - phi0;
    if (a < -pi) a += 2.0f * pi;
    else if (a > pi) a -= 2.0f * pi;
    e += k * a * a;
    coef += 2.0f * k * a;
  }
}
// Force projection (forces.py:584-603).
TMD_HD void torsion_forces(const TorsionGeom& g, float coef, Vec3& f0, Vec3& f1, Vec3& f2, Vec3& f3) {
  float ff0 = (-coef * g.n23) / g.nA2;
  float ff3 = (coef * g.n23) / g.nB2;
  Vec3 v0 = ff0 * g.cA;
  Vec3 v3 = ff3 * g.cB;
  Vec3 s = g.g1 * v0 - g.g2 * v3;
  f0 = -1.0f * v0;
  f1 = v0 + s;
  f2 = v3 - s;
  f3 = -1.0f * v3;
}

}  // namespace tmd
