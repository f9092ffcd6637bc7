// hostcheck.cpp -- UNIT-TEST SHIM, not product code.
// Compiles the scalar device functions of torchmd_b200/csrc/physics.cuh for the
// host so tests/test_physics_host.py can compare them with the oracle on a machine
// without a GPU.  Nothing in torchmd_b200 loads this library.
#include "physics.cuh"

using namespace tmd;

extern "C" {

float hc_squared_threshold(float rc) { return squared_threshold(rc); }

static int g_true_gradient = 0;  // PairParams::true_gradient for hc_pair_terms (forces the run-time term path)
void hc_set_true_gradient(int on) { g_true_gradient = on; }

// minimum image + squared distance + cutoff decision for n pairs
void hc_decide(int n, const float* pi, const float* pj, const float* box, int periodic, float s_max,
               float* w_out, float* s_out, int* in_out) {
  for (int k = 0; k < n; ++k) {
    Vec3 a = {pi[3 * k], pi[3 * k + 1], pi[3 * k + 2]}, b = {pj[3 * k], pj[3 * k + 1], pj[3 * k + 2]};
    Vec3 L = {box[0], box[1], box[2]};
    Vec3 iL = {1.0f / box[0], 1.0f / box[1], 1.0f / box[2]};
    Vec3 d = delta_ref(a, b, periodic, L, iL);
    w_out[3 * k] = d.x; w_out[3 * k + 1] = d.y; w_out[3 * k + 2] = d.z;
    s_out[k] = norm2_ref(d.x, d.y, d.z);
    in_out[k] = s_out[k] <= s_max;
  }
}

// pair energies (4 terms) and dE/dr for n pairs at squared distance s
void hc_pair_terms(int n, const float* s, const float* qq, const float* A, const float* B, unsigned terms,
                   int has_cutoff, float cutoff, int has_switch, float switch_dist, int rfa, float krf, float crf,
                   float* e_el, float* e_lj, float* e_rep, float* e_cg, float* dedr) {
  PairParams pp{};
  pp.terms = terms;
  pp.has_cutoff = has_cutoff;
  pp.cutoff = cutoff;
  pp.has_switch = has_switch;
  pp.switch_dist = switch_dist;
  pp.inv_sw_width = has_switch ? 1.0f / (cutoff - switch_dist) : 0.f;
  pp.rfa = rfa;
  pp.krf = krf;
  pp.crf = crf;
  pp.two_krf = 2.0f * krf;
  pp.true_gradient = g_true_gradient;
  for (int k = 0; k < n; ++k) {
    float a = 0, b = 0, c = 0, d = 0, rinv;
    dedr[k] = (terms == (T_LJ | T_ELEC) && has_switch && rfa && !g_true_gradient) ? pair_terms<1>(pp, s[k], qq[k], A[k], B[k], a, b, c, d, rinv) : pair_terms<0>(pp, s[k], qq[k], A[k], B[k], a, b, c, d, rinv);
    e_el[k] = a; e_lj[k] = b; e_rep[k] = c; e_cg[k] = d;
  }
}

void hc_bond(int n, const float* r, const float* k0, const float* r0, float* e, float* dedr) {
  for (int k = 0; k < n; ++k) bond_term<float>(r[k], k0[k], r0[k], e[k], dedr[k]);
}

void hc_angle(int n, const float* r21, const float* r23, const float* k0, const float* th0, float* e, float* f) {
  for (int k = 0; k < n; ++k) {
    Vec3 a = {r21[3 * k], r21[3 * k + 1], r21[3 * k + 2]}, b = {r23[3 * k], r23[3 * k + 1], r23[3 * k + 2]};
    Vec3 f0, f1, f2;
    e[k] = angle_term(a, b, k0[k], th0[k], f0, f1, f2);
    float* o = f + 9 * k;
    o[0] = f0.x; o[1] = f0.y; o[2] = f0.z; o[3] = f1.x; o[4] = f1.y; o[5] = f1.z; o[6] = f2.x; o[7] = f2.y; o[8] = f2.z;
  }
}

void hc_torsion(int n, const float* r12, const float* r23, const float* r34, const int* term_ptr, const float* terms,
                int amber, float* e, float* f) {
  for (int k = 0; k < n; ++k) {
    Vec3 a = {r12[3 * k], r12[3 * k + 1], r12[3 * k + 2]}, b = {r23[3 * k], r23[3 * k + 1], r23[3 * k + 2]},
         c = {r34[3 * k], r34[3 * k + 1], r34[3 * k + 2]};
    TorsionGeom<float> g = torsion_geom(a, b, c);
    float ee = 0, coef = 0;
    for (int m = term_ptr[k]; m < term_ptr[k + 1]; ++m)
      torsion_term(g.phi, terms[3 * m], terms[3 * m + 1], terms[3 * m + 2], amber, ee, coef);
    Vec3 f0, f1, f2, f3;
    torsion_forces(g, coef, f0, f1, f2, f3);
    e[k] = ee;
    float* o = f + 12 * k;
    o[0] = f0.x; o[1] = f0.y; o[2] = f0.z; o[3] = f1.x; o[4] = f1.y; o[5] = f1.z;
    o[6] = f2.x; o[7] = f2.y; o[8] = f2.z; o[9] = f3.x; o[10] = f3.y; o[11] = f3.z;
  }
}

// ---- fixed-point separations (k_pair_fx) ---------------------------------------------
// class per pair: 0 outside, 1 inside, 2 in the decision band (kernel re-does the reference
// arithmetic); w/s are the fixed-point values.  Margin coefficients exactly as finalize() does.
void hc_fx_decide(int n, const float* pi, const float* pj, const float* box, float s_max, double rmax, float pmax,
                  float* w_out, float* s_out, int* cls_out, float* margin_out) {
  double lmax = 0, c0, c1, inv[3];
  float unit[3];
  for (int k = 0; k < 3; ++k) {
    unit[k] = (float)((double)box[k] / 4294967296.0);
    inv[k] = 4294967296.0 / (double)box[k];
    if (box[k] > lmax) lmax = box[k];
  }
  fx_margin(rmax, lmax, &c0, &c1);
  const float fc0 = (float)(c0 * 1.0000002), fc1 = (float)(c1 * 1.0000002);
  const float margin = fmaf(fc1, pmax, fc0);
  *margin_out = margin;
  const float s_hi = s_max + margin, s_lo = s_max - margin;
  for (int q = 0; q < n; ++q) {
    float w[3];
    for (int k = 0; k < 3; ++k)
      w[k] = fx_delta(fx_encode(pi[3 * q + k], inv[k]), fx_encode(pj[3 * q + k], inv[k]), unit[k]);
    const float s = fmaf(w[2], w[2], fmaf(w[1], w[1], w[0] * w[0]));
    w_out[3 * q] = w[0]; w_out[3 * q + 1] = w[1]; w_out[3 * q + 2] = w[2];
    s_out[q] = s;
    cls_out[q] = (s < s_lo) ? 1 : ((s <= s_hi) ? 2 : 0);
  }
}

// Host emulation of the pair kernels' VALUE arithmetic over a given list of in-cutoff pairs
// (i<j, LJ+switch+RF or any term set): variant 0 = k_pair (reference-rounded separation,
// sub_err across the boundary), 1 = k_pair_fx (fixed-point separation).  Forces are
// accumulated in fp32 like the kernel (different order).  For error budgeting only.
void hc_pair_forces(int variant, int natoms, int npairs, const int* pairs, const float* pos, const float* qs,
                    const int* type, int ntypes, const float* AB, const float* box, unsigned terms, float cutoff,
                    int has_switch, float switch_dist, int rfa, float krf, float crf, float* forces) {
  PairParams pp{};
  pp.terms = terms;
  pp.has_cutoff = 1;
  pp.cutoff = cutoff;
  pp.has_switch = has_switch;
  pp.switch_dist = switch_dist;
  pp.inv_sw_width = has_switch ? 1.0f / (cutoff - switch_dist) : 0.f;
  pp.rfa = rfa;
  pp.krf = krf;
  pp.crf = crf;
  pp.two_krf = 2.0f * krf;
  const bool mode1 = (terms == (T_LJ | T_ELEC) && has_switch && rfa);
  float unit[3], iL[3];
  double inv[3];
  for (int k = 0; k < 3; ++k) {
    unit[k] = (float)((double)box[k] / 4294967296.0);
    inv[k] = 4294967296.0 / (double)box[k];
    iL[k] = 1.0f / box[k];
  }
  for (int e = 0; e < natoms * 3; ++e) forces[e] = 0.f;
  if (variant == 3) {  // k_pair2_open: no box, float records, two pairs per packed evaluation
    const SwitchConsts sc = make_switch_consts(pp);
    for (int q = 0; q < npairs; q += 2) {
      const int q1 = (q + 1 < npairs) ? q + 1 : q;
      const int ia[2] = {pairs[2 * q], pairs[2 * q1]}, ja[2] = {pairs[2 * q + 1], pairs[2 * q1 + 1]};
      F2 w[3];
      for (int k = 0; k < 3; ++k) w[k] = f2_add(f2(pos[3 * ia[0] + k], pos[3 * ia[1] + k]), f2(-pos[3 * ja[0] + k], -pos[3 * ja[1] + k]));
      const F2 s = f2_fma(w[2], w[2], f2_fma(w[1], w[1], f2_mul(w[0], w[0])));
      const float* ab0 = AB + 2 * (type[ia[0]] * ntypes + type[ja[0]]);
      const float* ab1 = AB + 2 * (type[ia[1]] * ntypes + type[ja[1]]);
      const F2 nqq = f2(-(qs[ia[0]] * qs[ja[0]]), -(qs[ia[1]] * qs[ja[1]]));
      const F2 nc = pair_coef2(sc, s, nqq, f2(ab0[0], ab1[0]), f2(ab0[1], ab1[1]), f2(rsqrt_seed(s.x), rsqrt_seed(s.y)),
                               f2(neg_rcp_seed(s.x), neg_rcp_seed(s.y)));
      const float ncs[2] = {nc.x, (q1 == q) ? 0.f : nc.y};
      for (int h = 0; h < 2; ++h)
        for (int k = 0; k < 3; ++k) {
          const float wk = h == 0 ? w[k].x : w[k].y;
          forces[3 * ia[h] + k] += wk * ncs[h];
          forces[3 * ja[h] + k] -= wk * ncs[h];
        }
    }
    return;
  }
  if (variant == 2) {  // k_pair_fx2: fixed-point separations, two pairs per packed evaluation
    const SwitchConsts sc = make_switch_consts(pp);
    for (int q = 0; q < npairs; q += 2) {
      const int q1 = (q + 1 < npairs) ? q + 1 : q;
      const int ia[2] = {pairs[2 * q], pairs[2 * q1]}, ja[2] = {pairs[2 * q + 1], pairs[2 * q1 + 1]};
      F2 w[3];
      for (int k = 0; k < 3; ++k)
        w[k] = f2_mul(f2((float)(int32_t)((uint32_t)fx_encode(pos[3 * ia[0] + k], inv[k]) - (uint32_t)fx_encode(pos[3 * ja[0] + k], inv[k])),
                         (float)(int32_t)((uint32_t)fx_encode(pos[3 * ia[1] + k], inv[k]) - (uint32_t)fx_encode(pos[3 * ja[1] + k], inv[k]))),
                      f2(unit[k]));
      const F2 s = f2_fma(w[2], w[2], f2_fma(w[1], w[1], f2_mul(w[0], w[0])));
      const float* ab0 = AB + 2 * (type[ia[0]] * ntypes + type[ja[0]]);
      const float* ab1 = AB + 2 * (type[ia[1]] * ntypes + type[ja[1]]);
      const F2 nqq = f2(-(qs[ia[0]] * qs[ja[0]]), -(qs[ia[1]] * qs[ja[1]]));
      const F2 nc = pair_coef2(sc, s, nqq, f2(ab0[0], ab1[0]), f2(ab0[1], ab1[1]), f2(rsqrt_seed(s.x), rsqrt_seed(s.y)),
                               f2(neg_rcp_seed(s.x), neg_rcp_seed(s.y)));
      const float ncs[2] = {nc.x, (q1 == q) ? 0.f : nc.y};
      for (int h = 0; h < 2; ++h)
        for (int k = 0; k < 3; ++k) {
          const float wk = h == 0 ? w[k].x : w[k].y;
          forces[3 * ia[h] + k] += wk * ncs[h];
          forces[3 * ja[h] + k] -= wk * ncs[h];
        }
    }
    return;
  }
  for (int q = 0; q < npairs; ++q) {
    const int i = pairs[2 * q], j = pairs[2 * q + 1];
    float w[3], s;
    if (variant == 1) {
      for (int k = 0; k < 3; ++k) w[k] = fx_delta(fx_encode(pos[3 * i + k], inv[k]), fx_encode(pos[3 * j + k], inv[k]), unit[k]);
      s = fmaf(w[2], w[2], fmaf(w[1], w[1], w[0] * w[0]));
    } else {
      bool straddle = false;
      float d0[3], img[3];
      for (int k = 0; k < 3; ++k) {
        float r;
        d0[k] = sub_rn(pos[3 * i + k], pos[3 * j + k]);
        w[k] = min_image_exact(d0[k], box[k], iL[k], r);
        img[k] = r;
        straddle |= (r != 0.f);
      }
      s = norm2_ref(w[0], w[1], w[2]);
      if (straddle) {
        for (int k = 0; k < 3; ++k) w[k] = straddle_value(pos[3 * i + k], pos[3 * j + k], d0[k], box[k], img[k]);
        s = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
      }
    }
    const float* ab = AB + 2 * (type[i] * ntypes + type[j]);
    float a = 0, b = 0, c = 0, d = 0, rinv;
    const float dedr = mode1 ? pair_terms<1>(pp, s, qs[i] * qs[j], ab[0], ab[1], a, b, c, d, rinv)
                             : pair_terms<0>(pp, s, qs[i] * qs[j], ab[0], ab[1], a, b, c, d, rinv);
    const float cf = dedr * rinv;
    for (int k = 0; k < 3; ++k) {
      forces[3 * i + k] -= w[k] * cf;
      forces[3 * j + k] += w[k] * cf;
    }
  }
}

// packed two-partner coefficient of k_pair_fx2: returns c = (dE/dr)/r for n pairs (n even)
void hc_pair_coef2(int n, const float* s, const float* qq, const float* A, const float* B, float cutoff,
                   float switch_dist, float krf, float* c_out) {
  PairParams pp{};
  pp.cutoff = cutoff;
  pp.has_switch = 1;
  pp.switch_dist = switch_dist;
  pp.inv_sw_width = 1.0f / (cutoff - switch_dist);
  pp.rfa = 1;
  pp.krf = krf;
  pp.two_krf = 2.0f * krf;
  const SwitchConsts sc = make_switch_consts(pp);
  for (int k = 0; k + 1 < n; k += 2) {
    const F2 ss = f2(s[k], s[k + 1]);
    const F2 nc = pair_coef2(sc, ss, f2(-qq[k], -qq[k + 1]), f2(A[k], A[k + 1]), f2(B[k], B[k + 1]),
                             f2(rsqrt_seed(ss.x), rsqrt_seed(ss.y)), f2(neg_rcp_seed(ss.x), neg_rcp_seed(ss.y)));
    c_out[k] = -nc.x;
    c_out[k + 1] = -nc.y;
  }
}

// the same with the energies of the ENERGY instantiation
void hc_pair_coef2e(int n, const float* s, const float* qq, const float* A, const float* B, float cutoff,
                    float switch_dist, float krf, float crf, float* c_out, float* elj_out, float* eel_out) {
  PairParams pp{};
  pp.cutoff = cutoff;
  pp.has_switch = 1;
  pp.switch_dist = switch_dist;
  pp.inv_sw_width = 1.0f / (cutoff - switch_dist);
  pp.rfa = 1;
  pp.krf = krf;
  pp.crf = crf;
  pp.two_krf = 2.0f * krf;
  const SwitchConsts sc = make_switch_consts(pp);
  for (int k = 0; k + 1 < n; k += 2) {
    const F2 ss = f2(s[k], s[k + 1]);
    F2 elj, neel;
    const F2 nc = pair_coef2<true>(sc, ss, f2(-qq[k], -qq[k + 1]), f2(A[k], A[k + 1]), f2(B[k], B[k + 1]),
                                   f2(rsqrt_seed(ss.x), rsqrt_seed(ss.y)), f2(neg_rcp_seed(ss.x), neg_rcp_seed(ss.y)), elj, neel);
    c_out[k] = -nc.x;
    c_out[k + 1] = -nc.y;
    elj_out[k] = elj.x;
    elj_out[k + 1] = elj.y;
    eel_out[k] = -neel.x;
    eel_out[k + 1] = -neel.y;
  }
}

// general form: with / without switching and reaction field (make_switch_consts covers all four)
void hc_pair_coef2g(int n, const float* s, const float* qq, const float* A, const float* B, float cutoff, int has_switch,
                    float switch_dist, int rfa, float krf, float crf, float* c_out, float* elj_out, float* eel_out) {
  PairParams pp{};
  pp.cutoff = cutoff;
  pp.has_switch = has_switch;
  pp.switch_dist = has_switch ? switch_dist : 0.f;
  pp.inv_sw_width = has_switch ? 1.0f / (cutoff - switch_dist) : 0.f;
  pp.rfa = rfa;
  pp.krf = krf;
  pp.crf = crf;
  pp.two_krf = 2.0f * krf;
  const SwitchConsts sc = make_switch_consts(pp);
  for (int k = 0; k + 1 < n; k += 2) {
    const F2 ss = f2(s[k], s[k + 1]);
    F2 elj, neel;
    const F2 nc = pair_coef2<true>(sc, ss, f2(-qq[k], -qq[k + 1]), f2(A[k], A[k + 1]), f2(B[k], B[k + 1]),
                                   f2(rsqrt_seed(ss.x), rsqrt_seed(ss.y)), f2(neg_rcp_seed(ss.x), neg_rcp_seed(ss.y)), elj, neel);
    c_out[k] = -nc.x;
    c_out[k + 1] = -nc.y;
    elj_out[k] = elj.x;
    elj_out[k + 1] = elj.y;
    eel_out[k] = -neel.x;
    eel_out[k + 1] = -neel.y;
  }
}

// ---- Wrapper.wrap: the arithmetic of csrc/wrap.cuh k_wrap, lane for lane -----------------
void hc_wrap(int natoms, int ngroups, const int* group_ptr, const int* group_atoms, int nrep, float* pos,
             const float* box /* (R,3,3) */) {
  bool allzero = true;
  for (int r = 0; r < nrep; ++r)
    for (int d = 0; d < 3; ++d) allzero &= (box[r * 9 + d * 4] == 0.f);
  if (allzero) return;
  for (int r = 0; r < nrep; ++r) {
    float* p = pos + (size_t)r * natoms * 3;
    for (int g = 0; g < ngroups; ++g) {
      const int b = group_ptr[g], n = group_ptr[g + 1] - b;
      if (n <= 0) continue;
      float sum[3] = {0.f, 0.f, 0.f};
      if (n <= 32) {
        for (int m = 0; m < n; ++m)
          for (int d = 0; d < 3; ++d) {
            const float t = p[(size_t)group_atoms[b + m] * 3 + d];
            sum[d] = (m == 0) ? t : add_rn(sum[d], t);
          }
      } else {
        float lane[32][3];
        for (int l = 0; l < 32; ++l)
          for (int d = 0; d < 3; ++d) {
            float acc = 0.f;
            for (int e = l; e < n; e += 32) acc = add_rn(acc, p[(size_t)group_atoms[b + e] * 3 + d]);
            lane[l][d] = acc;
          }
        for (int o = 16; o; o >>= 1) {
          float nxt[32][3];
          for (int l = 0; l < 32; ++l)
            for (int d = 0; d < 3; ++d) nxt[l][d] = add_rn(lane[l][d], lane[l ^ o][d]);
          for (int l = 0; l < 32; ++l)
            for (int d = 0; d < 3; ++d) lane[l][d] = nxt[l][d];
        }
        for (int d = 0; d < 3; ++d) sum[d] = lane[0][d];
      }
      for (int d = 0; d < 3; ++d) {
        const float off = wrap_offset(sum[d], n, box[r * 9 + d * 4]);
        for (int e = 0; e < n; ++e) {
          float& x = p[(size_t)group_atoms[b + e] * 3 + d];
          x = sub_rn(x, off);
        }
      }
    }
  }
}

}  // extern "C"
