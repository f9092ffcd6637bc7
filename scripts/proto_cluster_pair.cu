// Prototype + microbenchmark of the cluster half-list pair kernel (round 2) before it goes into the
// library: i-clusters of C cell-sorted atoms (one warp each, the C atoms packed two by two into the halves of fp32x2
// operations), lane = one j-atom of the cluster's list, Newton's third law with one reduction per (cluster, j) entry.
// Builds the list on the host, checks the forces against a double-precision all-neighbours loop, times the variants.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I torchmd_b200/csrc -o /tmp/proto scripts/proto_cluster_pair.cu
//   /tmp/proto [n_waters] [rlist]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "physics.cuh"

using namespace tmd;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e__ = (x);                                                                 \
    if (e__ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);     \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int WARPS = 8;
constexpr int MAXT = 16;
static double JIT = 0.4;  // lattice jitter (A); PROTO_JITTER overrides

struct KParams {
  float L[3], iL[3];
  float s_max;
  SwitchConsts sc;
  int ntypes;
  int nslots;  // dummy record index
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int RED>
__device__ __forceinline__ void red_add(float4* f, int j, float x, float y, float z) {
  float* p = reinterpret_cast<float*>(f + j);
  if (RED == 1) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 1), "f"(y) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 2), "f"(z) : "memory");
  } else if (RED == 2) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
  } else if (RED == 3) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 2), "f"(z) : "memory");
  }
}

// pair_coef2 with -1/r^2 = -(1/r)^2 from the refined 1/r (no second hardware seed)
__device__ __forceinline__ F2 pair_coef2b(const SwitchConsts& c, F2 s, F2 nqq, F2 A, F2 B, F2 y) {
  const F2 t = f2_mul(s, y);
  const F2 u = f2_fma(f2_mul(t, f2(-0.5f)), y, f2(0.5f));
  const F2 rinv = f2_fma(y, u, y);
  const F2 r = f2_mul(s, rinv);
  const F2 nr2 = f2_mul(f2_mul(rinv, f2(-1.0f)), rinv);
  const F2 nr6 = f2_mul(f2_mul(nr2, nr2), nr2);
  const F2 a12 = f2_mul(f2_mul(A, nr6), nr6);
  const F2 nb6 = f2_mul(B, nr6);
  const F2 e = f2_add(a12, nb6);
  const F2 nf = f2_mul(f2_fma(nb6, f2(6.0f), f2_mul(a12, f2(12.0f))), rinv);
  F2 tt = f2_mul(f2_add(r, f2(c.neg_switch_dist)), f2(c.inv_sw_width));
  tt = f2(fmaxf(tt.x, 0.0f), fmaxf(tt.y, 0.0f));
  const F2 t2 = f2_mul(tt, tt);
  const F2 sw = f2_fma(f2_mul(t2, tt), f2_fma(tt, f2_fma(tt, f2(-6.0f), f2(15.0f)), f2(-10.0f)), f2(1.0f));
  const F2 ndsw = f2_mul(t2, f2_fma(tt, f2_fma(tt, f2(c.d1), f2(c.d2)), f2(c.d3)));
  const F2 nfsw = f2_fma(sw, nf, f2_mul(f2_mul(e, ndsw), rinv));
  const F2 ndedr = f2_fma(nqq, f2_fma(f2(c.two_krf), r, nr2), nfsw);
  return f2_mul(ndedr, rinv);
}

// C atoms per i-cluster (2, 4 or 8); RED: 0 no j reduction (compute only), 1 three scalar reds, 2 one red.v4, 3 v2 + scalar
// VAR bit 0: no per-block branch (both i-pair blocks straight-line: the compiler may interleave them);
//     bit 1: -1/r^2 from the refined 1/r (one MUFU less per pair)
template <int C, int RED, int MINB, int VAR>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_cpair(int nclusters, const float4* __restrict__ xq, const int* __restrict__ type_s, const float4* __restrict__ centre,
        const int* __restrict__ cl_ptr, const unsigned* __restrict__ entries, const int* __restrict__ sp_ptr,
        const uint2* __restrict__ sp_entries, const float2* __restrict__ AB, KParams P, float4* __restrict__ f_s) {
  constexpr int H = C / 2;
  __shared__ float4 tab[WARPS][MAXT][H];  // (A_i0j, A_i1j, B_i0j, B_i1j) per partner type and i-pair
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * WARPS + w;
  if (c >= nclusters) return;
  const int s0 = c * C;
  F2 XI[H], YI[H], ZI[H], NQI[H];
#pragma unroll
  for (int p = 0; p < H; ++p) {
    const float4 a = xq[s0 + 2 * p], b = xq[s0 + 2 * p + 1];
    XI[p] = f2(a.x, b.x);
    YI[p] = f2(a.y, b.y);
    ZI[p] = f2(a.z, b.z);
    NQI[p] = f2(-a.w, -b.w);
  }
  for (int e = lane; e < P.ntypes * H; e += 32) {
    const int t = e / H, p = e % H;
    const int t0 = type_s[s0 + 2 * p], t1 = type_s[s0 + 2 * p + 1];
    const float2 v0 = AB[t0 * P.ntypes + t], v1 = AB[t1 * P.ntypes + t];
    tab[w][t][p] = make_float4(v0.x, v1.x, v0.y, v1.y);
  }
  __syncwarp();
  const float4 ctr = centre[c];
  F2 FX[H], FY[H], FZ[H];
#pragma unroll
  for (int p = 0; p < H; ++p) FX[p] = FY[p] = FZ[p] = f2(0.f);
  const float magic = 12582912.0f;

  auto body = [&](unsigned entry, unsigned mask, const float4 pj) {
    const unsigned tj = entry >> 24;
    // image shift of this (cluster, j): the same integer for every atom of a compact cluster
    const float nx = __fadd_rn(__fmaf_rn(ctr.x - pj.x, P.iL[0], magic), -magic);
    const float ny = __fadd_rn(__fmaf_rn(ctr.y - pj.y, P.iL[1], magic), -magic);
    const float nz = __fadd_rn(__fmaf_rn(ctr.z - pj.z, P.iL[2], magic), -magic);
    const F2 mLx = f2(-__fmul_rn(P.L[0], nx)), mLy = f2(-__fmul_rn(P.L[1], ny)), mLz = f2(-__fmul_rn(P.L[2], nz));
    const F2 mxj = f2(-pj.x), myj = f2(-pj.y), mzj = f2(-pj.z);
    F2 GX = f2(0.f), GY = f2(0.f), GZ = f2(0.f);
#pragma unroll
    for (int p = 0; p < H; ++p) {
      // the reference's rounding chain: fl(fl(xi - xj) - fl(L n)), then fma(z,z,fma(y,y,x*x))
      const F2 dx = f2_add(f2_add(XI[p], mxj), mLx);
      const F2 dy = f2_add(f2_add(YI[p], myj), mLy);
      const F2 dz = f2_add(f2_add(ZI[p], mzj), mLz);
      const F2 s = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
      const bool in0 = s.x <= P.s_max && ((mask >> (2 * p)) & 1u);
      const bool in1 = s.y <= P.s_max && ((mask >> (2 * p + 1)) & 1u);
      if ((VAR & 1) || in0 || in1) {
        const float4 ab = tab[w][tj][p];
        const F2 nqq = f2_mul(NQI[p], f2(pj.w));
        F2 nc;
        if (VAR & 2) nc = pair_coef2b(P.sc, s, nqq, f2(ab.x, ab.y), f2(ab.z, ab.w), f2(rsqrt_seed(s.x), rsqrt_seed(s.y)));
        else nc = pair_coef2(P.sc, s, nqq, f2(ab.x, ab.y), f2(ab.z, ab.w), f2(rsqrt_seed(s.x), rsqrt_seed(s.y)),
                           f2(neg_rcp_seed(s.x), neg_rcp_seed(s.y)));
        nc = f2(in0 ? nc.x : 0.f, in1 ? nc.y : 0.f);
        FX[p] = f2_fma(dx, nc, FX[p]);
        FY[p] = f2_fma(dy, nc, FY[p]);
        FZ[p] = f2_fma(dz, nc, FZ[p]);
        GX = f2_fma(dx, nc, GX);
        GY = f2_fma(dy, nc, GY);
        GZ = f2_fma(dz, nc, GZ);
      }
    }
    const float gx = -(GX.x + GX.y), gy = -(GY.x + GY.y), gz = -(GZ.x + GZ.y);
    if (RED != 0 && (gx != 0.f || gy != 0.f || gz != 0.f)) red_add<RED>(f_s, entry & 0xffffffu, gx, gy, gz);
  };

  // special entries: partners with an excluded pair in this cluster, and the cluster's own atoms (triangle)
  for (int e = sp_ptr[c] + lane; e < sp_ptr[c + 1]; e += 32) {
    const uint2 se = sp_entries[e];
    body(se.x, se.y, xq[se.x & 0xffffffu]);
  }
  // plain entries, padded per cluster to a multiple of 32 with entries of the dummy record
  const int e1 = cl_ptr[c + 1];
  int e = cl_ptr[c] + lane;
  unsigned en = 0;
  float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e < e1) {
    en = entries[e];
    pj = xq[en & 0xffffffu];
  }
  while (e < e1) {
    const int e_next = e + 32;
    unsigned en_next = 0;
    float4 pj_next = pj;
    if (e_next < e1) {
      en_next = entries[e_next];
      pj_next = xq[en_next & 0xffffffu];
    }
    body(en, 0xffffffffu, pj);
    en = en_next;
    pj = pj_next;
    e = e_next;
  }
  // i forces: reduce over the lanes, one reduction per atom
#pragma unroll
  for (int p = 0; p < H; ++p) {
    const float ax = warp_sum(FX[p].x), ay = warp_sum(FY[p].x), az = warp_sum(FZ[p].x);
    const float bx = warp_sum(FX[p].y), by = warp_sum(FY[p].y), bz = warp_sum(FZ[p].y);
    if (lane == 0) {
      red_add<RED == 0 ? 1 : RED>(f_s, s0 + 2 * p, ax, ay, az);
      red_add<RED == 0 ? 1 : RED>(f_s, s0 + 2 * p + 1, bx, by, bz);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct Host {
  int N, nslots, nclusters;
  float L;
  std::vector<float4> xq;   // sorted records (+ dummy)
  std::vector<int> type_s, mol_s;
  std::vector<float4> centre;
  std::vector<int> cl_ptr, sp_ptr;
  std::vector<unsigned> entries;
  std::vector<uint2> sp_entries;
  long long slots_evaluated = 0, in_cut = 0;
};

static void make_water(int nw, float& L, std::vector<float>& pos, std::vector<float>& q, std::vector<int>& type, std::vector<int>& mol) {
  std::mt19937 rng(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  L = (float)cbrt(nw / 0.0334);
  int m = (int)ceil(cbrt((double)nw) - 1e-9);
  std::vector<int> sites(m * m * m);
  for (int i = 0; i < m * m * m; ++i) sites[i] = i;
  std::shuffle(sites.begin(), sites.end(), rng);
  sites.resize(nw);
  std::sort(sites.begin(), sites.end());
  const double roh = 0.9572, ang = 104.52 * M_PI / 180.0;
  const double loc[3][3] = {{0, 0, 0}, {roh * sin(ang / 2), roh * cos(ang / 2), 0}, {-roh * sin(ang / 2), roh * cos(ang / 2), 0}};
  pos.resize((size_t)nw * 9);
  q.resize(nw * 3);
  type.resize(nw * 3);
  mol.resize(nw * 3);
  std::normal_distribution<double> G(0.0, 1.0);
  for (int k = 0; k < nw; ++k) {
    const int s = sites[k];
    const int gx = s / (m * m), gy = (s / m) % m, gz = s % m;
    double o[3] = {(gx + 0.5) * L / m + (U(rng) - 0.5) * JIT, (gy + 0.5) * L / m + (U(rng) - 0.5) * JIT, (gz + 0.5) * L / m + (U(rng) - 0.5) * JIT};
    double qv[4], nn = 0;
    for (double& v : qv) {
      v = G(rng);
      nn += v * v;
    }
    nn = sqrt(nn);
    const double ww = qv[0] / nn, x = qv[1] / nn, y = qv[2] / nn, z = qv[3] / nn;
    const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)},
                            {2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)},
                            {2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)}};
    for (int a = 0; a < 3; ++a) {
      for (int d = 0; d < 3; ++d) pos[(size_t)(k * 3 + a) * 3 + d] = (float)(o[d] + R[d][0] * loc[a][0] + R[d][1] * loc[a][1] + R[d][2] * loc[a][2]);
      q[k * 3 + a] = a == 0 ? -0.834f : 0.417f;
      type[k * 3 + a] = a == 0 ? 1 : 0;
      mol[k * 3 + a] = k;
    }
  }
}

template <int C>
static void build(Host& H, float L, const std::vector<float>& pos, const std::vector<float>& q, const std::vector<int>& type,
                  const std::vector<int>& mol, double rl, double cellw, float s_max) {
  const int N = (int)q.size();
  H.N = N;
  H.L = L;
  const int nc = std::max(1, (int)floor(L / cellw));
  const double w = L / nc;
  // snake order of the cells; atoms inside a cell by x along the row's direction
  std::vector<std::vector<int>> cells((size_t)nc * nc * nc);
  auto fold = [&](double v) { v -= L * floor(v / L); return v; };
  for (int i = 0; i < N; ++i) {
    int cx = std::min(nc - 1, (int)(fold(pos[i * 3]) / w)), cy = std::min(nc - 1, (int)(fold(pos[i * 3 + 1]) / w)), cz = std::min(nc - 1, (int)(fold(pos[i * 3 + 2]) / w));
    cells[((size_t)cz * nc + cy) * nc + cx].push_back(i);
  }
  std::vector<int> order;
  order.reserve(N);
  for (int cz = 0; cz < nc; ++cz)
    for (int yy = 0; yy < nc; ++yy) {
      const int cy = (cz & 1) ? nc - 1 - yy : yy;
      const int row = cz * nc + yy;
      for (int xx = 0; xx < nc; ++xx) {
        const int cx = (row & 1) ? nc - 1 - xx : xx;
        auto& v = cells[((size_t)cz * nc + cy) * nc + cx];
        std::sort(v.begin(), v.end(), [&](int a, int b) {
          const double xa = fold(pos[a * 3]), xb = fold(pos[b * 3]);
          return (row & 1) ? (xa > xb || (xa == xb && a < b)) : (xa < xb || (xa == xb && a < b));
        });
        for (int a : v) order.push_back(a);
      }
    }
  H.nclusters = (N + C - 1) / C;
  H.nslots = H.nclusters * C;
  const float qs = (float)sqrt(332.06371307417066);
  H.xq.assign(H.nslots + 1, make_float4(NAN, NAN, NAN, 0.f));
  H.type_s.assign(H.nslots + 1, 0);
  H.mol_s.assign(H.nslots + 1, -1);
  for (int s = 0; s < N; ++s) {
    const int a = order[s];
    H.xq[s] = make_float4(pos[a * 3], pos[a * 3 + 1], pos[a * 3 + 2], q[a] * qs);
    H.type_s[s] = type[a];
    H.mol_s[s] = mol[a];
  }
  // cluster bounding boxes in folded coordinates, unwrapped relative to the first atom
  std::vector<double> lo((size_t)H.nclusters * 3), hi((size_t)H.nclusters * 3);
  H.centre.resize(H.nclusters);
  for (int c = 0; c < H.nclusters; ++c) {
    double r0[3] = {H.xq[c * C].x, H.xq[c * C].y, H.xq[c * C].z};
    for (int d = 0; d < 3; ++d) lo[c * 3 + d] = 1e30, hi[c * 3 + d] = -1e30;
    for (int k = 0; k < C; ++k) {
      const float4 p = H.xq[c * C + k];
      if (std::isnan(p.x)) continue;
      const double v[3] = {p.x, p.y, p.z};
      for (int d = 0; d < 3; ++d) {
        double u = v[d] - r0[d];
        u -= L * rint(u / L);
        u += r0[d];
        lo[c * 3 + d] = std::min(lo[c * 3 + d], u);
        hi[c * 3 + d] = std::max(hi[c * 3 + d], u);
      }
    }
    H.centre[c] = make_float4((float)(0.5 * (lo[c * 3] + hi[c * 3])), (float)(0.5 * (lo[c * 3 + 1] + hi[c * 3 + 1])), (float)(0.5 * (lo[c * 3 + 2] + hi[c * 3 + 2])), 0.f);
  }
  // slot cell grid for the search
  const int ng = std::max(1, (int)floor(L / 5.0));
  const double gw = L / ng;
  std::vector<std::vector<int>> g((size_t)ng * ng * ng);
  for (int s = 0; s < N; ++s) {
    int cx = std::min(ng - 1, (int)(fold(H.xq[s].x) / gw)), cy = std::min(ng - 1, (int)(fold(H.xq[s].y) / gw)), cz = std::min(ng - 1, (int)(fold(H.xq[s].z) / gw));
    g[((size_t)cz * ng + cy) * ng + cx].push_back(s);
  }
  H.cl_ptr.assign(H.nclusters + 1, 0);
  H.sp_ptr.assign(H.nclusters + 1, 0);
  H.entries.clear();
  H.sp_entries.clear();
  std::vector<int> cand;
  for (int c = 0; c < H.nclusters; ++c) {
    H.cl_ptr[c] = (int)H.entries.size();
    H.sp_ptr[c] = (int)H.sp_entries.size();
    double ctr[3], half[3];
    for (int d = 0; d < 3; ++d) ctr[d] = 0.5 * (lo[c * 3 + d] + hi[c * 3 + d]), half[d] = 0.5 * (hi[c * 3 + d] - lo[c * 3 + d]);
    cand.clear();
    int clo[3], chi[3];
    for (int d = 0; d < 3; ++d) {
      clo[d] = (int)floor((ctr[d] - half[d] - rl) / gw);
      chi[d] = (int)floor((ctr[d] + half[d] + rl) / gw);
      if (chi[d] - clo[d] + 1 > ng) clo[d] = 0, chi[d] = ng - 1;
    }
    for (int z = clo[2]; z <= chi[2]; ++z)
      for (int y = clo[1]; y <= chi[1]; ++y)
        for (int x = clo[0]; x <= chi[0]; ++x) {
          const int zz = ((z % ng) + ng) % ng, yy = ((y % ng) + ng) % ng, xx = ((x % ng) + ng) % ng;
          for (int s : g[((size_t)zz * ng + yy) * ng + xx]) {
            if (s < (c + 1) * C) continue;  // half list: partners in later clusters only
            const double v[3] = {H.xq[s].x, H.xq[s].y, H.xq[s].z};
            double d2 = 0;
            for (int d = 0; d < 3; ++d) {
              double u = v[d] - ctr[d];
              u -= L * rint(u / L);
              const double e = std::max(fabs(u) - half[d], 0.0);
              d2 += e * e;
            }
            if (d2 < rl * rl) cand.push_back(s);
          }
        }
    std::sort(cand.begin(), cand.end());
    // the cluster's own atoms: triangle
    for (int k = 1; k < C; ++k) {
      const int s = c * C + k;
      if (s >= N) break;
      unsigned mask = 0;
      for (int i = 0; i < k; ++i)
        if (H.mol_s[c * C + i] != H.mol_s[s]) mask |= 1u << i;
      if (mask) H.sp_entries.push_back(make_uint2((unsigned)s | ((unsigned)H.type_s[s] << 24), mask));
    }
    for (int s : cand) {
      unsigned mask = 0;
      for (int i = 0; i < C; ++i)
        if (c * C + i < N && H.mol_s[c * C + i] != H.mol_s[s]) mask |= 1u << i;
      const unsigned full = (c * C + C <= N) ? ((1u << C) - 1u) : ((1u << (N - c * C)) - 1u);
      if (mask == full && full == ((1u << C) - 1u)) H.entries.push_back((unsigned)s | ((unsigned)H.type_s[s] << 24));
      else if (mask) H.sp_entries.push_back(make_uint2((unsigned)s | ((unsigned)H.type_s[s] << 24), mask));
    }
    while ((H.entries.size() - H.cl_ptr[c]) % 32) H.entries.push_back((unsigned)H.nslots);
  }
  H.cl_ptr[H.nclusters] = (int)H.entries.size();
  H.sp_ptr[H.nclusters] = (int)H.sp_entries.size();
}

// double-precision forces in sorted order (the reference's explicit formulas, forces.py:381-491)
static void reference_forces(const Host& H, const PairParams& pp, const std::vector<float2>& AB, int ntypes, std::vector<double>& F, long long& npairs) {
  const int N = H.N;
  const double L = H.L;
  F.assign((size_t)H.nslots * 3, 0.0);
  const int ng = std::max(1, (int)floor(L / 9.01));
  const double gw = L / ng;
  auto fold = [&](double v) { v -= L * floor(v / L); return v; };
  std::vector<std::vector<int>> g((size_t)ng * ng * ng);
  for (int s = 0; s < N; ++s)
    g[((size_t)std::min(ng - 1, (int)(fold(H.xq[s].z) / gw)) * ng + std::min(ng - 1, (int)(fold(H.xq[s].y) / gw))) * ng + std::min(ng - 1, (int)(fold(H.xq[s].x) / gw))].push_back(s);
  npairs = 0;
  const double rc = pp.cutoff, rs = pp.switch_dist, krf = pp.krf, crf = pp.crf;
  for (int cz = 0; cz < ng; ++cz)
    for (int cy = 0; cy < ng; ++cy)
      for (int cx = 0; cx < ng; ++cx)
        for (int a : g[((size_t)cz * ng + cy) * ng + cx])
          for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx) {
                const int zz = (cz + dz + ng) % ng, yy = (cy + dy + ng) % ng, xx = (cx + dx + ng) % ng;
                for (int b : g[((size_t)zz * ng + yy) * ng + xx]) {
                  if (b <= a || H.mol_s[a] == H.mol_s[b]) continue;
                  // decision: the fp32 reference predicate
                  const float4 pa = H.xq[a], pb = H.xq[b];
                  const float fL = (float)L, iL = 1.0f / fL;
                  if (!ref_inside(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, fL, fL, fL, iL, iL, iL, pp.s_max)) continue;
                  double d[3] = {(double)pa.x - pb.x, (double)pa.y - pb.y, (double)pa.z - pb.z};
                  double r2 = 0;
                  for (int k = 0; k < 3; ++k) {
                    d[k] -= L * rint(d[k] / L);
                    r2 += d[k] * d[k];
                  }
                  ++npairs;
                  const double r = sqrt(r2), rinv = 1.0 / r, rinv6 = rinv * rinv * rinv * rinv * rinv * rinv;
                  const float2 ab = AB[H.type_s[a] * ntypes + H.type_s[b]];
                  double e = ab.x * rinv6 * rinv6 - ab.y * rinv6;
                  double f = (-12.0 * ab.x * rinv6 * rinv6 + 6.0 * ab.y * rinv6) * rinv;
                  if (r > rs) {
                    const double t = (r - rs) / (rc - rs);
                    const double sw = 1 + t * t * t * (-10 + t * (15 - t * 6));
                    const double dsw = t * t * (-30 + t * (60 - t * 30)) / (rc - rs);
                    f = sw * f + e * dsw / r;
                  }
                  const double qq = (double)pa.w * pb.w;
                  f += qq * (2.0 * krf * r - 1.0 / r2);
                  for (int k = 0; k < 3; ++k) {
                    F[(size_t)a * 3 + k] -= d[k] * rinv * f;
                    F[(size_t)b * 3 + k] += d[k] * rinv * f;
                  }
                }
              }
  (void)crf;
}

template <int C, int RED, int MINB, int VAR>
static float run_kernel(const Host& H, const KParams& P, const float4* d_xq, const int* d_type, const float4* d_ctr, const int* d_clptr,
                        const unsigned* d_ent, const int* d_spptr, const uint2* d_sp, const float2* d_ab, float4* d_f, int reps) {
  const int blocks = (H.nclusters + WARPS - 1) / WARPS;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int k = 0; k < 3; ++k) k_cpair<C, RED, MINB, VAR><<<blocks, WARPS * 32>>>(H.nclusters, d_xq, d_type, d_ctr, d_clptr, d_ent, d_spptr, d_sp, d_ab, P, d_f);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int k = 0; k < reps; ++k) k_cpair<C, RED, MINB, VAR><<<blocks, WARPS * 32>>>(H.nclusters, d_xq, d_type, d_ctr, d_clptr, d_ent, d_spptr, d_sp, d_ab, P, d_f);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps * 1000.f;
}

template <int C>
static void experiment(int nw, double rl, double cellw) {
  float L;
  std::vector<float> pos, q;
  std::vector<int> type, mol;
  make_water(nw, L, pos, q, type, mol);
  PairParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.terms = T_LJ | T_ELEC;
  pp.periodic = 1;
  pp.has_cutoff = pp.has_switch = pp.rfa = 1;
  pp.cutoff = 9.0f;
  pp.s_max = squared_threshold(9.0f);
  pp.switch_dist = 7.5f;
  pp.inv_sw_width = 1.0f / 1.5f;
  const double eps = 78.5, den = 2 * eps + 1;
  pp.krf = (float)((1.0 / 729.0) * (eps - 1) / den);
  pp.crf = (float)((1.0 / 9.0) * 3 * eps / den);
  pp.two_krf = (float)(2.0 * (1.0 / 729.0) * (eps - 1) / den);
  const double sig[2] = {0.40001352444501237, 3.150574226831496}, ep[2] = {-0.046, -0.1521};
  std::vector<float2> AB(4);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const double s = 0.5 * (sig[a] + sig[b]), e = sqrt(ep[a] * ep[b]);
      AB[a * 2 + b] = make_float2((float)(4 * e * pow(s, 12)), (float)(4 * e * pow(s, 6)));
    }
  Host H;
  build<C>(H, L, pos, q, type, mol, rl, cellw, pp.s_max);
  std::vector<double> Fref;
  long long npairs;
  reference_forces(H, pp, AB, 2, Fref, npairs);
  const long long slots = (long long)(H.entries.size()) * C;
  printf("C=%d  N=%d L=%.2f rlist=%.2f cellw=%.2f: clusters %d, plain entries %zu (%.1f MB), special %zu, pairs in cutoff %lld, eta %.3f\n", C, H.N, L, rl, cellw,
         H.nclusters, H.entries.size(), H.entries.size() * 4e-6, H.sp_entries.size(), npairs, (double)npairs / slots);

  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      printf("   (no GPU: list statistics only)\n");
      return;
    }
  }
  KParams P;
  for (int d = 0; d < 3; ++d) P.L[d] = L, P.iL[d] = 1.0f / L;
  P.s_max = pp.s_max;
  P.sc = make_switch_consts(pp);
  P.ntypes = 2;
  P.nslots = H.nslots;
  float4 *d_xq, *d_ctr, *d_f;
  int *d_type, *d_clptr, *d_spptr;
  unsigned* d_ent;
  uint2* d_sp;
  float2* d_ab;
  CK(cudaMalloc(&d_xq, H.xq.size() * 16));
  CK(cudaMalloc(&d_ctr, H.centre.size() * 16));
  CK(cudaMalloc(&d_f, (H.nslots + 1) * 16));
  CK(cudaMalloc(&d_type, H.type_s.size() * 4));
  CK(cudaMalloc(&d_clptr, H.cl_ptr.size() * 4));
  CK(cudaMalloc(&d_spptr, H.sp_ptr.size() * 4));
  CK(cudaMalloc(&d_ent, H.entries.size() * 4 + 16));
  CK(cudaMalloc(&d_sp, H.sp_entries.size() * 8 + 16));
  CK(cudaMalloc(&d_ab, AB.size() * 8));
  CK(cudaMemcpy(d_xq, H.xq.data(), H.xq.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ctr, H.centre.data(), H.centre.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_type, H.type_s.data(), H.type_s.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_clptr, H.cl_ptr.data(), H.cl_ptr.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_spptr, H.sp_ptr.data(), H.sp_ptr.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ent, H.entries.data(), H.entries.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_sp, H.sp_entries.data(), H.sp_entries.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ab, AB.data(), AB.size() * 8, cudaMemcpyHostToDevice));

  // correctness: one launch with the vector reduction
  CK(cudaMemset(d_f, 0, (H.nslots + 1) * 16));
  k_cpair<C, 2, 3, 0><<<(H.nclusters + WARPS - 1) / WARPS, WARPS * 32>>>(H.nclusters, d_xq, d_type, d_ctr, d_clptr, d_ent, d_spptr, d_sp, d_ab, P, d_f);
  CK(cudaDeviceSynchronize());
  std::vector<float4> f(H.nslots + 1);
  CK(cudaMemcpy(f.data(), d_f, f.size() * 16, cudaMemcpyDeviceToHost));
  double emax = 0, fmax = 0, erel = 0;
  for (int s = 0; s < H.N; ++s) {
    const double v[3] = {f[s].x, f[s].y, f[s].z};
    double fn = 0;
    for (int k = 0; k < 3; ++k) fn = std::max(fn, fabs(Fref[(size_t)s * 3 + k]));
    for (int k = 0; k < 3; ++k) {
      emax = std::max(emax, fabs(v[k] - Fref[(size_t)s * 3 + k]));
      erel = std::max(erel, fabs(v[k] - Fref[(size_t)s * 3 + k]) / std::max(fn, 10.0));
      fmax = std::max(fmax, fabs(Fref[(size_t)s * 3 + k]));
    }
  }
  printf("   max |dF| vs fp64 = %.3e  (max |F| %.1f), max |dF| / max(|F_atom|, 10) = %.3e\n", emax, fmax, erel);

#define RUN(RED, MINB, VAR) printf("   RED=%d minblocks=%d var=%d: %8.1f us\n", RED, MINB, VAR, run_kernel<C, RED, MINB, VAR>(H, P, d_xq, d_type, d_ctr, d_clptr, d_ent, d_spptr, d_sp, d_ab, d_f, reps));
  const int reps = getenv("PROTO_REPS") ? atoi(getenv("PROTO_REPS")) : 20;
  RUN(2, 3, 0)
  RUN(2, 3, 1)
  RUN(2, 3, 2)
  RUN(2, 3, 3)
  RUN(2, 2, 3)
  RUN(2, 4, 3)
  RUN(0, 3, 3)
  cudaFree(d_xq); cudaFree(d_ctr); cudaFree(d_f); cudaFree(d_type); cudaFree(d_clptr); cudaFree(d_spptr); cudaFree(d_ent); cudaFree(d_sp); cudaFree(d_ab);
}

int main(int argc, char** argv) {
  const int nw = argc > 1 ? atoi(argv[1]) : 33333;
  const double rl = argc > 2 ? atof(argv[2]) : 9.5;
  if (getenv("PROTO_JITTER")) JIT = atof(getenv("PROTO_JITTER"));
  if (argc > 3) {  // one configuration (for ncu)
    experiment<4>(nw, rl, 4.0);
    return 0;
  }
  experiment<2>(nw, rl, 4.0);
  experiment<4>(nw, rl, 4.0);
  experiment<8>(nw, rl, 4.0);
  return 0;
}
