// ptx.cuh -- the handful of inline-PTX idioms the kernels use, behind small wrappers.
//
// Device: exactly the instruction wanted (shared/global loads by explicit address, register
// pinning against re-materialisation, system-scope release/acquire).  Host: plain C++ with the same
// meaning -- used only when the kernels are compiled for the SIMT interpreter of tests/simt, which
// runs the real kernel code on the CPU to check its logic without a GPU (test infrastructure; the
// product never executes these alternatives).
#pragma once
#include <stdint.h>

namespace tmd {

#if defined(__CUDA_ARCH__)
#define TMD_PIN_R(x) asm volatile("" : "+r"(x))  // keep a 32-bit integer / float / 64-bit value in a register:
#define TMD_PIN_F(x) asm volatile("" : "+f"(x))  // the compiler can no longer re-derive it inside a loop
#define TMD_PIN_L(x) asm volatile("" : "+l"(x))
#define TMD_KEEP_STORES() asm volatile("" ::: "memory")  // shared-memory stores that are read back through lds_*

typedef unsigned smem_addr;  // address in the CTA's shared window
__device__ __forceinline__ smem_addr smem_address(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(smem_addr a) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
template <int OFFSET>
__device__ __forceinline__ float lds_f32_at(smem_addr a) {  // the same entry of a second plane OFFSET bytes further on
  float v;
  asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFFSET));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(smem_addr a) {
  float2 v;
  asm("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
// (volatile: the staged tile is rewritten between barriers, the load must not be merged with an earlier one)
__device__ __forceinline__ float4 lds_f32x4(smem_addr a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 ldg_s32x4(unsigned long long addr) {
  int4 v;
  asm("ld.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr));
  return v;
}
// two floats as one 64-bit register pair (the operand form of the packed fp32x2 instructions); pinning the packed
// value keeps the pair adjacent across a loop instead of re-assembling it with two moves per use
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// base + a * b as one IMAD.WIDE.U32 (the address of a 16-byte record from its index)
__device__ __forceinline__ unsigned long long mad_wide_u32(unsigned a, unsigned b, unsigned long long c) {
  unsigned long long d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void stg_u32(unsigned long long addr, int v) {
  asm volatile("st.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- bulk (TMA) copies global -> shared with an mbarrier, fire-and-forget vector reductions ----
// One elected lane announces the bytes and starts the copy; the warp then waits on the barrier's phase.
__device__ __forceinline__ void mbar_init(smem_addr bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(smem_addr bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// size and both addresses multiples of 16 bytes
__device__ __forceinline__ void bulk_g2s(smem_addr dst, const void* src, unsigned bytes, smem_addr bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(smem_addr bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ unsigned lds_u8(smem_addr a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u32(smem_addr a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
// f[0..2] += (x, y, z) as one 16-byte reduction at L2 (the fourth component adds zero)
__device__ __forceinline__ void red_add_f32x4(float4* p, float x, float y, float z) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
}
#else
#define TMD_PIN_R(x) ((void)0)
#define TMD_PIN_F(x) ((void)0)
#define TMD_PIN_L(x) ((void)0)
#define TMD_KEEP_STORES() ((void)0)

typedef const char* smem_addr;
inline smem_addr smem_address(const void* p) { return static_cast<const char*>(p); }
inline float lds_f32(smem_addr a) { return *reinterpret_cast<const float*>(a); }
template <int OFFSET>
inline float lds_f32_at(smem_addr a) { return *reinterpret_cast<const float*>(a + OFFSET); }
inline float2 lds_f32x2(smem_addr a) { return *reinterpret_cast<const float2*>(a); }
inline float4 lds_f32x4(smem_addr a) { return *reinterpret_cast<const float4*>(a); }
inline int4 ldg_s32x4(unsigned long long addr) { return *reinterpret_cast<const int4*>(addr); }
inline unsigned long long pack_f32x2(float lo, float hi) {
  unsigned a, b;
  memcpy(&a, &lo, 4);
  memcpy(&b, &hi, 4);
  return (unsigned long long)a | ((unsigned long long)b << 32);
}
inline void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  const unsigned a = (unsigned)v, b = (unsigned)(v >> 32);
  memcpy(&lo, &a, 4);
  memcpy(&hi, &b, 4);
}
inline unsigned long long mad_wide_u32(unsigned a, unsigned b, unsigned long long c) { return (unsigned long long)a * b + c; }
inline void stg_u32(unsigned long long addr, int v) { *reinterpret_cast<int*>(addr) = v; }
inline void st_release_sys(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned ld_acquire_sys(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

inline void mbar_init(smem_addr, unsigned) {}
inline void fence_mbar_init() {}
inline void mbar_expect_tx(smem_addr, unsigned) {}
inline void bulk_g2s(smem_addr dst, const void* src, unsigned bytes, smem_addr) { memcpy(const_cast<char*>(dst), src, bytes); }
inline void mbar_wait(smem_addr, unsigned) {}  // (the host copy above is synchronous)
inline unsigned lds_u32(smem_addr a) { return *reinterpret_cast<const unsigned*>(a); }
inline unsigned lds_u8(smem_addr a) { return *reinterpret_cast<const unsigned char*>(a); }
inline void red_add_f32x4(float4* p, float x, float y, float z) {  // (interpreter threads run one at a time)
  p->x += x;
  p->y += y;
  p->z += z;
}
#endif

}  // namespace tmd
