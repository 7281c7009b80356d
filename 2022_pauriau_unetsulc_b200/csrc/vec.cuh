// 8 x bf16 (16-byte) vector helpers for the bandwidth-bound NDHWC kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

struct f8 {
  float v[8];
};

__device__ __forceinline__ f8 unpack8(const uint4& u) {
  f8 r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}

__device__ __forceinline__ uint4 pack8(const f8& f) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(f.v[0], f.v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f.v[2], f.v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f.v[4], f.v[5]);
  __nv_bfloat162 d = __floats2bfloat162_rn(f.v[6], f.v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c);
  u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

// bf16 bits of a float whose value is exactly representable in bf16 (weights such as 0.25, 0.75, 0.5625)
__device__ __forceinline__ unsigned short bf16_bits_exact(float f) { return (unsigned short)(__float_as_uint(f) >> 16); }

// acc[k] += w * u[k] for the 8 bf16 values of a 16-byte vector: mixed-precision FMA (PTX fma.rn.f32.bf16 -> SASS
// FHFMA.BF16 on sm_100a): the bf16 operands are read straight from the register halves, the product is exact and the
// accumulation is fp32 — identical to unpacking to fp32 first, at half the instruction count (no shift / mask per value).
__device__ __forceinline__ void fma8_bf16(float (&acc)[8], const uint4& u, unsigned short w) {
  asm("{\n\t.reg .b16 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
      "mov.b32 {l0, h0}, %8;\n\tmov.b32 {l1, h1}, %9;\n\tmov.b32 {l2, h2}, %10;\n\tmov.b32 {l3, h3}, %11;\n\t"
      "fma.rn.f32.bf16 %0, l0, %12, %0;\n\tfma.rn.f32.bf16 %1, h0, %12, %1;\n\t"
      "fma.rn.f32.bf16 %2, l1, %12, %2;\n\tfma.rn.f32.bf16 %3, h1, %12, %3;\n\t"
      "fma.rn.f32.bf16 %4, l2, %12, %4;\n\tfma.rn.f32.bf16 %5, h2, %12, %5;\n\t"
      "fma.rn.f32.bf16 %6, l3, %12, %6;\n\tfma.rn.f32.bf16 %7, h3, %12, %7;\n\t}"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w), "h"(w));
}

// acc[k] += u[k] * v[k] for two 16-byte vectors of 8 bf16 values (exact products, fp32 accumulation)
__device__ __forceinline__ void fma8_bf16_vv(float (&acc)[8], const uint4& u, const uint4& v) {
  asm("{\n\t.reg .b16 l0, h0, l1, h1, l2, h2, l3, h3, m0, n0, m1, n1, m2, n2, m3, n3;\n\t"
      "mov.b32 {l0, h0}, %8;\n\tmov.b32 {l1, h1}, %9;\n\tmov.b32 {l2, h2}, %10;\n\tmov.b32 {l3, h3}, %11;\n\t"
      "mov.b32 {m0, n0}, %12;\n\tmov.b32 {m1, n1}, %13;\n\tmov.b32 {m2, n2}, %14;\n\tmov.b32 {m3, n3}, %15;\n\t"
      "fma.rn.f32.bf16 %0, l0, m0, %0;\n\tfma.rn.f32.bf16 %1, h0, n0, %1;\n\t"
      "fma.rn.f32.bf16 %2, l1, m1, %2;\n\tfma.rn.f32.bf16 %3, h1, n1, %3;\n\t"
      "fma.rn.f32.bf16 %4, l2, m2, %4;\n\tfma.rn.f32.bf16 %5, h2, n2, %5;\n\t"
      "fma.rn.f32.bf16 %6, l3, m3, %6;\n\tfma.rn.f32.bf16 %7, h3, n3, %7;\n\t}"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// packed bf16x2 helpers on raw 32-bit words (HMNMX2 / HSET2 / HFMA2 .BF16_V2: one instruction per pair)
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bf2_eq_mask(uint32_t a, uint32_t b) {   // 0xffff per equal half
  return __heq2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
}
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) {       // correctly rounded bf16 sums
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// 16-byte shared-memory load through an explicit shared-window address (pointers derived from an aligned dynamic
// shared-memory base by integer arithmetic lose their address space and would compile to generic LD)
__device__ __forceinline__ uint4 lds16(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr)
               : "memory");
  return v;
}

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& u) { *reinterpret_cast<uint4*>(p) = u; }

}  // namespace b2
