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
