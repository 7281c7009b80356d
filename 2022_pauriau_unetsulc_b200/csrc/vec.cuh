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

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& u) { *reinterpret_cast<uint4*>(p) = u; }

}  // namespace b2
