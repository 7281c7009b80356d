// MaxPool3d(2) backward, trilinear upsample (+ skip-concat: the result is written into the channel window
// [coff, coff+C) of the pre-allocated concat buffer whose first channels already hold the skip tensor) and its
// backward.  NDHWC bf16; one 16-byte channel octet per thread; HBM/L2-bandwidth bound.
#include "common.h"
#include "vec.cuh"

namespace b2 {

static inline int ew_blocks(long long total) {
  long long nb = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

// out[v, c] = dskip[v, c] + (v is the arg-max of its 2x2x2 cell ? dpool[cell, c] : 0)
// arg-max is recomputed from the stored forward tensor y; first maximum in (d, h, w) scan order wins, as in
// PyTorch's max_pool3d_with_indices.
__global__ void __launch_bounds__(256)
pool_bwd_add_kernel(const __nv_bfloat16* __restrict__ y, int ldy, int y_coff, const __nv_bfloat16* __restrict__ dskip,
                    int ldd, int d_coff, const __nv_bfloat16* __restrict__ dpool, __nv_bfloat16* __restrict__ out,
                    int N, int D, int H, int W, int C) {
  const int C8 = C >> 3;
  const int Dc = (D + 1) >> 1, Hc = (H + 1) >> 1, Wc = (W + 1) >> 1;
  const int Dp = D >> 1, Hp = H >> 1, Wp = W >> 1;
  const long long total = (long long)N * Dc * Hc * Wc * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int oct = (int)(i % C8);
    long long t = i / C8;
    const int cw = (int)(t % Wc); t /= Wc;
    const int ch = (int)(t % Hc); t /= Hc;
    const int cd = (int)(t % Dc);
    const int n = (int)(t / Dc);
    const bool pooled = (cd < Dp && ch < Hp && cw < Wp);
    int arg[8];
    f8 gp;
    if (pooled) {
      float mx[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { mx[k] = -INFINITY; arg[k] = 0; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int d = 2 * cd + (j >> 2), h = 2 * ch + ((j >> 1) & 1), w = 2 * cw + (j & 1);
        const long long v = (((long long)n * D + d) * H + h) * W + w;
        const f8 x = unpack8(ldg16(y + v * ldy + y_coff + oct * 8));
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (x.v[k] > mx[k]) { mx[k] = x.v[k]; arg[k] = j; }
      }
      const long long pv = (((long long)n * Dp + cd) * Hp + ch) * Wp + cw;
      gp = unpack8(ldg16(dpool + pv * C + oct * 8));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = 2 * cd + (j >> 2), h = 2 * ch + ((j >> 1) & 1), w = 2 * cw + (j & 1);
      if (d < D && h < H && w < W) {
        const long long v = (((long long)n * D + d) * H + h) * W + w;
        f8 g;
        if (dskip) {
          g = unpack8(ldg16(dskip + v * ldd + d_coff + oct * 8));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) g.v[k] = 0.f;
        }
        if (pooled) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (arg[k] == j) g.v[k] += gp.v[k];
        }
        stg16(out + v * C + oct * 8, pack8(g));
      }
    }
  }
}

// PyTorch upsample_trilinear3d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0, scale = in/out (fp32)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256)
upsample_cat_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int Di, int Hi, int Wi, int C,
                        __nv_bfloat16* __restrict__ cat, int ldc, int coff, int Do, int Ho, int Wo) {
  const int C8 = C >> 3;
  const float sd = (float)Di / (float)Do, shh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  const long long total = (long long)N * Do * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int oct = (int)(i % C8);
    long long t = i / C8;
    const long long vo = t;
    const int w = (int)(t % Wo); t /= Wo;
    const int h = (int)(t % Ho); t /= Ho;
    const int d = (int)(t % Do);
    const int n = (int)(t / Do);
    int d0, d1, h0, h1, w0, w1;
    float ld0, ld1, lh0, lh1, lw0, lw1;
    src_index(d, sd, Di, d0, d1, ld0, ld1);
    src_index(h, shh, Hi, h0, h1, lh0, lh1);
    src_index(w, sw, Wi, w0, w1, lw0, lw1);
    const __nv_bfloat16* xb = x + (size_t)n * Di * Hi * Wi * C + oct * 8;
    auto at = [&](int dd, int hh, int ww) { return unpack8(ldg16(xb + (((size_t)dd * Hi + hh) * Wi + ww) * C)); };
    const f8 a000 = at(d0, h0, w0), a001 = at(d0, h0, w1), a010 = at(d0, h1, w0), a011 = at(d0, h1, w1);
    const f8 a100 = at(d1, h0, w0), a101 = at(d1, h0, w1), a110 = at(d1, h1, w0), a111 = at(d1, h1, w1);
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // same association order as ATen's upsample_trilinear3d kernel
      o.v[k] = ld0 * (lh0 * (lw0 * a000.v[k] + lw1 * a001.v[k]) + lh1 * (lw0 * a010.v[k] + lw1 * a011.v[k])) +
               ld1 * (lh0 * (lw0 * a100.v[k] + lw1 * a101.v[k]) + lh1 * (lw0 * a110.v[k] + lw1 * a111.v[k]));
    }
    stg16(cat + vo * ldc + coff + oct * 8, pack8(o));
  }
}

// range of output indices whose interpolation touches input index i (conservative; exact test done per element)
__device__ __forceinline__ void touch_range(int i, float scale, int out_size, int& lo, int& hi) {
  // src(dst) = scale*(dst+0.5)-0.5 in [i-1, i+1)  =>  dst in ((i-0.5)/scale-0.5, (i+1.5)/scale-0.5)
  float a = ((float)i - 0.5f) / scale - 0.5f;
  float b = ((float)i + 1.5f) / scale - 0.5f;
  lo = (int)floorf(a) - 1;
  hi = (int)ceilf(b) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

__global__ void __launch_bounds__(256)
upsample_cat_bwd_kernel(const __nv_bfloat16* __restrict__ dcat, int ldc, int coff, int N, int Do, int Ho, int Wo,
                        __nv_bfloat16* __restrict__ dx, int Di, int Hi, int Wi, int C) {
  const int C8 = C >> 3;
  const float sd = (float)Di / (float)Do, shh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  const long long total = (long long)N * Di * Hi * Wi * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int oct = (int)(i % C8);
    long long t = i / C8;
    const long long vi = t;
    const int w = (int)(t % Wi); t /= Wi;
    const int h = (int)(t % Hi); t /= Hi;
    const int d = (int)(t % Di);
    const int n = (int)(t / Di);
    int dlo, dhi, hlo, hhi, wlo, whi;
    touch_range(d, sd, Do, dlo, dhi);
    touch_range(h, shh, Ho, hlo, hhi);
    touch_range(w, sw, Wo, wlo, whi);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const __nv_bfloat16* gb = dcat + (size_t)n * Do * Ho * Wo * ldc + coff + oct * 8;
    for (int od = dlo; od <= dhi; ++od) {
      int a0, a1; float la0, la1;
      src_index(od, sd, Di, a0, a1, la0, la1);
      const float wd = (a0 == d ? la0 : 0.f) + (a1 == d ? la1 : 0.f);
      if (wd == 0.f) continue;
      for (int oh = hlo; oh <= hhi; ++oh) {
        int b0, b1; float lb0, lb1;
        src_index(oh, shh, Hi, b0, b1, lb0, lb1);
        const float wh = (b0 == h ? lb0 : 0.f) + (b1 == h ? lb1 : 0.f);
        if (wh == 0.f) continue;
        for (int ow = wlo; ow <= whi; ++ow) {
          int c0, c1; float lc0, lc1;
          src_index(ow, sw, Wi, c0, c1, lc0, lc1);
          const float ww = (c0 == w ? lc0 : 0.f) + (c1 == w ? lc1 : 0.f);
          if (ww == 0.f) continue;
          const float wt = wd * wh * ww;
          const f8 g = unpack8(ldg16(gb + (((size_t)od * Ho + oh) * Wo + ow) * ldc));
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(wt, g.v[k], acc[k]);
        }
      }
    }
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
    stg16(dx + vi * C + oct * 8, pack8(o));
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_maxpool3d_bwd_add(const void* y, int ldy, int y_coff, const void* dskip, int ldd, int d_coff,
                                    const void* dpool, void* out, int N, int D, int H, int W, int C,
                                    cudaStream_t stream) {
  B2_REQUIRE(y && dpool && out, "b2_maxpool3d_bwd_add: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldy % 8 == 0 && y_coff % 8 == 0 && ldd % 8 == 0 && d_coff % 8 == 0,
             "b2_maxpool3d_bwd_add: channel counts must be multiples of 8");
  const long long total = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  pool_bwd_add_kernel<<<ew_blocks(total), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), ldy, y_coff, reinterpret_cast<const __nv_bfloat16*>(dskip), ldd, d_coff,
      reinterpret_cast<const __nv_bfloat16*>(dpool), reinterpret_cast<__nv_bfloat16*>(out), N, D, H, W, C);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_upcat_fwd(const void* x, int N, int Di, int Hi, int Wi, int C, void* cat, int ldc, int coff, int Do,
                            int Ho, int Wo, cudaStream_t stream) {
  B2_REQUIRE(x && cat, "b2_upcat_fwd: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldc % 8 == 0 && coff % 8 == 0, "b2_upcat_fwd: channel counts must be multiples of 8");
  const long long total = (long long)N * Do * Ho * Wo * (C / 8);
  upsample_cat_fwd_kernel<<<ew_blocks(total), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), N, Di, Hi,
                                                                Wi, C, reinterpret_cast<__nv_bfloat16*>(cat), ldc, coff,
                                                                Do, Ho, Wo);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_upcat_bwd(const void* dcat, int ldc, int coff, int N, int Do, int Ho, int Wo, void* dx, int Di,
                            int Hi, int Wi, int C, cudaStream_t stream) {
  B2_REQUIRE(dcat && dx, "b2_upcat_bwd: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldc % 8 == 0 && coff % 8 == 0, "b2_upcat_bwd: channel counts must be multiples of 8");
  const long long total = (long long)N * Di * Hi * Wi * (C / 8);
  upsample_cat_bwd_kernel<<<ew_blocks(total), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dcat), ldc, coff,
                                                                N, Do, Ho, Wo, reinterpret_cast<__nv_bfloat16*>(dx), Di,
                                                                Hi, Wi, C);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
