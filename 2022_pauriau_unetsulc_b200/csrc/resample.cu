// MaxPool3d(2) backward, trilinear upsample (+ skip-concat: the result is written into the channel window
// [coff, coff+C) of the pre-allocated concat buffer whose first channels already hold the skip tensor) and its
// backward.  NDHWC bf16; one 16-byte channel octet per thread; HBM/L2-bandwidth bound.
#include "common.h"
#include "ptx.cuh"
#include "vec.cuh"
#include <stdlib.h>

namespace b2 {

static inline int ew_blocks(long long total) {
  long long nb = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

// GroupNorm-backward statistics of a tensor this file PRODUCES (it is the gradient at a GroupNorm output): per channel
// sum(g) and sum(g * r) over the stored (bf16-rounded) values, r = the layer's saved relu(conv).  A thread keeps the
// sums of its fixed channel octet in registers; the block combines them in a fixed order and adds ONE exact,
// order-independent contribution per channel to the [C][4] accumulators (common.h) — the separate statistics pass
// over (dy, r) disappears.  Requires blockDim.x == 256 and 256 % (C/8) == 0.
__device__ __forceinline__ void block_bstats_commit(const float (&s)[8], const float (&q)[8], int C, int oct, int vloc,
                                                    long long* __restrict__ acc) {
  __shared__ float red[256 * 16];   // [256 / C8][C][2]
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[((size_t)vloc * C + oct * 8 + k) * 2 + 0] = s[k];
    red[((size_t)vloc * C + oct * 8 + k) * 2 + 1] = q[k];
  }
  __syncthreads();
  const int rows = 256 / (C >> 3);
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < rows; ++j) { a += red[((size_t)j * C + c) * 2]; b += red[((size_t)j * C + c) * 2 + 1]; }
    stat_atomic_add(acc + 4 * c, a);
    stat_atomic_add(acc + 4 * c + 2, b);
  }
}
__device__ __forceinline__ void bstats_accumulate(float (&s)[8], float (&q)[8], const uint4& stored, const uint4& rr) {
  fma8_bf16(s, stored, (unsigned short)0x3f80);   // s += g * 1
  fma8_bf16_vv(q, stored, rr);                    // q += g * r   (FHFMA.BF16: no unpacking)
}

// out[v, c] = dskip[v, c] + (v is the arg-max of its 2x2x2 cell ? dpool[cell, c] : 0)
// arg-max is recomputed from the stored forward tensor y; first maximum in (d, h, w) scan order wins, as in
// PyTorch's max_pool3d_with_indices.
__global__ void __launch_bounds__(256, 4)
pool_bwd_add_kernel(const __nv_bfloat16* __restrict__ y, int ldy, int y_coff, const __nv_bfloat16* __restrict__ dskip,
                    int ldd, int d_coff, const __nv_bfloat16* __restrict__ dpool, __nv_bfloat16* __restrict__ out,
                    int N, int D, int H, int W, int C, const __nv_bfloat16* __restrict__ stat_r,
                    long long* __restrict__ stat_acc) {
  pdl_prologue();
  float st_s[8], st_q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { st_s[k] = 0.f; st_q[k] = 0.f; }
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
    // packed bf16x2 arithmetic throughout (HMNMX2 / HSET2 / HFMA2): the channel-wise maximum of the 8 voxels, then the
    // FIRST voxel in (d, h, w) scan order that equals it takes the pooled gradient (PyTorch's max_pool3d_with_indices:
    // a later voxel replaces the running maximum only if strictly greater), added to the skip gradient with one
    // correctly rounded bf16 add.  No unpacking to fp32: ~2x fewer instructions than the fp32 form.
    uint32_t mx[4] = {0u, 0u, 0u, 0u}, gp[4] = {0u, 0u, 0u, 0u}, found[4] = {0u, 0u, 0u, 0u};
    if (pooled) {   // pass 1: channel-wise maximum of the cell (the 8 rows are read again in pass 2: L1 hits, and 32
                    // registers less than keeping them: 4 instead of 2 resident blocks per SM)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int d = 2 * cd + (j >> 2), h = 2 * ch + ((j >> 1) & 1), w = 2 * cw + (j & 1);
        const long long v = (((long long)n * D + d) * H + h) * W + w;
        const uint4 yv = ldg16(y + v * ldy + y_coff + oct * 8);
        if (j == 0) { mx[0] = yv.x; mx[1] = yv.y; mx[2] = yv.z; mx[3] = yv.w; }
        else {
          mx[0] = bf2_max(mx[0], yv.x); mx[1] = bf2_max(mx[1], yv.y);
          mx[2] = bf2_max(mx[2], yv.z); mx[3] = bf2_max(mx[3], yv.w);
        }
      }
      const long long pv = (((long long)n * Dp + cd) * Hp + ch) * Wp + cw;
      const uint4 g4 = ldg16(dpool + pv * C + oct * 8);
      gp[0] = g4.x; gp[1] = g4.y; gp[2] = g4.z; gp[3] = g4.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = 2 * cd + (j >> 2), h = 2 * ch + ((j >> 1) & 1), w = 2 * cw + (j & 1);
      if (d < D && h < H && w < W) {
        const long long v = (((long long)n * D + d) * H + h) * W + w;
        uint4 g = make_uint4(0u, 0u, 0u, 0u);
        if (dskip) g = ldg16(dskip + v * ldd + d_coff + oct * 8);
        if (pooled) {
          const uint4 yj = ldg16(y + v * ldy + y_coff + oct * 8);
          const uint32_t yw[4] = {yj.x, yj.y, yj.z, yj.w};
          uint32_t add[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t eq = bf2_eq_mask(yw[k], mx[k]) & ~found[k];
            found[k] |= eq;
            add[k] = gp[k] & eq;
          }
          g.x = bf2_add(g.x, add[0]); g.y = bf2_add(g.y, add[1]);
          g.z = bf2_add(g.z, add[2]); g.w = bf2_add(g.w, add[3]);
        }
        stg16(out + v * C + oct * 8, g);
        if (stat_acc != nullptr) bstats_accumulate(st_s, st_q, g, ldg16(stat_r + v * C + oct * 8));
      }
    }
  }
  if (stat_acc != nullptr)   // grid stride is a multiple of C8: the octet of a thread never changes
    block_bstats_commit(st_s, st_q, C, (int)(threadIdx.x % C8), (int)(threadIdx.x / C8), stat_acc);
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

static constexpr int kMaxUpW = 512;   // widest output row the per-row tables support
static constexpr int kMaxTouch = 8;   // fine samples that can touch one coarse sample along a dimension (ratio <= 3.5)

// one block per output row (n, d, h): the D/H interpolation is uniform per block, the W interpolation comes from a
// shared-memory table built once per block; threads sweep (w, channel octet) with 32-bit index math.
__global__ void __launch_bounds__(256)
upsample_cat_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int Di, int Hi, int Wi, int C,
                        __nv_bfloat16* __restrict__ cat, int ldc, int coff, int Do, int Ho, int Wo) {
  pdl_prologue();
  __shared__ int s_i0[kMaxUpW], s_i1[kMaxUpW];
  __shared__ float s_l1[kMaxUpW];
  const int C8 = C >> 3;
  const int row = blockIdx.x;
  const int h = row % Ho, d = (row / Ho) % Do, n = row / (Ho * Do);
  const float sd = (float)Di / (float)Do, shh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  for (int w = threadIdx.x; w < Wo; w += blockDim.x) {
    int i0, i1; float l0, l1;
    src_index(w, sw, Wi, i0, i1, l0, l1);
    s_i0[w] = i0 * C; s_i1[w] = i1 * C; s_l1[w] = l1;
  }
  int d0, d1, h0, h1; float ld0, ld1, lh0, lh1;
  src_index(d, sd, Di, d0, d1, ld0, ld1);
  src_index(h, shh, Hi, h0, h1, lh0, lh1);
  __syncthreads();
  const __nv_bfloat16* xb = x + (size_t)n * Di * Hi * Wi * C;
  const __nv_bfloat16* r00 = xb + ((size_t)d0 * Hi + h0) * Wi * C;
  const __nv_bfloat16* r01 = xb + ((size_t)d0 * Hi + h1) * Wi * C;
  const __nv_bfloat16* r10 = xb + ((size_t)d1 * Hi + h0) * Wi * C;
  const __nv_bfloat16* r11 = xb + ((size_t)d1 * Hi + h1) * Wi * C;
  __nv_bfloat16* out = cat + (size_t)row * Wo * ldc + coff;
  const int total = Wo * C8;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int w = e / C8, oc = (e - w * C8) * 8;
    const int o0 = s_i0[w] + oc, o1 = s_i1[w] + oc;
    const float lw1 = s_l1[w], lw0 = 1.f - lw1;
    const f8 a000 = unpack8(ldg16(r00 + o0)), a001 = unpack8(ldg16(r00 + o1));
    const f8 a010 = unpack8(ldg16(r01 + o0)), a011 = unpack8(ldg16(r01 + o1));
    const f8 a100 = unpack8(ldg16(r10 + o0)), a101 = unpack8(ldg16(r10 + o1));
    const f8 a110 = unpack8(ldg16(r11 + o0)), a111 = unpack8(ldg16(r11 + o1));
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // same association order as ATen's upsample_trilinear3d kernel
      o.v[k] = ld0 * (lh0 * (lw0 * a000.v[k] + lw1 * a001.v[k]) + lh1 * (lw0 * a010.v[k] + lw1 * a011.v[k])) +
               ld1 * (lh0 * (lw0 * a100.v[k] + lw1 * a101.v[k]) + lh1 * (lw0 * a110.v[k] + lw1 * a111.v[k]));
    }
    stg16(out + (size_t)w * ldc + oc, pack8(o));
  }
}

// Exact 2x upsampling (every level of a volume whose sides are multiples of 8): src = dst/2 - 0.25, so the fine
// samples 2j+1 and 2j+2 both interpolate between the coarse samples j and j+1 with weights (.75,.25) / (.25,.75).
// One thread per CELL (jd, jh, jw) in [-1, Di-1] x [-1, Hi-1] x [-1, Wi-1] and channel octet: 8 coarse loads feed
// 8 fine stores (the generic kernel needs 8 loads per store).  Edge cells clamp the coarse index exactly like
// src_index does (fine 0 = coarse 0 with weights (1,0); the last fine sample = .75*a + .25*a of the last coarse one).
struct Cell2x {
  int c0, c1;      // coarse indices
  int f[2];        // fine indices (2j+1, 2j+2), -1 when outside
  float l0[2], l1[2];
};
__device__ __forceinline__ Cell2x cell2x(int j, int n_in) {
  Cell2x c;
  c.c0 = j < 0 ? 0 : j;
  c.c1 = (j + 1 > n_in - 1) ? n_in - 1 : j + 1;
  c.f[0] = j >= 0 ? 2 * j + 1 : -1;
  c.f[1] = (j + 1 <= n_in - 1) ? 2 * j + 2 : -1;
  c.l0[0] = 0.75f; c.l1[0] = 0.25f;
  c.l0[1] = j < 0 ? 1.f : 0.25f;
  c.l1[1] = j < 0 ? 0.f : 0.75f;
  return c;
}

__global__ void __launch_bounds__(256)
upsample2x_cat_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int Di, int Hi, int Wi, int C,
                          __nv_bfloat16* __restrict__ cat, int ldc, int coff) {
  pdl_prologue();
  const int C8 = C >> 3;
  const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
  const long long total = (long long)N * (Di + 1) * (Hi + 1) * (Wi + 1) * C8;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int oc = (int)(i % C8) * 8;
  long long t = i / C8;
  const int jw = (int)(t % (Wi + 1)) - 1; t /= (Wi + 1);
  const int jh = (int)(t % (Hi + 1)) - 1; t /= (Hi + 1);
  const int jd = (int)(t % (Di + 1)) - 1;
  const int n = (int)(t / (Di + 1));
  const Cell2x cd = cell2x(jd, Di), ch = cell2x(jh, Hi), cw = cell2x(jw, Wi);
  const __nv_bfloat16* xb = x + (size_t)n * Di * Hi * Wi * C + oc;
  uint4 a[2][2][2];
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int hz = 0; hz < 2; ++hz)
#pragma unroll
      for (int wz = 0; wz < 2; ++wz)
        a[dz][hz][wz] = ldg16(xb + (((size_t)(dz ? cd.c1 : cd.c0) * Hi + (hz ? ch.c1 : ch.c0)) * Wi +
                                    (wz ? cw.c1 : cw.c0)) * C);
  uint32_t o[2][2][2][4];   // [fd][fh][fw][channel pair]
#pragma unroll
  for (int cp = 0; cp < 4; ++cp) {
    float lo[2][2][2], hi[2][2][2];
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int hz = 0; hz < 2; ++hz)
#pragma unroll
        for (int wz = 0; wz < 2; ++wz) {
          const uint4& u = a[dz][hz][wz];
          const uint32_t wd = cp == 0 ? u.x : (cp == 1 ? u.y : (cp == 2 ? u.z : u.w));
          lo[dz][hz][wz] = __uint_as_float(wd << 16);
          hi[dz][hz][wz] = __uint_as_float(wd & 0xffff0000u);
        }
#pragma unroll
    for (int fd = 0; fd < 2; ++fd)
#pragma unroll
      for (int fh = 0; fh < 2; ++fh)
#pragma unroll
        for (int fw = 0; fw < 2; ++fw) {
          const float lw0 = cw.l0[fw], lw1 = cw.l1[fw], lh0 = ch.l0[fh], lh1 = ch.l1[fh];
          const float ld0 = cd.l0[fd], ld1 = cd.l1[fd];
          // same association order as ATen's upsample_trilinear3d kernel (and upsample_cat_fwd_kernel)
          const float vl = ld0 * (lh0 * (lw0 * lo[0][0][0] + lw1 * lo[0][0][1]) + lh1 * (lw0 * lo[0][1][0] + lw1 * lo[0][1][1])) +
                           ld1 * (lh0 * (lw0 * lo[1][0][0] + lw1 * lo[1][0][1]) + lh1 * (lw0 * lo[1][1][0] + lw1 * lo[1][1][1]));
          const float vh = ld0 * (lh0 * (lw0 * hi[0][0][0] + lw1 * hi[0][0][1]) + lh1 * (lw0 * hi[0][1][0] + lw1 * hi[0][1][1])) +
                           ld1 * (lh0 * (lw0 * hi[1][0][0] + lw1 * hi[1][0][1]) + lh1 * (lw0 * hi[1][1][0] + lw1 * hi[1][1][1]));
          __nv_bfloat162 pk = __floats2bfloat162_rn(vl, vh);
          o[fd][fh][fw][cp] = *reinterpret_cast<uint32_t*>(&pk);
        }
  }
  __nv_bfloat16* ob = cat + (size_t)n * Do * Ho * Wo * ldc + coff + oc;
#pragma unroll
  for (int fd = 0; fd < 2; ++fd)
#pragma unroll
    for (int fh = 0; fh < 2; ++fh)
#pragma unroll
      for (int fw = 0; fw < 2; ++fw) {
        if (cd.f[fd] < 0 || ch.f[fh] < 0 || cw.f[fw] < 0) continue;
        const size_t v = ((size_t)cd.f[fd] * Ho + ch.f[fh]) * Wo + cw.f[fw];
        stg16(ob + v * ldc, make_uint4(o[fd][fh][fw][0], o[fd][fh][fw][1], o[fd][fh][fw][2], o[fd][fh][fw][3]));
      }
}

// fine samples whose interpolation touches coarse index i, with their weights (exact: same src_index as forward)
__device__ __forceinline__ int touch_list(int i, float scale, int in_size, int out_size, int* idx, float* wt) {
  const float a = ((float)i - 0.5f) / scale - 0.5f;
  const float b = ((float)i + 1.5f) / scale - 0.5f;
  int lo = (int)floorf(a) - 1, hi = (int)ceilf(b) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
  int cnt = 0;
  for (int o = lo; o <= hi; ++o) {
    int i0, i1; float l0, l1;
    src_index(o, scale, in_size, i0, i1, l0, l1);
    const float w = (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
    if (w != 0.f && cnt < kMaxTouch) { idx[cnt] = o; wt[cnt] = w; ++cnt; }
  }
  return cnt;
}

// adjoint of the trilinear upsample as a deterministic gather: one block per coarse row (n, d, h)
__global__ void __launch_bounds__(256)
upsample_cat_bwd_kernel(const __nv_bfloat16* __restrict__ dcat, int ldc, int coff, int N, int Do, int Ho, int Wo,
                        __nv_bfloat16* __restrict__ dx, int Di, int Hi, int Wi, int C) {
  pdl_prologue();
  __shared__ int s_cnt[kMaxUpW];
  __shared__ int s_idx[kMaxUpW][kMaxTouch];
  __shared__ float s_wt[kMaxUpW][kMaxTouch];
  const int C8 = C >> 3;
  const int row = blockIdx.x;
  const int h = row % Hi, d = (row / Hi) % Di, n = row / (Hi * Di);
  const float sd = (float)Di / (float)Do, shh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  for (int w = threadIdx.x; w < Wi; w += blockDim.x) {
    int idx[kMaxTouch]; float wt[kMaxTouch];
    const int c = touch_list(w, sw, Wi, Wo, idx, wt);
    s_cnt[w] = c;
    for (int k = 0; k < c; ++k) { s_idx[w][k] = idx[k] * ldc; s_wt[w][k] = wt[k]; }
  }
  int didx[kMaxTouch], hidx[kMaxTouch];
  float dwt[kMaxTouch], hwt[kMaxTouch];
  const int nd = touch_list(d, sd, Di, Do, didx, dwt);
  const int nh = touch_list(h, shh, Hi, Ho, hidx, hwt);
  __syncthreads();
  const __nv_bfloat16* gb = dcat + (size_t)n * Do * Ho * Wo * ldc + coff;
  __nv_bfloat16* out = dx + (size_t)row * Wi * C;
  const int total = Wi * C8;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int w = e / C8, oc = (e - w * C8) * 8;
    const int nw = s_cnt[w];
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int a = 0; a < nd; ++a) {
      for (int b = 0; b < nh; ++b) {
        const __nv_bfloat16* rp = gb + ((size_t)didx[a] * Ho + hidx[b]) * Wo * ldc + oc;
        const float wdh = dwt[a] * hwt[b];
        for (int c = 0; c < nw; ++c) {
          const float wgt = wdh * s_wt[w][c];
          const f8 g = unpack8(ldg16(rp + s_idx[w][c]));
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, g.v[k], acc[k]);
        }
      }
    }
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
    stg16(out + (size_t)w * C + oc, pack8(o));
  }
}

// Separable adjoint, pass 1 (W): t[n, od, oh, wc, c] = sum_ow ww(wc, ow) * dcat[n, od, oh, ow, c]   (bf16 out)
// one block per fine row (n, od, oh); instruction-light streaming pass over the big tensor.
__global__ void __launch_bounds__(256)
upsample_bwd_w_kernel(const __nv_bfloat16* __restrict__ dcat, int ldc, int coff, int Wo, __nv_bfloat16* __restrict__ t,
                      int Wi, int C) {
  pdl_prologue();
  __shared__ int s_cnt[kMaxUpW];
  __shared__ int s_idx[kMaxUpW][kMaxTouch];
  __shared__ float s_wt[kMaxUpW][kMaxTouch];
  const int C8 = C >> 3;
  const float sw = (float)Wi / (float)Wo;
  for (int w = threadIdx.x; w < Wi; w += blockDim.x) {
    int idx[kMaxTouch]; float wt[kMaxTouch];
    const int c = touch_list(w, sw, Wi, Wo, idx, wt);
    s_cnt[w] = c;
    for (int k = 0; k < c; ++k) { s_idx[w][k] = idx[k] * ldc; s_wt[w][k] = wt[k]; }
  }
  __syncthreads();
  const __nv_bfloat16* rp = dcat + (size_t)blockIdx.x * Wo * ldc + coff;
  __nv_bfloat16* out = t + (size_t)blockIdx.x * Wi * C;
  const int total = Wi * C8;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int w = e / C8, oc = (e - w * C8) * 8;
    const int nw = s_cnt[w];
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int c = 0; c < nw; ++c) {
      const float wgt = s_wt[w][c];
      const f8 g = unpack8(ldg16(rp + s_idx[w][c] + oc));
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, g.v[k], acc[k]);
    }
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
    stg16(out + (size_t)w * C + oc, pack8(o));
  }
}

// pass 2 (H, D): dx[n, d, h, w, c] = sum_{od, oh} wd * wh * t[n, od, oh, w, c]; one block per coarse row (n, d, h)
__global__ void __launch_bounds__(256)
upsample_bwd_hd_kernel(const __nv_bfloat16* __restrict__ t, int Do, int Ho, __nv_bfloat16* __restrict__ dx, int Di,
                       int Hi, int Wi, int C, const __nv_bfloat16* __restrict__ stat_r,
                       long long* __restrict__ stat_acc) {
  pdl_prologue();
  float st_s[8], st_q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { st_s[k] = 0.f; st_q[k] = 0.f; }
  const int row = blockIdx.x;
  const int h = row % Hi, d = (row / Hi) % Di, n = row / (Hi * Di);
  const float sd = (float)Di / (float)Do, shh = (float)Hi / (float)Ho;
  int didx[kMaxTouch], hidx[kMaxTouch];
  float dwt[kMaxTouch], hwt[kMaxTouch];
  const int nd = touch_list(d, sd, Di, Do, didx, dwt);
  const int nh = touch_list(h, shh, Hi, Ho, hidx, hwt);
  const int rowlen = Wi * C;           // elements per (od, oh) row of t
  const __nv_bfloat16* tb = t + (size_t)n * Do * Ho * rowlen;
  __nv_bfloat16* out = dx + (size_t)row * rowlen;
  for (int e = threadIdx.x * 8; e < rowlen; e += blockDim.x * 8) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int a = 0; a < nd; ++a)
      for (int b = 0; b < nh; ++b) {
        const float wgt = dwt[a] * hwt[b];
        const f8 g = unpack8(ldg16(tb + ((size_t)didx[a] * Ho + hidx[b]) * rowlen + e));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, g.v[k], acc[k]);
      }
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
    const uint4 pk = pack8(o);
    stg16(out + e, pk);
    if (stat_acc != nullptr) bstats_accumulate(st_s, st_q, pk, ldg16(stat_r + (size_t)row * rowlen + e));
  }
  if (stat_acc != nullptr) {   // 2048 % C == 0: a thread's channel octet is the same for every element it visits
    const int C8 = C >> 3;
    block_bstats_commit(st_s, st_q, C, (int)(threadIdx.x % C8), (int)(threadIdx.x / C8), stat_acc);
  }
}

// ---- exact-2x trilinear adjoint in ONE pass -------------------------------------------------------------------------
// Every level of a volume whose sides are multiples of 8 is an exact 2x upsampling (src = dst/2 - 0.25): the adjoint is a
// fixed 4-tap stencil per axis,
//     dx[i] = .25 g[2i-1] + .75 g[2i] + .75 g[2i+1] + .25 g[2i+2],   with weight 1 for g[0] and g[2n-1]
// (fine sample 0 copies coarse 0; the last fine sample puts both of its weights on the clamped last coarse sample;
// g[-1] and g[2n] do not exist: TMA zero-fills them).  The two-pass separable form moved 1.9x the algorithmic bytes
// through a bf16 intermediate (round 1: 1.33 TB/s); here a CTA owns a coarse tile of kUbTH x kUbTW voxels x 64 channels,
// marches along D over a segment of coarse planes and streams the fine planes of its (2 kUbTH + 2) x (2 kUbTW + 2) x 64
// window through a TMA ring: every fine value crosses L2 -> SM once per tile (halo 1.2-1.4x, from L2), DRAM sees
// the fine tensor once, the coarse tensor is written once, and the GroupNorm-backward statistics of the result are
// accumulated on the way out.  Thread = (coarse h, coarse w, channel octet): 16 x LDS.128 + 128 FMA per fine plane
// reduce H and W; the D taps are a two-register rolling combination (a fine plane pair (2j-1, 2j) finishes coarse
// plane j-1 and opens j).
static constexpr int kUbTH = 4, kUbTW = 8, kUbC = 64, kUbStages = 4;
static constexpr int kUbThreads = kUbTH * kUbTW * (kUbC / 8);                                  // 256
static constexpr int kUbPlaneBytes = (2 * kUbTH + 2) * (2 * kUbTW + 2) * kUbC * 2;            // 23 040

__device__ __forceinline__ float ub_weight(int f, int k, int n_fine) {
  // tap k (0..3) of a coarse sample reads fine index f = 2i - 1 + k
  const float base = (k == 0 || k == 3) ? 0.25f : 0.75f;
  return (f == 0 || f == n_fine - 1) ? 1.f : base;
}

__global__ void __launch_bounds__(kUbThreads, 2)
upsample2x_bwd_kernel(const __grid_constant__ CUtensorMap tmap, __nv_bfloat16* __restrict__ dx, int Di, int Hi, int Wi,
                      int C, int tiles_h, int tiles_w, int n_slices, int Ds, int n_dseg, int n_items,
                      const __nv_bfloat16* __restrict__ stat_r, long long* __restrict__ stat_acc) {
  pdl_prologue();
  extern __shared__ uint8_t ub_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ub_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full[kUbStages];
  __shared__ float red[kUbTH * kUbTW][kUbC][2];
  const int tid = threadIdx.x;
  const int oct = tid & 7, lw = (tid >> 3) % kUbTW, lh = tid / (8 * kUbTW);
  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < kUbStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  float st_s[8], st_q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { st_s[k] = 0.f; st_q[k] = 0.f; }
  const int Ho = 2 * Hi, Wo = 2 * Wi, Do = 2 * Di;
  uint32_t gp = 0;   // planes consumed so far by this CTA (ring position / parity), identical in every thread
  int slice = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int t = item;
    slice = t % n_slices; t /= n_slices;     // gridDim.x is a multiple of n_slices: a CTA keeps its channel slice
    const int tw = t % tiles_w; t /= tiles_w;
    const int th = t % tiles_h; t /= tiles_h;
    const int dseg = t % n_dseg;
    const int n = t / n_dseg;
    const int h0 = th * kUbTH, w0 = tw * kUbTW, d0 = dseg * Ds;
    const int d1 = min(d0 + Ds, Di);
    const int n_planes = 2 * (d1 - d0 + 1);            // fine planes 2 d0 - 1 .. 2 d1
    const int h = h0 + lh, w = w0 + lw;
    unsigned short whw[4][4];   // bf16 bits of wh[a] * ww[b]: products of {.25, .75, 1} are exact in bf16
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        whw[a][b] = bf16_bits_exact(ub_weight(2 * h - 1 + a, a, Ho) * ub_weight(2 * w - 1 + b, b, Wo));
    const bool valid_hw = (h < Hi) && (w < Wi);
    float prev[8], cur[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { prev[k] = 0.f; cur[k] = 0.f; }
    // prologue: planes 0 .. S-2 of this item (the previous item's planes are all consumed: see the barrier below)
    __syncthreads();
    if (tid == 0) {
      for (int p = 0; p < kUbStages - 1 && p < n_planes; ++p) {
        const int s = (int)((gp + p) % kUbStages);
        mbar_arrive_expect_tx(&full[s], (uint32_t)kUbPlaneBytes);
        tma_load_5d(ring + (size_t)s * kUbPlaneBytes, &tmap, &full[s], slice * kUbC, 2 * w0 - 1, 2 * h0 - 1,
                    2 * d0 - 1 + p, n);
      }
    }
    for (int p = 0; p < n_planes; ++p, ++gp) {
      __syncthreads();                                   // plane p-1 is consumed by everybody: its stage is free
      if (tid == 0 && p + kUbStages - 1 < n_planes) {
        const int s = (int)((gp + kUbStages - 1) % kUbStages);
        mbar_arrive_expect_tx(&full[s], (uint32_t)kUbPlaneBytes);
        tma_load_5d(ring + (size_t)s * kUbPlaneBytes, &tmap, &full[s], slice * kUbC, 2 * w0 - 1, 2 * h0 - 1,
                    2 * d0 - 1 + p + kUbStages - 1, n);
      }
      const int s = (int)(gp % kUbStages);
      mbar_wait(&full[s], (gp / kUbStages) & 1u);
      const uint32_t pl = smem_u32(ring) + (uint32_t)s * kUbPlaneBytes +
                          (uint32_t)(((2 * lh) * (2 * kUbTW + 2) + 2 * lw) * kUbC + oct * 8) * 2u;
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          fma8_bf16(acc, lds16(pl + (uint32_t)((a * (2 * kUbTW + 2) + b) * kUbC) * 2u), whw[a][b]);
      // D taps: fine plane f = 2 d0 - 1 + p.  Even p: f = 2j - 1 (tap 0 of coarse j, tap 2 of coarse j - 1);
      // odd p: f = 2j (tap 1 of j, tap 3 of j - 1, which it completes), j = d0 + p / 2.
      const int f = 2 * d0 - 1 + p;
      const int j = d0 + (p >> 1);
      if ((p & 1) == 0) {
        const float w_open = ub_weight(f, 0, Do), w_close = ub_weight(f, 2, Do);
#pragma unroll
        for (int k = 0; k < 8; ++k) { cur[k] = w_open * acc[k]; prev[k] = fmaf(w_close, acc[k], prev[k]); }
      } else {
        const float w_open = ub_weight(f, 1, Do), w_close = ub_weight(f, 3, Do);
        f8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) { o.v[k] = fmaf(w_close, acc[k], prev[k]); prev[k] = fmaf(w_open, acc[k], cur[k]); }
        if (j - 1 >= d0 && valid_hw) {                    // coarse plane j - 1 is complete
          const size_t v = (((size_t)n * Di + (j - 1)) * Hi + h) * Wi + w;
          const uint4 pk = pack8(o);
          stg16(dx + v * C + slice * kUbC + oct * 8, pk);
          if (stat_acc != nullptr) bstats_accumulate(st_s, st_q, pk, ldg16(stat_r + v * C + slice * kUbC + oct * 8));
        }
      }
    }
  }
  if (stat_acc != nullptr) {
    __syncthreads();
    const int vloc = tid >> 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[vloc][oct * 8 + k][0] = st_s[k];
      red[vloc][oct * 8 + k][1] = st_q[k];
    }
    __syncthreads();
    if (tid < kUbC) {
      float a = 0.f, b = 0.f;
      for (int r = 0; r < kUbTH * kUbTW; ++r) { a += red[r][tid][0]; b += red[r][tid][1]; }
      stat_atomic_add(stat_acc + 4 * (slice * kUbC + tid), a);
      stat_atomic_add(stat_acc + 4 * (slice * kUbC + tid) + 2, b);
    }
  }
}

int make_act_tmap_plain(CUtensorMap* map, const void* base, int N, int D, int H, int W, int C, int ld, int coff,
                        int box_c, int bw, int bh, int bd);

}  // namespace b2

using namespace b2;

static int maxpool3d_bwd_add_impl(const void* y, int ldy, int y_coff, const void* dskip, int ldd, int d_coff,
                                  const void* dpool, void* out, int N, int D, int H, int W, int C, const void* stat_r,
                                  long long* stat_acc, cudaStream_t stream) {
  B2_REQUIRE(y && dpool && out, "b2_maxpool3d_bwd_add: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldy % 8 == 0 && y_coff % 8 == 0 && ldd % 8 == 0 && d_coff % 8 == 0,
             "b2_maxpool3d_bwd_add: channel counts must be multiples of 8");
  const long long total = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  int blocks = ew_blocks(total);
  if (stat_acc) {
    B2_REQUIRE(N == 1 && 256 % (C / 8) == 0, "b2_maxpool3d_bwd_add_bstats: needs batch 1 and C/8 dividing 256 (C=%d)", C);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;   // one resident wave: fewer atomic contributions
  }
  B2_LAUNCH(pool_bwd_add_kernel, blocks, 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(y), ldy, y_coff,
            reinterpret_cast<const __nv_bfloat16*>(dskip), ldd, d_coff, reinterpret_cast<const __nv_bfloat16*>(dpool),
            reinterpret_cast<__nv_bfloat16*>(out), N, D, H, W, C, reinterpret_cast<const __nv_bfloat16*>(stat_r),
            stat_acc);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_maxpool3d_bwd_add(const void* y, int ldy, int y_coff, const void* dskip, int ldd, int d_coff,
                                    const void* dpool, void* out, int N, int D, int H, int W, int C,
                                    cudaStream_t stream) {
  return maxpool3d_bwd_add_impl(y, ldy, y_coff, dskip, ldd, d_coff, dpool, out, N, D, H, W, C, nullptr, nullptr,
                                stream);
}

// Same (batch 1); `out` is the gradient at a GroupNorm output: also accumulates that layer's GroupNorm-backward
// statistics (sum out, sum out*r; r = the layer's saved relu(conv), dense bf16 [V][C]) into stat_acc int64 [C][4].
extern "C" int b2_maxpool3d_bwd_add_bstats(const void* y, int ldy, int y_coff, const void* dskip, int ldd, int d_coff,
                                           const void* dpool, void* out, int N, int D, int H, int W, int C,
                                           const void* r, long long* stat_acc, cudaStream_t stream) {
  B2_REQUIRE(r && stat_acc, "b2_maxpool3d_bwd_add_bstats: null pointer");
  return maxpool3d_bwd_add_impl(y, ldy, y_coff, dskip, ldd, d_coff, dpool, out, N, D, H, W, C, r, stat_acc, stream);
}

extern "C" int b2_upcat_fwd(const void* x, int N, int Di, int Hi, int Wi, int C, void* cat, int ldc, int coff, int Do,
                            int Ho, int Wo, cudaStream_t stream) {
  B2_REQUIRE(x && cat, "b2_upcat_fwd: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldc % 8 == 0 && coff % 8 == 0, "b2_upcat_fwd: channel counts must be multiples of 8");
  B2_REQUIRE(Wo <= kMaxUpW && Wi <= kMaxUpW, "b2_upcat_fwd: row width %d > %d", Wo, kMaxUpW);
  B2_REQUIRE((long long)Wi * C < (1LL << 31) && (long long)Wo * ldc < (1LL << 31), "b2_upcat_fwd: row too large");
  if (Do == 2 * Di && Ho == 2 * Hi && Wo == 2 * Wi) {
    const long long total = (long long)N * (Di + 1) * (Hi + 1) * (Wi + 1) * (C / 8);
    B2_LAUNCH(upsample2x_cat_fwd_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, 
        reinterpret_cast<const __nv_bfloat16*>(x), N, Di, Hi, Wi, C, reinterpret_cast<__nv_bfloat16*>(cat), ldc, coff);
    B2_CHECK_CUDA(cudaGetLastError());
    return B2_OK;
  }
  B2_LAUNCH(upsample_cat_fwd_kernel, (unsigned)(N * Do * Ho), 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), N, Di, Hi,
                                                                Wi, C, reinterpret_cast<__nv_bfloat16*>(cat), ldc, coff,
                                                                Do, Ho, Wo);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" long long b2_upcat_bwd_workspace_bytes(int N, int Do, int Ho, int Wi, int C) {
  return (long long)N * Do * Ho * Wi * C * 2;
}

// separable variant: W pass into `workspace` (bf16 [N][Do][Ho][Wi][C]), then the H/D pass
static int upcat_bwd_separable_impl(const void* dcat, int ldc, int coff, int N, int Do, int Ho, int Wo, void* dx,
                                    int Di, int Hi, int Wi, int C, void* workspace, long long workspace_bytes,
                                    const void* stat_r, long long* stat_acc, cudaStream_t stream) {
  B2_REQUIRE(dcat && dx && workspace, "b2_upcat_bwd_separable: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldc % 8 == 0 && coff % 8 == 0, "b2_upcat_bwd_separable: channel counts must be multiples of 8");
  B2_REQUIRE(Wo <= kMaxUpW && Wi <= kMaxUpW, "b2_upcat_bwd_separable: row width %d > %d", Wo, kMaxUpW);
  B2_REQUIRE(2 * Do <= 7 * Di && 2 * Ho <= 7 * Hi && 2 * Wo <= 7 * Wi, "b2_upcat_bwd_separable: ratio > 3.5 unsupported");
  B2_REQUIRE((long long)Wo * ldc < (1LL << 31) && (long long)Wi * C < (1LL << 31), "b2_upcat_bwd_separable: row too large");
  B2_REQUIRE(workspace_bytes >= b2_upcat_bwd_workspace_bytes(N, Do, Ho, Wi, C), "b2_upcat_bwd_separable: workspace too small");
  if (stat_acc)
    B2_REQUIRE(N == 1 && 2048 % C == 0 && C >= 8, "b2_upcat_bwd_separable_bstats: needs batch 1 and C dividing 2048 (C=%d)", C);
  static const bool no_onepass = getenv("B2_NO_UP1PASS") != nullptr;
  if (!no_onepass && Do == 2 * Di && Ho == 2 * Hi && Wo == 2 * Wi && C % kUbC == 0) {
    // exact 2x: single-pass stencil kernel (TMA-staged fine planes), see upsample2x_bwd_kernel
    CUtensorMap tm;
    int rc = make_act_tmap_plain(&tm, dcat, N, Do, Ho, Wo, C, ldc, coff, kUbC, 2 * kUbTW + 2, 2 * kUbTH + 2, 1);
    if (rc) return rc;
    const int tiles_h = ceil_div(Hi, kUbTH), tiles_w = ceil_div(Wi, kUbTW), n_slices = C / kUbC;
    const int slots = 2 * num_sms() / n_slices * n_slices;      // resident CTAs, a multiple of the slice count
    // D segment length: trade the 2-plane halo per segment against filling the last wave
    int best_ds = Di;
    double best = -1.0;
    for (int ds = 1; ds <= Di; ++ds) {
      const long long items = (long long)N * ceil_div(Di, ds) * tiles_h * tiles_w * n_slices;
      const double waves = (double)items / slots;
      const double eff = (waves / (double)((items + slots - 1) / slots)) * (2.0 * ds / (2.0 * ds + 2.0));
      if (eff > best + 1e-9) { best = eff; best_ds = ds; }
    }
    const int n_dseg = ceil_div(Di, best_ds);
    const long long items = (long long)N * n_dseg * tiles_h * tiles_w * n_slices;
    B2_REQUIRE(items < (1LL << 31), "b2_upcat_bwd: too many tiles");
    const int grid = (int)(items < slots ? items : slots);
    const size_t smem = (size_t)kUbStages * kUbPlaneBytes + 128;
    B2_CHECK_CUDA(cudaFuncSetAttribute(upsample2x_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_LAUNCH(upsample2x_bwd_kernel, grid, kUbThreads, smem, stream, tm, reinterpret_cast<__nv_bfloat16*>(dx), Di, Hi,
              Wi, C, tiles_h, tiles_w, n_slices, best_ds, n_dseg, (int)items,
              reinterpret_cast<const __nv_bfloat16*>(stat_r), stat_acc);
    B2_CHECK_CUDA(cudaGetLastError());
    return B2_OK;
  }
  __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(workspace);
  B2_LAUNCH(upsample_bwd_w_kernel, (unsigned)(N * Do * Ho), 256, 0, stream,
            reinterpret_cast<const __nv_bfloat16*>(dcat), ldc, coff, Wo, t, Wi, C);
  B2_CHECK_CUDA(cudaGetLastError());
  B2_LAUNCH(upsample_bwd_hd_kernel, (unsigned)(N * Di * Hi), 256, 0, stream, t, Do, Ho,
            reinterpret_cast<__nv_bfloat16*>(dx), Di, Hi, Wi, C, reinterpret_cast<const __nv_bfloat16*>(stat_r),
            stat_acc);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_upcat_bwd_separable(const void* dcat, int ldc, int coff, int N, int Do, int Ho, int Wo, void* dx,
                                      int Di, int Hi, int Wi, int C, void* workspace, long long workspace_bytes,
                                      cudaStream_t stream) {
  return upcat_bwd_separable_impl(dcat, ldc, coff, N, Do, Ho, Wo, dx, Di, Hi, Wi, C, workspace, workspace_bytes,
                                  nullptr, nullptr, stream);
}

// Same (batch 1); dx is the gradient at a GroupNorm output: also accumulates that layer's GroupNorm-backward
// statistics (sum dx, sum dx*r; r dense bf16 [Di*Hi*Wi][C]) into stat_acc int64 [C][4].
extern "C" int b2_upcat_bwd_separable_bstats(const void* dcat, int ldc, int coff, int N, int Do, int Ho, int Wo,
                                             void* dx, int Di, int Hi, int Wi, int C, void* workspace,
                                             long long workspace_bytes, const void* r, long long* stat_acc,
                                             cudaStream_t stream) {
  B2_REQUIRE(r && stat_acc, "b2_upcat_bwd_separable_bstats: null pointer");
  return upcat_bwd_separable_impl(dcat, ldc, coff, N, Do, Ho, Wo, dx, Di, Hi, Wi, C, workspace, workspace_bytes, r,
                                  stat_acc, stream);
}

extern "C" int b2_upcat_bwd(const void* dcat, int ldc, int coff, int N, int Do, int Ho, int Wo, void* dx, int Di,
                            int Hi, int Wi, int C, cudaStream_t stream) {
  B2_REQUIRE(dcat && dx, "b2_upcat_bwd: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldc % 8 == 0 && coff % 8 == 0, "b2_upcat_bwd: channel counts must be multiples of 8");
  B2_REQUIRE(Wo <= kMaxUpW && Wi <= kMaxUpW, "b2_upcat_bwd: row width %d > %d", Wo, kMaxUpW);
  B2_REQUIRE(2 * Do <= 7 * Di && 2 * Ho <= 7 * Hi && 2 * Wo <= 7 * Wi, "b2_upcat_bwd: upsampling ratio > 3.5 unsupported");
  B2_REQUIRE((long long)Wo * ldc < (1LL << 31), "b2_upcat_bwd: row too large");
  B2_LAUNCH(upsample_cat_bwd_kernel, (unsigned)(N * Di * Hi), 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(dcat), ldc, coff,
                                                                N, Do, Ho, Wo, reinterpret_cast<__nv_bfloat16*>(dx), Di,
                                                                Hi, Wi, C);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
