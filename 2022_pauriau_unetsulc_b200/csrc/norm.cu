// GroupNorm over post-ReLU activations ('crg' order: conv -> ReLU -> GroupNorm), NDHWC bf16, fp32 statistics.
//
// forward : stats pass  (sum x, sum x^2 per channel, per-block partials, deterministic two-stage reduce)
//           finalize    (per (n, group): mean, rstd; per (n, channel): scale = rstd*gamma, shift = beta - mean*scale)
//           apply pass  y = r*scale + shift  (optionally also writes the 2x2x2 max-pooled tensor in the same read)
// backward: stats pass  (sum dy, sum dy*xhat per channel), finalize (group sums, dgamma, dbeta, per-channel coefs),
//           apply pass  dr = relu'(r) * rstd*(dy*gamma - mean_g(dy*gamma) - xhat*mean_g(dy*gamma*xhat))
// All kernels are HBM-bound: one 16-byte (8-channel) vector per thread per voxel, coalesced along channels.
#include "common.h"
#include "vec.cuh"

namespace b2 {

static constexpr int kStatBlocks = 592;  // 4 x 148 SMs
static constexpr int kStatThreads = 256;

// ------------------------------------------------------------------------------------------------- forward stats
__global__ void __launch_bounds__(kStatThreads)
gn_stats_kernel(const __nv_bfloat16* __restrict__ r, long long V, int C, float* __restrict__ partial) {
  extern __shared__ float sh[];  // [vpi][C][2]
  const int C8 = C >> 3;
  const int vpi = kStatThreads / C8;
  const int oct = threadIdx.x % C8;
  const int vloc = threadIdx.x / C8;
  const int n = blockIdx.y;
  const __nv_bfloat16* rn = r + (size_t)n * V * C;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (vloc < vpi) {
    for (long long v = (long long)blockIdx.x * vpi + vloc; v < V; v += (long long)gridDim.x * vpi) {
      const f8 x = unpack8(ldg16(rn + v * C + oct * 8));
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += x.v[i]; q[i] = fmaf(x.v[i], x.v[i], q[i]); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sh[((size_t)vloc * C + oct * 8 + i) * 2 + 0] = s[i];
      sh[((size_t)vloc * C + oct * 8 + i) * 2 + 1] = q[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < vpi; ++j) { a += sh[((size_t)j * C + c) * 2]; b += sh[((size_t)j * C + c) * 2 + 1]; }
    float* dst = partial + (((size_t)n * gridDim.x + blockIdx.x) * C + c) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

// one block per (group, n)
__global__ void __launch_bounds__(128)
gn_finalize_kernel(const float* __restrict__ partial, int nblk, int C, int G, long long V, float eps,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ mean_rstd /*[N][C][2]*/, float* __restrict__ scale_shift /*[N][C][2]*/) {
  const int g = blockIdx.x, n = blockIdx.y;
  const int cpg = C / G;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblk * cpg; i += blockDim.x) {
    const int blk = i / cpg, c = g * cpg + i % cpg;
    const float* src = partial + (((size_t)n * nblk + blk) * C + c) * 2;
    a += (double)src[0];
    b += (double)src[1];
  }
  __shared__ double sa[128], sb[128];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  const double m = (double)V * cpg;
  const double mean = sa[0] / m;
  double var = sb[0] / m - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  for (int j = threadIdx.x; j < cpg; j += blockDim.x) {
    const int c = g * cpg + j;
    const float sc = rstd * gamma[c];
    mean_rstd[((size_t)n * C + c) * 2 + 0] = (float)mean;
    mean_rstd[((size_t)n * C + c) * 2 + 1] = rstd;
    scale_shift[((size_t)n * C + c) * 2 + 0] = sc;
    scale_shift[((size_t)n * C + c) * 2 + 1] = beta[c] - (float)mean * sc;
  }
}

// ------------------------------------------------------------------------------------------------- forward apply
__global__ void __launch_bounds__(256)
gn_apply_kernel(const __nv_bfloat16* __restrict__ r, long long V, int C, const float* __restrict__ scale_shift,
                __nv_bfloat16* __restrict__ y, int ldy, int y_coff, int N) {
  const int C8 = C >> 3;  // C8 divides 256, so a thread keeps the same channel octet across the grid-stride loop
  const long long total = (long long)N * V * C8;
  const int oct = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % C8);
  int cur_n = -1;
  float sc[8], sh[8];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long nv = i / C8;
    const int n = (int)(nv / V);
    if (n != cur_n) {
      cur_n = n;
      const float4* ss = reinterpret_cast<const float4*>(scale_shift + ((size_t)n * C + oct * 8) * 2);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(ss + k);
        sc[2 * k] = t.x; sh[2 * k] = t.y; sc[2 * k + 1] = t.z; sh[2 * k + 1] = t.w;
      }
    }
    const f8 x = unpack8(ldg16(r + nv * C + oct * 8));
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = fmaf(x.v[k], sc[k], sh[k]);
    stg16(y + nv * ldy + y_coff + oct * 8, pack8(o));
  }
}

// apply + MaxPool3d(2,2,0): one thread per (2x2x2 cell, channel octet); also covers odd tails (no pooled output)
__global__ void __launch_bounds__(256)
gn_apply_pool_kernel(const __nv_bfloat16* __restrict__ r, int N, int D, int H, int W, int C,
                     const float* __restrict__ scale_shift, __nv_bfloat16* __restrict__ y, int ldy, int y_coff,
                     __nv_bfloat16* __restrict__ pooled /*[N][D/2][H/2][W/2][C]*/) {
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
    float sc[8], sh[8];
    const float4* ss = reinterpret_cast<const float4*>(scale_shift + ((size_t)n * C + oct * 8) * 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 q = __ldg(ss + k);
      sc[2 * k] = q.x; sh[2 * k] = q.y; sc[2 * k + 1] = q.z; sh[2 * k + 1] = q.w;
    }
    f8 mx;
#pragma unroll
    for (int k = 0; k < 8; ++k) mx.v[k] = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = 2 * cd + (j >> 2), h = 2 * ch + ((j >> 1) & 1), w = 2 * cw + (j & 1);
      if (d < D && h < H && w < W) {
        const long long v = (((long long)n * D + d) * H + h) * W + w;
        const f8 x = unpack8(ldg16(r + v * C + oct * 8));
        f8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = fmaf(x.v[k], sc[k], sh[k]);
        const uint4 pk = pack8(o);
        stg16(y + v * ldy + y_coff + oct * 8, pk);
        const f8 ro = unpack8(pk);  // pool the bf16-rounded values (what the next layer reads)
#pragma unroll
        for (int k = 0; k < 8; ++k) mx.v[k] = fmaxf(mx.v[k], ro.v[k]);
      }
    }
    if (cd < Dp && ch < Hp && cw < Wp) {
      const long long pv = (((long long)n * Dp + cd) * Hp + ch) * Wp + cw;
      stg16(pooled + pv * C + oct * 8, pack8(mx));
    }
  }
}

// ------------------------------------------------------------------------------------------------- backward stats
__global__ void __launch_bounds__(kStatThreads)
gn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int dy_coff, const __nv_bfloat16* __restrict__ r,
                    long long V, int C, const float* __restrict__ mean_rstd, float* __restrict__ partial) {
  extern __shared__ float sh[];
  const int C8 = C >> 3;
  const int vpi = kStatThreads / C8;
  const int oct = threadIdx.x % C8;
  const int vloc = threadIdx.x / C8;
  const int n = blockIdx.y;
  const __nv_bfloat16* rn = r + (size_t)n * V * C;
  const __nv_bfloat16* dyn = dy + (size_t)n * V * lddy + dy_coff;
  float mu[8], rs[8], s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean_rstd[((size_t)n * C + oct * 8 + i) * 2];
    rs[i] = mean_rstd[((size_t)n * C + oct * 8 + i) * 2 + 1];
    s[i] = 0.f;
    q[i] = 0.f;
  }
  if (vloc < vpi) {
    for (long long v = (long long)blockIdx.x * vpi + vloc; v < V; v += (long long)gridDim.x * vpi) {
      const f8 x = unpack8(ldg16(rn + v * C + oct * 8));
      const f8 g = unpack8(ldg16(dyn + v * lddy + oct * 8));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += g.v[i];
        q[i] = fmaf(g.v[i], (x.v[i] - mu[i]) * rs[i], q[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sh[((size_t)vloc * C + oct * 8 + i) * 2 + 0] = s[i];
      sh[((size_t)vloc * C + oct * 8 + i) * 2 + 1] = q[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < vpi; ++j) { a += sh[((size_t)j * C + c) * 2]; b += sh[((size_t)j * C + c) * 2 + 1]; }
    float* dst = partial + (((size_t)n * gridDim.x + blockIdx.x) * C + c) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

// one block per group; loops over samples. coef[n][c] = {a, b, c0, pad}: dr = a*dy + b*xhat + c0
__global__ void __launch_bounds__(128)
gn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int N, int C, int G, long long V,
                       const float* __restrict__ gamma, const float* __restrict__ mean_rstd,
                       float* __restrict__ coef /*[N][C][4]*/, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int g = blockIdx.x;
  const int cpg = C / G;
  __shared__ double csum[2][64];   // per channel-in-group sums for the current sample (cpg <= 64)
  __shared__ double dgs[64], dbs[64];
  for (int j = threadIdx.x; j < cpg; j += blockDim.x) { dgs[j] = 0.0; dbs[j] = 0.0; }
  __syncthreads();
  for (int n = 0; n < N; ++n) {
    // each channel of the group: reduce over blocks with a sub-team of threads
    for (int j = 0; j < cpg; ++j) {
      const int c = g * cpg + j;
      double a = 0.0, b = 0.0;
      for (int blk = threadIdx.x; blk < nblk; blk += blockDim.x) {
        const float* src = partial + (((size_t)n * nblk + blk) * C + c) * 2;
        a += (double)src[0];
        b += (double)src[1];
      }
      __shared__ double ra[128], rb[128];
      ra[threadIdx.x] = a;
      rb[threadIdx.x] = b;
      __syncthreads();
      for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) { ra[threadIdx.x] += ra[threadIdx.x + o]; rb[threadIdx.x] += rb[threadIdx.x + o]; }
        __syncthreads();
      }
      if (threadIdx.x == 0) { csum[0][j] = ra[0]; csum[1][j] = rb[0]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      double S1 = 0.0, S2 = 0.0;
      for (int j = 0; j < cpg; ++j) {
        const double gm = (double)gamma[g * cpg + j];
        S1 += gm * csum[0][j];
        S2 += gm * csum[1][j];
        dbs[j] += csum[0][j];
        dgs[j] += csum[1][j];
      }
      const double m = (double)V * cpg;
      for (int j = 0; j < cpg; ++j) {
        const int c = g * cpg + j;
        const double rstd = (double)mean_rstd[((size_t)n * C + c) * 2 + 1];
        float* o = coef + ((size_t)n * C + c) * 4;
        o[0] = (float)(rstd * (double)gamma[c]);
        o[1] = (float)(-rstd * S2 / m);
        o[2] = (float)(-rstd * S1 / m);
        o[3] = 0.f;
      }
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < cpg; j += blockDim.x) {
    if (dgamma) dgamma[g * cpg + j] = (float)dgs[j];
    if (dbeta) dbeta[g * cpg + j] = (float)dbs[j];
  }
}

__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int dy_coff, const __nv_bfloat16* __restrict__ r,
                    long long V, int C, int N, const float* __restrict__ mean_rstd, const float* __restrict__ coef,
                    __nv_bfloat16* __restrict__ dr) {
  const int C8 = C >> 3;
  const long long total = (long long)N * V * C8;
  const int oct = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % C8);
  int cur_n = -1;
  float mu[8], rs[8], ca[8], cb[8], cc[8];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long nv = i / C8;
    const int n = (int)(nv / V);
    if (n != cur_n) {
      cur_n = n;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = oct * 8 + k;
        const float2 mr = __ldg(reinterpret_cast<const float2*>(mean_rstd + ((size_t)n * C + c) * 2));
        const float4 cf = __ldg(reinterpret_cast<const float4*>(coef + ((size_t)n * C + c) * 4));
        mu[k] = mr.x; rs[k] = mr.y; ca[k] = cf.x; cb[k] = cf.y; cc[k] = cf.z;
      }
    }
    const f8 x = unpack8(ldg16(r + nv * C + oct * 8));
    const f8 g = unpack8(ldg16(dy + nv * lddy + dy_coff + oct * 8));
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (x.v[k] - mu[k]) * rs[k];
      const float d = fmaf(ca[k], g.v[k], fmaf(cb[k], xh, cc[k]));
      o.v[k] = x.v[k] > 0.f ? d : 0.f;
    }
    stg16(dr + nv * C + oct * 8, pack8(o));
  }
}

static inline int stat_blocks(long long V, int C) {
  const int vpi = kStatThreads / (C / 8);
  long long nb = (V + vpi - 1) / vpi;
  if (nb > kStatBlocks) nb = kStatBlocks;
  return (int)nb;
}
static inline int ew_blocks(long long total) {
  long long nb = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

}  // namespace b2

using namespace b2;

extern "C" long long b2_gn_workspace_bytes(int N, int C) {
  return (long long)N * kStatBlocks * C * 2 * (long long)sizeof(float);
}

static int check_gn_shape(const char* who, int N, long long V, int C, int G) {
  B2_REQUIRE(N > 0 && V > 0, "%s: bad shape", who);
  B2_REQUIRE(C % 8 == 0 && C >= 8 && C / 8 <= kStatThreads && kStatThreads % (C / 8) == 0,
             "%s: C=%d unsupported (C/8 must divide %d)", who, C, kStatThreads);
  B2_REQUIRE(G > 0 && C % G == 0 && C / G <= 64, "%s: groups=%d unsupported for C=%d", who, G, C);
  return B2_OK;
}

// r: bf16 [N][V][C] dense.  Outputs mean_rstd [N][C][2], scale_shift [N][C][2] (fp32).
extern "C" int b2_relu_gn_stats(const void* r, int N, long long V, int C, int G, float eps, const float* gamma,
                                const float* beta, float* mean_rstd, float* scale_shift, void* workspace,
                                long long workspace_bytes, cudaStream_t stream) {
  B2_REQUIRE(r && gamma && beta && mean_rstd && scale_shift && workspace, "b2_relu_gn_stats: null pointer");
  int rc = check_gn_shape("b2_relu_gn_stats", N, V, C, G);
  if (rc) return rc;
  B2_REQUIRE(workspace_bytes >= b2_gn_workspace_bytes(N, C), "b2_relu_gn_stats: workspace too small");
  const int nblk = stat_blocks(V, C);
  const int vpi = kStatThreads / (C / 8);
  const size_t sh = (size_t)vpi * C * 2 * sizeof(float);
  gn_stats_kernel<<<dim3(nblk, N), kStatThreads, sh, stream>>>(reinterpret_cast<const __nv_bfloat16*>(r), V, C,
                                                               reinterpret_cast<float*>(workspace));
  B2_CHECK_CUDA(cudaGetLastError());
  gn_finalize_kernel<<<dim3(G, N), 128, 0, stream>>>(reinterpret_cast<const float*>(workspace), nblk, C, G, V, eps,
                                                     gamma, beta, mean_rstd, scale_shift);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

// y[(n,v)*ldy + y_coff + c] = r*scale + shift.  If pooled != NULL also writes MaxPool3d(2) of y (needs D,H,W).
extern "C" int b2_relu_gn_apply(const void* r, int N, int D, int H, int W, int C, const float* scale_shift, void* y,
                                int ldy, int y_coff, void* pooled, cudaStream_t stream) {
  B2_REQUIRE(r && scale_shift && y, "b2_relu_gn_apply: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldy % 8 == 0 && y_coff % 8 == 0, "b2_relu_gn_apply: channel counts must be multiples of 8");
  const long long V = (long long)D * H * W;
  if (pooled) {
    const long long total = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
    gn_apply_pool_kernel<<<ew_blocks(total), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(r), N, D, H, W, C, scale_shift, reinterpret_cast<__nv_bfloat16*>(y), ldy,
        y_coff, reinterpret_cast<__nv_bfloat16*>(pooled));
  } else {
    const long long total = (long long)N * V * (C / 8);
    gn_apply_kernel<<<ew_blocks(total), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(r), V, C, scale_shift,
                                                          reinterpret_cast<__nv_bfloat16*>(y), ldy, y_coff, N);
  }
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

// dr = relu'(r) * GroupNorm-backward(dy);  dgamma/dbeta fp32 [C] (overwritten; may be NULL)
extern "C" int b2_relu_gn_bwd(const void* dy, int lddy, int dy_coff, const void* r, int N, long long V, int C, int G,
                              const float* gamma, const float* mean_rstd, void* dr, float* dgamma, float* dbeta,
                              void* workspace, long long workspace_bytes, cudaStream_t stream) {
  B2_REQUIRE(dy && r && gamma && mean_rstd && dr && workspace, "b2_relu_gn_bwd: null pointer");
  int rc = check_gn_shape("b2_relu_gn_bwd", N, V, C, G);
  if (rc) return rc;
  B2_REQUIRE(lddy % 8 == 0 && dy_coff % 8 == 0, "b2_relu_gn_bwd: lddy/dy_coff must be multiples of 8");
  const long long need = b2_gn_workspace_bytes(N, C) + (long long)N * C * 4 * (long long)sizeof(float);
  B2_REQUIRE(workspace_bytes >= need, "b2_relu_gn_bwd: workspace %lld < %lld", workspace_bytes, need);
  float* partial = reinterpret_cast<float*>(workspace);
  float* coef = partial + (size_t)N * kStatBlocks * C * 2;
  const int nblk = stat_blocks(V, C);
  const int vpi = kStatThreads / (C / 8);
  const size_t sh = (size_t)vpi * C * 2 * sizeof(float);
  gn_bwd_stats_kernel<<<dim3(nblk, N), kStatThreads, sh, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), lddy, dy_coff, reinterpret_cast<const __nv_bfloat16*>(r), V, C,
      mean_rstd, partial);
  B2_CHECK_CUDA(cudaGetLastError());
  gn_bwd_finalize_kernel<<<G, 128, 0, stream>>>(partial, nblk, N, C, G, V, gamma, mean_rstd, coef, dgamma, dbeta);
  B2_CHECK_CUDA(cudaGetLastError());
  const long long total = (long long)N * V * (C / 8);
  gn_bwd_apply_kernel<<<ew_blocks(total), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy, dy_coff,
                                                            reinterpret_cast<const __nv_bfloat16*>(r), V, C, N,
                                                            mean_rstd, coef, reinterpret_cast<__nv_bfloat16*>(dr));
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" long long b2_relu_gn_bwd_workspace_bytes(int N, int C) {
  return b2_gn_workspace_bytes(N, C) + (long long)N * C * 4 * (long long)sizeof(float);
}
