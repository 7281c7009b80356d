// encoders.0.conv1: Cin = 1, K = 27 — a bandwidth-bound direct convolution (not tensor-core shaped).
//   fwd  : y[v, co] = relu( sum_tap x[v + off(tap)] * w[co][tap] )      x fp32 [N,D,H,W], y bf16 [N,D,H,W,Cout]
//   wgrad: dw[co][tap] = sum_v dy[v, co] * x[v + off(tap)]
// The input is a binary skeleton (~3 % ones): taps whose input voxel is 0 are skipped.
#include "common.h"
#include "ptx.cuh"

namespace b2 {

static constexpr int kMaxC1 = 64;

template <int COUT>
__global__ void __launch_bounds__(128)
conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w /*[COUT][27]*/,
                      __nv_bfloat16* __restrict__ y, int N, int D, int H, int W, int ldy, int y_coff, int relu) {
  __shared__ float ws[27][COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) ws[i % 27][i / 27] = w[i];  // w[co*27+tap]
  __syncthreads();
  const long long V = (long long)N * D * H * W;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < V;
       v += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(v % W);
    long long r = v / W;
    const int hq = (int)(r % H);
    r /= H;
    const int dq = (int)(r % D);
    const long long nbase = (r / D) * (long long)D * H * W;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 27; ++tap) {
      const int d = dq + tap / 9 - 1, h = hq + (tap / 3) % 3 - 1, ww = wq + tap % 3 - 1;
      if ((unsigned)d >= (unsigned)D || (unsigned)h >= (unsigned)H || (unsigned)ww >= (unsigned)W) continue;
      const float xv = __ldg(x + nbase + ((long long)d * H + h) * W + ww);
      if (xv != 0.f) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = fmaf(xv, ws[tap][c], acc[c]);
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(y + v * ldy + y_coff);
#pragma unroll
    for (int j = 0; j < COUT / 8; ++j) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = relu ? fmaxf(acc[8 * j + e], 0.f) : acc[8 * j + e];
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      dst[j] = o;
    }
  }
}

// dw[co][tap] = sum_u x[u] * dy[u - off(tap), co]: a warp scans 32 consecutive INPUT voxels, and for every non-zero
// one (~3 % of a skeleton volume) adds the 27 neighbouring dy rows (64-byte coalesced loads, lane = channel).
__global__ void __launch_bounds__(256)
conv_first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int lddy, int dy_coff,
                        float* __restrict__ partial /*[grid][27][Cout]*/, int N, int D, int H, int W, int Cout) {
  __shared__ float red[27][kMaxC1];
  for (int i = threadIdx.x; i < 27 * kMaxC1; i += blockDim.x) red[i / kMaxC1][i % kMaxC1] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long V = (long long)N * D * H * W;
  const long long gw = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * warps_per_block;
  float acc0[27], acc1[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) { acc0[t] = 0.f; acc1[t] = 0.f; }
  const bool has1 = (lane + 32) < Cout;
  const bool has0 = lane < Cout;
  const __nv_bfloat16* dyc = dy + dy_coff;
  for (long long base = gw * 32; base < V; base += nw * 32) {
    const long long u = base + lane;
    const float xu = (u < V) ? __ldg(x + u) : 0.f;
    unsigned any = __ballot_sync(0xffffffffu, xu != 0.f);
    while (any) {
      const int src = __ffs(any) - 1;
      any &= any - 1;
      const float xv = __shfl_sync(0xffffffffu, xu, src);
      const long long uu = base + src;
      const int wq = (int)(uu % W);
      long long r = uu / W;
      const int hq = (int)(r % H);
      r /= H;
      const int dq = (int)(r % D);
      const long long nbase = (r / D) * (long long)D * H * W;
      // issue all 27 row loads first (independent, predicated), then accumulate: the loop is latency-bound otherwise
      float g0[27], g1[27];
#pragma unroll
      for (int tap = 0; tap < 27; ++tap) {
        // output voxel v with v + off(tap) = u
        const int d = dq - (tap / 9 - 1), h = hq - ((tap / 3) % 3 - 1), ww = wq - (tap % 3 - 1);
        const bool ok = (unsigned)d < (unsigned)D && (unsigned)h < (unsigned)H && (unsigned)ww < (unsigned)W;
        const __nv_bfloat16* row = dyc + (nbase + ((long long)(ok ? d : dq) * H + (ok ? h : hq)) * W + (ok ? ww : wq)) * lddy;
        g0[tap] = (ok && has0) ? __bfloat162float(row[lane]) : 0.f;
        g1[tap] = (ok && has1) ? __bfloat162float(row[lane + 32]) : 0.f;
      }
#pragma unroll
      for (int tap = 0; tap < 27; ++tap) {
        acc0[tap] = fmaf(xv, g0[tap], acc0[tap]);
        acc1[tap] = fmaf(xv, g1[tap], acc1[tap]);
      }
    }
  }
  // deterministic block reduction: warps add their registers one after another (no float atomics)
  for (int w = 0; w < warps_per_block; ++w) {
    if ((int)(threadIdx.x >> 5) == w) {
#pragma unroll
      for (int tap = 0; tap < 27; ++tap) {
        if (has0) red[tap][lane] += acc0[tap];
        if (has1) red[tap][lane + 32] += acc1[tap];
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x)
    partial[(size_t)blockIdx.x * 27 * Cout + i] = red[i / Cout][i % Cout];
}

__global__ void conv_first_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nblocks,
                                               int Cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = tap*Cout + co
  if (i >= 27 * Cout) return;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) acc += (double)partial[(size_t)b * 27 * Cout + i];
  const int tap = i / Cout, co = i % Cout;
  dw[co * 27 + tap] = (float)acc;
}

static constexpr int kFirstWgradBlocks = 296;

}  // namespace b2

using namespace b2;

extern "C" int b2_conv3d_first_fwd(const float* x, const float* w, void* y, int ldy, int y_coff, int N, int D, int H,
                                   int W, int Cout, int relu, cudaStream_t stream) {
  B2_REQUIRE(x && w && y, "b2_conv3d_first_fwd: null pointer");
  B2_REQUIRE(ldy % 8 == 0 && y_coff % 8 == 0, "b2_conv3d_first_fwd: ldy/y_coff must be multiples of 8");
  const long long V = (long long)N * D * H * W;
  long long blocks = (V + 127) / 128;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
  switch (Cout) {
    case 16: conv_first_fwd_kernel<16><<<(unsigned)blocks, 128, 0, stream>>>(x, w, yy, N, D, H, W, ldy, y_coff, relu); break;
    case 32: conv_first_fwd_kernel<32><<<(unsigned)blocks, 128, 0, stream>>>(x, w, yy, N, D, H, W, ldy, y_coff, relu); break;
    case 64: conv_first_fwd_kernel<64><<<(unsigned)blocks, 128, 0, stream>>>(x, w, yy, N, D, H, W, ldy, y_coff, relu); break;
    default:
      set_error("b2_conv3d_first_fwd: Cout=%d unsupported (16, 32 or 64)", Cout);
      return B2_ERR_UNSUPPORTED;
  }
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" long long b2_conv3d_first_wgrad_workspace_bytes(int Cout) {
  return (long long)kFirstWgradBlocks * 27 * Cout * (long long)sizeof(float);
}

extern "C" int b2_conv3d_first_wgrad(const float* x, const void* dy, int lddy, int dy_coff, float* dw, void* workspace,
                                     long long workspace_bytes, int N, int D, int H, int W, int Cout,
                                     cudaStream_t stream) {
  B2_REQUIRE(x && dy && dw && workspace, "b2_conv3d_first_wgrad: null pointer");
  B2_REQUIRE(Cout >= 1 && Cout <= kMaxC1, "b2_conv3d_first_wgrad: Cout=%d unsupported (<= 64)", Cout);
  B2_REQUIRE(workspace_bytes >= b2_conv3d_first_wgrad_workspace_bytes(Cout), "b2_conv3d_first_wgrad: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  conv_first_wgrad_kernel<<<kFirstWgradBlocks, 256, 0, stream>>>(x, reinterpret_cast<const __nv_bfloat16*>(dy), lddy,
                                                                 dy_coff, partial, N, D, H, W, Cout);
  B2_CHECK_CUDA(cudaGetLastError());
  conv_first_wgrad_reduce_kernel<<<(27 * Cout + 127) / 128, 128, 0, stream>>>(partial, dw, kFirstWgradBlocks, Cout);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
