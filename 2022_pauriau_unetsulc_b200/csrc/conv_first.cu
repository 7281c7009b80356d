// encoders.0.conv1: Cin = 1, K = 27 — a bandwidth-bound direct convolution (not tensor-core shaped).
//   fwd  : y[v, co] = relu( sum_tap x[v + off(tap)] * w[co][tap] )      x fp32 [N,D,H,W], y bf16 [N,D,H,W,Cout]
//   wgrad: dw[co][tap] = sum_v dy[v, co] * x[v + off(tap)]
// The input is a binary skeleton (~3 % ones): taps whose input voxel is 0 are skipped.
#include "common.h"
#include "ptx.cuh"

namespace b2 {

static constexpr int kMaxC1 = 64;

// One warp per 32 consecutive voxels, lane = voxel.  All 27 neighbour loads are issued before any arithmetic (the
// round-1 kernel walked the taps one dependent L1 load at a time and was latency-bound at 680 GB/s); a tap is skipped
// when no lane of the warp has a non-zero input there.  When stat_partial != NULL the kernel also accumulates, per
// channel, sum and sum of squares of the STORED (ReLU'd, bf16-rounded) outputs — the GroupNorm statistics — into
// one fp32 [COUT][2] row per block (same partial layout as b2_conv3d_igemm_stats; finalised by b2_relu_gn_finalize).
template <int COUT>
__global__ void __launch_bounds__(128)
conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w /*[COUT][27]*/,
                      __nv_bfloat16* __restrict__ y, int N, int D, int H, int W, int ldy, int y_coff, int relu,
                      float* __restrict__ stat_partial) {
  __shared__ __align__(16) float ws[27][COUT];
  __shared__ float2 sred[4][COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) ws[i % 27][i / 27] = w[i];  // w[co*27+tap]
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long V = (long long)N * D * H * W;
  float st_s[COUT / 32], st_q[COUT / 32];   // lane L: channel 32*k + L, summed over this warp's voxels
#pragma unroll
  for (int k = 0; k < COUT / 32; ++k) { st_s[k] = 0.f; st_q[k] = 0.f; }
  for (long long v0 = ((long long)blockIdx.x * 4 + warp) * 32; v0 < V; v0 += (long long)gridDim.x * 128) {
    const long long v = v0 + lane;
    const bool active = v < V;
    const long long vv = active ? v : V - 1;
    const int wq = (int)(vv % W);
    long long r = vv / W;
    const int hq = (int)(r % H);
    r /= H;
    const int dq = (int)(r % D);
    const float* xc = x + vv;
    float xv[27];
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      const int dd = tap / 9 - 1, dh = (tap / 3) % 3 - 1, dw = tap % 3 - 1;
      const bool ok = active && (unsigned)(dq + dd) < (unsigned)D && (unsigned)(hq + dh) < (unsigned)H &&
                      (unsigned)(wq + dw) < (unsigned)W;
      xv[tap] = ok ? __ldg(xc + ((long long)dd * H + dh) * W + dw) : 0.f;
    }
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      if (__any_sync(0xffffffffu, xv[tap] != 0.f)) {
        const float4* wr = reinterpret_cast<const float4*>(&ws[tap][0]);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 q = wr[c4];
          acc[4 * c4 + 0] = fmaf(xv[tap], q.x, acc[4 * c4 + 0]);
          acc[4 * c4 + 1] = fmaf(xv[tap], q.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(xv[tap], q.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(xv[tap], q.w, acc[4 * c4 + 3]);
        }
      }
    }
    if (relu) {
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[c] = fmaxf(acc[c], 0.f);
    }
    if (active) {
      uint4* dst = reinterpret_cast<uint4*>(y + v * ldy + y_coff);
#pragma unroll
      for (int j = 0; j < COUT / 8; ++j) {
        uint4 o;
        o.x = pack_bf16x2(acc[8 * j + 0], acc[8 * j + 1]);
        o.y = pack_bf16x2(acc[8 * j + 2], acc[8 * j + 3]);
        o.z = pack_bf16x2(acc[8 * j + 4], acc[8 * j + 5]);
        o.w = pack_bf16x2(acc[8 * j + 6], acc[8 * j + 7]);
        dst[j] = o;
      }
    }
    if (stat_partial != nullptr) {
#pragma unroll
      for (int k = 0; k < COUT / 32; ++k) {
        float xs[32], xq[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float f = active ? __bfloat162float(__float2bfloat16_rn(acc[32 * k + e])) : 0.f;
          xs[e] = f;
          xq[e] = f * f;
        }
        warp_column_sums(xs, lane);
        warp_column_sums(xq, lane);
        st_s[k] += xs[0];
        st_q[k] += xq[0];
      }
    }
  }
  if (stat_partial != nullptr) {
#pragma unroll
    for (int k = 0; k < COUT / 32; ++k) sred[warp][32 * k + lane] = make_float2(st_s[k], st_q[k]);
    __syncthreads();
    for (int c = threadIdx.x; c < COUT; c += blockDim.x) {
      const float2 a = sred[0][c], b = sred[1][c], cc = sred[2][c], d = sred[3][c];
      reinterpret_cast<float2*>(stat_partial)[(size_t)blockIdx.x * COUT + c] =
          make_float2((a.x + b.x) + (cc.x + d.x), (a.y + b.y) + (cc.y + d.y));
    }
  }
}

// dw[co][tap] = sum_u x[u] * dy[u - off(tap), co] over the NON-ZERO input voxels u (~3 % of a skeleton volume).
// Every block owns a contiguous range of input voxels: it first compacts the non-zero ones of a 2048-voxel
// sub-chunk into shared memory (ballot + prefix: ascending voxel order, deterministic), then warp w accumulates taps
// w, w+8, w+16, w+24 over that list (lane = output channel, 64-byte coalesced dy rows, four voxels x four taps of
// independent loads in flight).  The round-1 kernel scanned and gathered one voxel at a time per warp and was
// latency-bound (0.18 ms for 31 k voxels).  No atomics: a (tap, channel) sum lives in one lane's register.
static constexpr int kFwChunk = 2048;

__global__ void __launch_bounds__(256)
conv_first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int lddy, int dy_coff,
                        float* __restrict__ partial /*[grid][27][Cout]*/, int N, int D, int H, int W, int Cout,
                        long long per_block) {
  __shared__ int s_u[kFwChunk];
  __shared__ uint32_t s_dhw[kFwChunk];
  __shared__ float s_x[kFwChunk];
  __shared__ int s_cnt[kFwChunk / 32 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long V = (long long)N * D * H * W;
  const long long r_begin = (long long)blockIdx.x * per_block;
  const long long r_end = (r_begin + per_block < V) ? r_begin + per_block : V;
  const bool has0 = lane < Cout, has1 = (lane + 32) < Cout;
  const __nv_bfloat16* dyc = dy + dy_coff;
  float acc0[4], acc1[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) { acc0[t] = 0.f; acc1[t] = 0.f; }
  // this warp's taps and their (dd, dh, dw)
  int tdd[4], tdh[4], tdw[4], toff[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tap = warp + 8 * t;
    tdd[t] = tap / 9 - 1; tdh[t] = (tap / 3) % 3 - 1; tdw[t] = tap % 3 - 1;
    toff[t] = (tdd[t] * H + tdh[t]) * W + tdw[t];
  }
  for (long long c0 = r_begin; c0 < r_end; c0 += kFwChunk) {
    // ---- compaction of the non-zero voxels of [c0, c0 + kFwChunk) in ascending order
    float xv[kFwChunk / 256];
    unsigned bal[kFwChunk / 256];
#pragma unroll
    for (int k = 0; k < kFwChunk / 256; ++k) {
      const long long u = c0 + k * 256 + threadIdx.x;
      xv[k] = (u < r_end) ? __ldg(x + u) : 0.f;
      bal[k] = __ballot_sync(0xffffffffu, xv[k] != 0.f);
      if (lane == 0) s_cnt[k * 8 + warp] = __popc(bal[k]);
    }
    __syncthreads();
    if (warp == 0) {   // exclusive prefix over the 64 (k, warp) counts
      int a = s_cnt[lane], b = s_cnt[lane + 32];
      int ia = a, ib = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
      }
      const int tot_a = __shfl_sync(0xffffffffu, ia, 31);
      s_cnt[lane] = ia - a;
      s_cnt[lane + 32] = tot_a + ib - b;
      if (lane == 31) s_cnt[64] = tot_a + ib;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFwChunk / 256; ++k) {
      if (xv[k] != 0.f) {
        const int pos = s_cnt[k * 8 + warp] + __popc(bal[k] & ((1u << lane) - 1u));
        const long long u = c0 + k * 256 + threadIdx.x;
        const int wq = (int)(u % W);
        long long r = u / W;
        const int hq = (int)(r % H);
        r /= H;
        const int dq = (int)(r % D);
        s_u[pos] = (int)u;
        s_dhw[pos] = ((uint32_t)dq << 20) | ((uint32_t)hq << 10) | (uint32_t)wq;
        s_x[pos] = xv[k];
      }
    }
    __syncthreads();
    const int cnt = s_cnt[64];
    // ---- gather: 4 voxels x (up to) 4 taps of loads in flight per warp
    for (int j0 = 0; j0 < cnt; j0 += 4) {
      float g0[4][4], g1[4][4], xs[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = (j0 + jj < cnt) ? j0 + jj : cnt - 1;
        const bool live = j0 + jj < cnt;
        const int u = s_u[j];
        const uint32_t c = s_dhw[j];
        const int dq = (int)(c >> 20), hq = (int)((c >> 10) & 1023u), wq = (int)(c & 1023u);
        xs[jj] = live ? s_x[j] : 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          // output voxel v with v + off(tap) = u
          const bool ok = live && (warp + 8 * t < 27) && (unsigned)(dq - tdd[t]) < (unsigned)D &&
                          (unsigned)(hq - tdh[t]) < (unsigned)H && (unsigned)(wq - tdw[t]) < (unsigned)W;
          const __nv_bfloat16* row = dyc + (long long)(ok ? u - toff[t] : u) * lddy;
          g0[jj][t] = (ok && has0) ? __bfloat162float(row[lane]) : 0.f;
          g1[jj][t] = (ok && has1) ? __bfloat162float(row[lane + 32]) : 0.f;
        }
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc0[t] = fmaf(xs[jj], g0[jj][t], acc0[t]);
          acc1[t] = fmaf(xs[jj], g1[jj][t], acc1[t]);
        }
    }
    __syncthreads();   // the list is rebuilt by the next sub-chunk
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tap = warp + 8 * t;
    if (tap < 27) {
      float* dst = partial + ((size_t)blockIdx.x * 27 + tap) * Cout;
      if (has0) dst[lane] = acc0[t];
      if (has1) dst[lane + 32] = acc1[t];
    }
  }
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

static constexpr int kFirstWgradBlocks = 592;
static constexpr int kFirstFwdStatBlocks = 148 * 4;

}  // namespace b2

using namespace b2;

static int conv_first_fwd_impl(const float* x, const float* w, void* y, int ldy, int y_coff, int N, int D, int H,
                               int W, int Cout, int relu, float* stat_partial, int* n_partials, cudaStream_t stream) {
  B2_REQUIRE(x && w && y, "b2_conv3d_first_fwd: null pointer");
  B2_REQUIRE(ldy % 8 == 0 && y_coff % 8 == 0, "b2_conv3d_first_fwd: ldy/y_coff must be multiples of 8");
  const long long V = (long long)N * D * H * W;
  long long blocks = (V + 127) / 128;
  const long long cap = stat_partial ? (long long)kFirstFwdStatBlocks : (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
  switch (Cout) {
    case 32: conv_first_fwd_kernel<32><<<(unsigned)blocks, 128, 0, stream>>>(x, w, yy, N, D, H, W, ldy, y_coff, relu, stat_partial); break;
    case 64: conv_first_fwd_kernel<64><<<(unsigned)blocks, 128, 0, stream>>>(x, w, yy, N, D, H, W, ldy, y_coff, relu, stat_partial); break;
    default:
      set_error("b2_conv3d_first_fwd: Cout=%d unsupported (32 or 64)", Cout);
      return B2_ERR_UNSUPPORTED;
  }
  B2_CHECK_CUDA(cudaGetLastError());
  if (n_partials) *n_partials = (int)blocks;
  return B2_OK;
}

extern "C" int b2_conv3d_first_fwd(const float* x, const float* w, void* y, int ldy, int y_coff, int N, int D, int H,
                                   int W, int Cout, int relu, cudaStream_t stream) {
  return conv_first_fwd_impl(x, w, y, ldy, y_coff, N, D, H, W, Cout, relu, nullptr, nullptr, stream);
}

// Same, with the GroupNorm statistics of the stored output fused in (batch 1): stat_partial fp32
// [b2_conv3d_first_stats_max_partials()][Cout][2]; *n_partials (HOST) receives the rows written.
extern "C" int b2_conv3d_first_stats_max_partials(void) { return kFirstFwdStatBlocks; }
extern "C" int b2_conv3d_first_fwd_stats(const float* x, const float* w, void* y, int ldy, int y_coff, int N, int D,
                                         int H, int W, int Cout, int relu, float* stat_partial, int* n_partials,
                                         cudaStream_t stream) {
  B2_REQUIRE(stat_partial && n_partials, "b2_conv3d_first_fwd_stats: null pointer");
  B2_REQUIRE(N == 1, "b2_conv3d_first_fwd_stats: fused statistics need batch 1");
  return conv_first_fwd_impl(x, w, y, ldy, y_coff, N, D, H, W, Cout, relu, stat_partial, n_partials, stream);
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
  B2_REQUIRE(D < 1024 && H < 1024 && W < 1024 && (long long)N * D * H * W < (1LL << 31),
             "b2_conv3d_first_wgrad: volume %dx%dx%dx%d too large", N, D, H, W);
  float* partial = reinterpret_cast<float*>(workspace);
  const long long V = (long long)N * D * H * W;
  long long per_block = (V + kFirstWgradBlocks - 1) / kFirstWgradBlocks;
  per_block = (per_block + 255) / 256 * 256;
  const int blocks = (int)((V + per_block - 1) / per_block);
  conv_first_wgrad_kernel<<<blocks, 256, 0, stream>>>(x, reinterpret_cast<const __nv_bfloat16*>(dy), lddy, dy_coff,
                                                     partial, N, D, H, W, Cout, per_block);
  B2_CHECK_CUDA(cudaGetLastError());
  conv_first_wgrad_reduce_kernel<<<(27 * Cout + 63) / 64, 64, 0, stream>>>(partial, dw, blocks, Cout);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
