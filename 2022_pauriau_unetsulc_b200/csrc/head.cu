// final_conv (1x1x1, Cin<=64 -> Cout<=64, bias) fused with what follows it in the reference:
//   * training  : nn.CrossEntropyLoss(ignore_index=-1) + torch.max(out,1)  (reference training.py:206-211)
//                 -> only labelled voxels (2-4 % of the volume) are ever touched: loss, argmax, d(logits),
//                    dW, db and the sparse dX rows in ONE kernel (head_ce_kernel)
//   * inference : Softmax(dim=1) scores gathered at the skeleton voxels (reference pattern_class.py:266-277)
//   * dense     : the full [N,Cout,D,H,W] fp32 tensor, for callers that use the nn.Module surface directly
// x is NDHWC bf16 (dense, ld = Cin); weights/bias fp32.
#include "common.h"
#include "vec.cuh"

namespace b2 {

static constexpr int kMaxCo = 64;
static constexpr int kCeBlocks = 296;
static constexpr int kCeThreads = 256;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void count_labelled_kernel(const long long* __restrict__ labels, long long n, int* __restrict__ count) {
  pdl_prologue();
  int c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += (labels[i] >= 0) ? 1 : 0;
  c = (int)warp_sum((float)c);  // < 2^24 per warp pass, exact
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

struct WarpHead {
  // per-lane state for one voxel: logits of channels (lane, lane+32)
  float l0, l1;
};

// logits for one voxel; xp = this lane's packed bf16 pair (channels 2*lane, 2*lane+1); Wt = [Cin][kMaxCo] in smem
template <int CIN>
__device__ __forceinline__ WarpHead warp_logits(uint32_t xp, const float* __restrict__ Wt, const float* __restrict__ bs,
                                                int lane) {
  WarpHead r;
  r.l0 = bs[lane];
  r.l1 = bs[lane + 32];
#pragma unroll
  for (int ci = 0; ci < CIN; ci += 2) {
    const uint32_t p = __shfl_sync(0xffffffffu, xp, ci >> 1);
    const float xa = __uint_as_float(p << 16), xb = __uint_as_float(p & 0xffff0000u);
    r.l0 = fmaf(xa, Wt[ci * kMaxCo + lane], r.l0);
    r.l1 = fmaf(xa, Wt[ci * kMaxCo + lane + 32], r.l1);
    r.l0 = fmaf(xb, Wt[(ci + 1) * kMaxCo + lane], r.l0);
    r.l1 = fmaf(xb, Wt[(ci + 1) * kMaxCo + lane + 32], r.l1);
  }
  return r;
}

// softmax over Cout channels spread as (lane, lane+32); returns probabilities in p0/p1, lse in *lse
__device__ __forceinline__ void warp_softmax(float l0, float l1, int Cout, int lane, float& p0, float& p1, float& lse) {
  const bool v0 = lane < Cout, v1 = lane + 32 < Cout;
  const float m = warp_max(fmaxf(v0 ? l0 : -INFINITY, v1 ? l1 : -INFINITY));
  const float e0 = v0 ? expf(l0 - m) : 0.f, e1 = v1 ? expf(l1 - m) : 0.f;
  const float s = warp_sum(e0 + e1);
  p0 = e0 / s;
  p1 = e1 / s;
  lse = m + logf(s);
}

// argmax with ties -> lowest index (torch.max semantics)
__device__ __forceinline__ int warp_argmax(float l0, float l1, int Cout, int lane) {
  float bv = (lane < Cout) ? l0 : -INFINITY;
  int bi = lane;
  if (lane + 32 < Cout && l1 > bv) { bv = l1; bi = lane + 32; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  return bi;
}

__device__ __forceinline__ void load_head_weights(const float* __restrict__ W, const float* __restrict__ b, int Cin,
                                                  int Cout, float* Wt, float* Ws, float* bs) {
  for (int i = threadIdx.x; i < Cin * kMaxCo; i += blockDim.x) {
    const int ci = i / kMaxCo, co = i % kMaxCo;
    Wt[i] = (co < Cout) ? W[co * Cin + ci] : 0.f;
  }
  if (Ws)
    for (int i = threadIdx.x; i < kMaxCo * Cin; i += blockDim.x) Ws[i] = (i < Cout * Cin) ? W[i] : 0.f;
  for (int i = threadIdx.x; i < kMaxCo; i += blockDim.x) bs[i] = (i < Cout && b) ? b[i] : 0.f;
}

// ------------------------------------------------------------------------------------------------- fused head + CE
// Every block owns a contiguous range of voxels.  Per 2048-voxel sub-chunk it compacts the LABELLED voxels (2-4 %)
// into shared memory (ballot + prefix: ascending order, deterministic), then works through them 64 at a time:
//   * four adjacent lanes per voxel: lane q computes the logits of output channels [q*Cout/4, (q+1)*Cout/4) from the
//     voxel's 64-channel row held in registers (weights: conflict-free broadcast float4 reads, rows padded by 4),
//     softmax / argmax / loss combine over the 4 lanes by shuffles, d(logits) goes to shared memory;
//   * the same four lanes compute the voxel's dX row (lane q owns channels 16m + 4q .. +3);
//   * dW / db: all 256 threads as a [64 x CIN] register tile (thread = (co, CIN/4 channels)) accumulate
//     d(logits)^T . x over the 64 staged voxels — the accumulators live for the whole kernel, no atomics.
// The round-1 kernel gave each labelled voxel to a whole warp, one after another (128 accumulator registers per lane,
// 8 warps per SM): latency-bound at 0.19 ms for 31 k voxels.
// partial layout per block: [Cin][kMaxCo] dW (transposed) | [kMaxCo] db | loss
static constexpr int kCeChunk = 2048;
static constexpr int kCeGroup = 64;

template <int CIN>
__global__ void __launch_bounds__(kCeThreads, 2)
head_ce_kernel(const __nv_bfloat16* __restrict__ x, const long long* __restrict__ labels, long long NV,
               const float* __restrict__ W, const float* __restrict__ b, int Cout, const int* __restrict__ count,
               float grad_scale, const float* __restrict__ grad_scale_dev, int compute_grad, int eval_softmax,
               int* __restrict__ preds, __nv_bfloat16* __restrict__ dx, float* __restrict__ partial,
               long long per_block, const __nv_bfloat16* __restrict__ stat_r, long long* __restrict__ stat_acc,
               const float* __restrict__ x_scale_shift) {
  pdl_prologue();
  constexpr int WS = CIN + 4;        // padded row strides (floats)
  constexpr int DS = kMaxCo + 1;
  constexpr int CPT = CIN / 4;       // channels per thread in the dX and dW phases
  extern __shared__ float shm[];
  float* Ws = shm;                            // [kMaxCo][WS]
  float* bs = Ws + kMaxCo * WS;               // [kMaxCo]
  float* ds = bs + kMaxCo;                    // [kCeGroup][DS]  logits, then d(logits)
  float* xs = ds + kCeGroup * DS;             // [kCeGroup][WS]  x rows (fp32)
  float* lossv = xs + kCeGroup * WS;          // [kCeGroup]
  // dX is the gradient at the last GroupNorm output: with stat_acc the kernel also accumulates that layer's
  // GroupNorm-backward statistics (sum dX, sum dX*r) — dX is zero away from the labelled voxels, so these rows are all
  float* dxs = lossv + kCeGroup;              // [kCeGroup][WS]  stored (bf16-rounded) dX rows
  float* rs = dxs + kCeGroup * WS;            // [kCeGroup][WS]  r rows
  int* s_u = reinterpret_cast<int*>(rs + kCeGroup * WS);   // [kCeChunk] voxel offsets relative to r_begin
  int* s_lab = s_u + kCeChunk;                           // [kCeChunk]
  int* s_cnt = s_lab + kCeChunk;                         // [kCeChunk / 32 + 1]
  for (int i = threadIdx.x; i < kMaxCo * WS; i += blockDim.x) {
    const int co = i / WS, ci = i % WS;
    Ws[i] = (co < Cout && ci < CIN) ? W[co * CIN + ci] : 0.f;
  }
  for (int i = threadIdx.x; i < kMaxCo; i += blockDim.x) bs[i] = (i < Cout && b) ? b[i] : 0.f;
  for (int i = threadIdx.x; i < kCeGroup * DS; i += blockDim.x) ds[i] = 0.f;   // columns >= Cout stay zero
  // x_scale_shift != NULL: x holds the last layer's relu(conv) and its GroupNorm apply y = bf16(r*scale + shift) is
  // done here, on the gathered rows only — the dense apply pass over the whole volume is skipped
  __shared__ float s_ss[2 * 64];
  const bool gn_x = x_scale_shift != nullptr;
  if (gn_x)
    for (int i = threadIdx.x; i < 2 * CIN; i += blockDim.x) s_ss[i] = x_scale_shift[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = threadIdx.x >> 2, q = threadIdx.x & 3;   // voxel slot in the group, quarter
  const int cpq = (Cout + 3) >> 2;
  const int co_begin = q * cpq, co_end = min(Cout, co_begin + cpq);
  const int cnt_all = *count;
  const float gsc = grad_scale_dev ? grad_scale * (*grad_scale_dev) : grad_scale;
  const float gs = (cnt_all > 0) ? gsc / (float)cnt_all : 0.f;
  const int wco = threadIdx.x >> 2, wcb = (threadIdx.x & 3) * CPT;   // dW tile of this thread
  float accW[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) accW[i] = 0.f;
  float accb = 0.f, block_loss = 0.f;
  float st_s = 0.f, st_q = 0.f;   // threads < CIN: running (sum dX, sum dX*r) of channel threadIdx.x

  const long long r_begin = (long long)blockIdx.x * per_block;
  const long long r_end = (r_begin + per_block < NV) ? r_begin + per_block : NV;
  for (long long c0 = r_begin; c0 < r_end; c0 += kCeChunk) {
    // ---- compaction of the labelled voxels of [c0, c0 + kCeChunk)
    int lab[kCeChunk / kCeThreads];
    unsigned bal[kCeChunk / kCeThreads];
#pragma unroll
    for (int k = 0; k < kCeChunk / kCeThreads; ++k) {
      const long long u = c0 + k * kCeThreads + threadIdx.x;
      const long long l = (u < r_end) ? labels[u] : -1;
      lab[k] = (l >= 0) ? (int)l : -1;
      bal[k] = __ballot_sync(0xffffffffu, lab[k] >= 0);
      if (lane == 0) s_cnt[k * 8 + warp] = __popc(bal[k]);
    }
    __syncthreads();
    if (warp == 0) {   // exclusive prefix over the 64 (k, warp) counts
      const int a = s_cnt[lane], bb = s_cnt[lane + 32];
      int ia = a, ib = bb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
      }
      const int tot_a = __shfl_sync(0xffffffffu, ia, 31);
      s_cnt[lane] = ia - a;
      s_cnt[lane + 32] = tot_a + ib - bb;
      if (lane == 31) s_cnt[64] = tot_a + ib;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kCeChunk / kCeThreads; ++k) {
      if (lab[k] >= 0) {
        const int pos = s_cnt[k * 8 + warp] + __popc(bal[k] & ((1u << lane) - 1u));
        s_u[pos] = (int)(c0 - r_begin) + k * kCeThreads + threadIdx.x;
        s_lab[pos] = lab[k];
      }
    }
    __syncthreads();
    const int cnt = s_cnt[64];

    for (int g0 = 0; g0 < cnt; g0 += kCeGroup) {
      const int nj = min(kCeGroup, cnt - g0);
      const bool live = j < nj;
      const long long vv = r_begin + (live ? s_u[g0 + j] : s_u[g0]);
      const int label = live ? s_lab[g0 + j] : 0;
      // ---- logits of this lane's output channels
      float xr[CIN];
#pragma unroll
      for (int i8 = 0; i8 < CIN / 8; ++i8) {
        const f8 t = unpack8(ldg16(x + vv * CIN + i8 * 8));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float v = t.v[k];
          if (gn_x) v = __bfloat162float(__float2bfloat16_rn(fmaf(v, s_ss[2 * (i8 * 8 + k)], s_ss[2 * (i8 * 8 + k) + 1])));
          xr[i8 * 8 + k] = v;
        }
      }
      // this lane's quarter of the row, staged for the dW tile (re-loaded: indexing xr[] by q would spill it)
#pragma unroll
      for (int i8 = 0; i8 < CPT / 8; ++i8) {
        const f8 t = unpack8(ldg16(x + vv * CIN + q * CPT + i8 * 8));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = q * CPT + i8 * 8 + k;
          float v = t.v[k];
          if (gn_x) v = __bfloat162float(__float2bfloat16_rn(fmaf(v, s_ss[2 * c], s_ss[2 * c + 1])));
          xs[j * WS + c] = live ? v : 0.f;
        }
      }
      float m = -INFINITY;
      int am = 0;
      for (int co = co_begin; co < co_end; ++co) {
        float a0 = bs[co], a1 = 0.f;
        const float4* wr = reinterpret_cast<const float4*>(Ws + co * WS);
#pragma unroll
        for (int c4 = 0; c4 < CIN / 4; c4 += 2) {
          const float4 w0 = wr[c4], w1 = wr[c4 + 1];
          a0 = fmaf(xr[4 * c4 + 0], w0.x, a0); a0 = fmaf(xr[4 * c4 + 1], w0.y, a0);
          a0 = fmaf(xr[4 * c4 + 2], w0.z, a0); a0 = fmaf(xr[4 * c4 + 3], w0.w, a0);
          a1 = fmaf(xr[4 * c4 + 4], w1.x, a1); a1 = fmaf(xr[4 * c4 + 5], w1.y, a1);
          a1 = fmaf(xr[4 * c4 + 6], w1.z, a1); a1 = fmaf(xr[4 * c4 + 7], w1.w, a1);
        }
        const float l = a0 + a1;
        ds[j * DS + co] = l;
        if (l > m) { m = l; am = co; }   // first maximum wins inside the ascending range
      }
      // combine max / argmax over the 4 lanes of the voxel (ties -> lowest index, torch.max semantics)
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
      }
      float se = 0.f;
      for (int co = co_begin; co < co_end; ++co) se += expf(ds[j * DS + co] - m);
      se += __shfl_xor_sync(0xffffffffu, se, 1);
      se += __shfl_xor_sync(0xffffffffu, se, 2);
      const float lse = m + logf(se);
      __syncwarp();
      float lv;
      if (eval_softmax) {  // reference val phase: CrossEntropyLoss applied to Softmax outputs (training.py:189,205-208)
        float m2 = -INFINITY;
        for (int co = co_begin; co < co_end; ++co) m2 = fmaxf(m2, expf(ds[j * DS + co] - m) / se);
        m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
        m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
        float s2 = 0.f;
        for (int co = co_begin; co < co_end; ++co) s2 += expf(expf(ds[j * DS + co] - m) / se - m2);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        lv = (m2 + logf(s2)) - expf(ds[j * DS + label] - m) / se;
      } else {
        lv = lse - ds[j * DS + label];
      }
      if (q == 0) {
        lossv[j] = live ? lv : 0.f;
        if (live && preds) preds[vv] = am;
      }
      __syncwarp();
      if (compute_grad) {
        for (int co = co_begin; co < co_end; ++co) {
          const float pco = expf(ds[j * DS + co] - m) / se;
          ds[j * DS + co] = live ? (pco - ((co == label) ? 1.f : 0.f)) * gs : 0.f;
        }
        __syncwarp();
        if (dx != nullptr && live) {
          float g[CPT];   // channels 16m + 4q + e
#pragma unroll
          for (int i = 0; i < CPT; ++i) g[i] = 0.f;
          for (int co = 0; co < Cout; ++co) {
            const float dco = ds[j * DS + co];
#pragma unroll
            for (int mm = 0; mm < CPT / 4; ++mm) {
              const float4 w4 = *reinterpret_cast<const float4*>(Ws + co * WS + 16 * mm + 4 * q);
              g[4 * mm + 0] = fmaf(dco, w4.x, g[4 * mm + 0]);
              g[4 * mm + 1] = fmaf(dco, w4.y, g[4 * mm + 1]);
              g[4 * mm + 2] = fmaf(dco, w4.z, g[4 * mm + 2]);
              g[4 * mm + 3] = fmaf(dco, w4.w, g[4 * mm + 3]);
            }
          }
#pragma unroll
          for (int mm = 0; mm < CPT / 4; ++mm) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(g[4 * mm + 0], g[4 * mm + 1]);
            __nv_bfloat162 hi = __floats2bfloat162_rn(g[4 * mm + 2], g[4 * mm + 3]);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo);
            o.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dx + vv * CIN + 16 * mm + 4 * q) = o;
            if (stat_acc != nullptr) {
              const uint2 rr = __ldg(reinterpret_cast<const uint2*>(stat_r + vv * CIN + 16 * mm + 4 * q));
              float* dd = dxs + j * WS + 16 * mm + 4 * q;
              float* rd = rs + j * WS + 16 * mm + 4 * q;
              dd[0] = __uint_as_float(o.x << 16); dd[1] = __uint_as_float(o.x & 0xffff0000u);
              dd[2] = __uint_as_float(o.y << 16); dd[3] = __uint_as_float(o.y & 0xffff0000u);
              rd[0] = __uint_as_float(rr.x << 16); rd[1] = __uint_as_float(rr.x & 0xffff0000u);
              rd[2] = __uint_as_float(rr.y << 16); rd[3] = __uint_as_float(rr.y & 0xffff0000u);
            }
          }
        }
      }
      __syncthreads();
      // ---- dW / db register tile over the staged voxels; loss in a fixed order
      if (compute_grad) {
        for (int jv = 0; jv < nj; ++jv) {
          const float dco = ds[jv * DS + wco];
          const float4* xr4 = reinterpret_cast<const float4*>(xs + jv * WS + wcb);
#pragma unroll
          for (int i4 = 0; i4 < CPT / 4; ++i4) {
            const float4 xv = xr4[i4];
            accW[4 * i4 + 0] = fmaf(dco, xv.x, accW[4 * i4 + 0]);
            accW[4 * i4 + 1] = fmaf(dco, xv.y, accW[4 * i4 + 1]);
            accW[4 * i4 + 2] = fmaf(dco, xv.z, accW[4 * i4 + 2]);
            accW[4 * i4 + 3] = fmaf(dco, xv.w, accW[4 * i4 + 3]);
          }
          if ((threadIdx.x & 3) == 0) accb += dco;
        }
        if (stat_acc != nullptr && dx != nullptr && threadIdx.x < CIN) {
          for (int jv = 0; jv < nj; ++jv) {
            const float g = dxs[jv * WS + threadIdx.x];
            st_s += g;
            st_q = fmaf(g, rs[jv * WS + threadIdx.x], st_q);
          }
        }
      }
      if (warp == 0) {
        float v = lossv[lane] + lossv[lane + 32];
        v = warp_sum(v);
        block_loss += v;
      }
      __syncthreads();
    }
  }
  float* dst = partial + (size_t)blockIdx.x * (kMaxCo * CIN + kMaxCo + 1);
#pragma unroll
  for (int i = 0; i < CPT; ++i) dst[(wcb + i) * kMaxCo + wco] = accW[i];
  if ((threadIdx.x & 3) == 0) dst[kMaxCo * CIN + wco] = accb;
  if (threadIdx.x == 0) dst[kMaxCo * CIN + kMaxCo] = block_loss;
  if (stat_acc != nullptr && compute_grad && dx != nullptr && threadIdx.x < CIN) {
    stat_atomic_add(stat_acc + 4 * threadIdx.x, st_s);
    stat_atomic_add(stat_acc + 4 * threadIdx.x + 2, st_q);
  }
}

// Sums the per-block partials in a fixed order: block = 32 consecutive outputs x 8 row groups (row r goes to group
// r mod 8), fp64 accumulation, the 8 groups combined in ascending order.  (A single thread per output walking all
// 296 rows took 25 us.)
__global__ void __launch_bounds__(256)
head_ce_finalize_kernel(const float* __restrict__ partial, int nblocks, int Cin, int Cout,
                        const int* __restrict__ count, float* __restrict__ dW, float* __restrict__ db,
                        float* __restrict__ loss_out /*[2]: mean loss, sum*/) {
  pdl_prologue();
  __shared__ double red[8][32];
  const int stride = kMaxCo * Cin + kMaxCo + 1;
  const int o = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + o;
  double acc = 0.0;
  if (i < stride) {
    int r = rg;
    for (; r + 24 < nblocks; r += 32) {
      const float a0 = partial[(size_t)r * stride + i], a1 = partial[(size_t)(r + 8) * stride + i];
      const float a2 = partial[(size_t)(r + 16) * stride + i], a3 = partial[(size_t)(r + 24) * stride + i];
      acc += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
    }
    for (; r < nblocks; r += 8) acc += (double)partial[(size_t)r * stride + i];
  }
  red[rg][o] = acc;
  __syncthreads();
  if (rg != 0 || i >= stride) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) acc += red[k][o];
  if (i < kMaxCo * Cin) {
    const int ci = i / kMaxCo, co = i % kMaxCo;   // partials are laid out [ci][co]
    if (dW && co < Cout) dW[co * Cin + ci] = (float)acc;
  } else if (i < kMaxCo * Cin + kMaxCo) {
    const int co = i - kMaxCo * Cin;
    if (db && co < Cout) db[co] = (float)acc;
  } else {
    const int c = *count;
    loss_out[0] = (c > 0) ? (float)(acc / (double)c) : __int_as_float(0x7fc00000);  // NaN like PyTorch when empty
    loss_out[1] = (float)acc;
  }
}

// ------------------------------------------------------------------------------------------------- gather (inference)
template <int CIN>
__global__ void __launch_bounds__(256)
head_gather_kernel(const __nv_bfloat16* __restrict__ x, const long long* __restrict__ index, long long nidx,
                   const float* __restrict__ W, const float* __restrict__ b, int Cout, int softmax,
                   float* __restrict__ scores /*[nidx][Cout]*/, int* __restrict__ preds,
                   const float* __restrict__ x_scale_shift) {
  pdl_prologue();
  extern __shared__ float shm[];
  float* Wt = shm;
  float* bs = Wt + CIN * kMaxCo;
  load_head_weights(W, b, CIN, Cout, Wt, nullptr, bs);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  // deferred GroupNorm apply of the last layer on the gathered rows (see head_ce_kernel): this lane's two channels
  float4 ss = make_float4(1.f, 0.f, 1.f, 0.f);
  const bool gn_x = x_scale_shift != nullptr;
  if (gn_x && lane < CIN / 2) ss = __ldg(reinterpret_cast<const float4*>(x_scale_shift) + lane);
  for (long long k = (long long)blockIdx.x * nwarp + warp; k < nidx; k += (long long)gridDim.x * nwarp) {
    const long long vv = index[k];
    uint32_t xp = 0;
    if (lane < CIN / 2) xp = __ldg(reinterpret_cast<const uint32_t*>(x + vv * CIN) + lane);
    if (gn_x) {
      __nv_bfloat162 y2 = __floats2bfloat162_rn(fmaf(__uint_as_float(xp << 16), ss.x, ss.y),
                                                fmaf(__uint_as_float(xp & 0xffff0000u), ss.z, ss.w));
      xp = *reinterpret_cast<uint32_t*>(&y2);
    }
    WarpHead h = warp_logits<CIN>(xp, Wt, bs, lane);
    const int am = warp_argmax(h.l0, h.l1, Cout, lane);
    float o0 = h.l0, o1 = h.l1;
    if (softmax) {
      float lse;
      warp_softmax(h.l0, h.l1, Cout, lane, o0, o1, lse);
    }
    if (lane < Cout) scores[k * Cout + lane] = o0;
    if (lane + 32 < Cout) scores[k * Cout + lane + 32] = o1;
    if (lane == 0 && preds) preds[k] = am;
  }
}

// ------------------------------------------------------------------------------------------------- dense forward
template <int CIN>
__global__ void __launch_bounds__(128)
head_dense_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long V, int N, const float* __restrict__ W,
                      const float* __restrict__ b, int Cout, int softmax, float* __restrict__ out /*[N][Cout][V]*/) {
  pdl_prologue();
  __shared__ float Ws[kMaxCo * CIN];
  __shared__ float bs[kMaxCo];
  for (int i = threadIdx.x; i < kMaxCo * CIN; i += blockDim.x) Ws[i] = (i < Cout * CIN) ? W[i] : 0.f;
  for (int i = threadIdx.x; i < kMaxCo; i += blockDim.x) bs[i] = (i < Cout && b) ? b[i] : 0.f;
  __syncthreads();
  const long long NV = (long long)N * V;
  for (long long nv = blockIdx.x * (long long)blockDim.x + threadIdx.x; nv < NV;
       nv += (long long)gridDim.x * blockDim.x) {
    float xr[CIN];
#pragma unroll
    for (int j = 0; j < CIN / 8; ++j) {
      const f8 t = unpack8(ldg16(x + nv * CIN + j * 8));
#pragma unroll
      for (int k = 0; k < 8; ++k) xr[j * 8 + k] = t.v[k];
    }
    float lg[kMaxCo];
    float m = -INFINITY;
#pragma unroll
    for (int co = 0; co < kMaxCo; ++co) {
      float acc = bs[co];
      if (co < Cout) {
#pragma unroll
        for (int ci = 0; ci < CIN; ci += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(Ws + co * CIN + ci);
          acc = fmaf(xr[ci], w4.x, acc);
          acc = fmaf(xr[ci + 1], w4.y, acc);
          acc = fmaf(xr[ci + 2], w4.z, acc);
          acc = fmaf(xr[ci + 3], w4.w, acc);
        }
        m = fmaxf(m, acc);
      }
      lg[co] = acc;
    }
    const long long n = nv / V, v = nv % V;
    float* o = out + (size_t)n * Cout * V + v;
    if (softmax) {
      float s = 0.f;
#pragma unroll
      for (int co = 0; co < kMaxCo; ++co)
        if (co < Cout) { lg[co] = __expf(lg[co] - m); s += lg[co]; }
      const float inv = 1.f / s;
#pragma unroll
      for (int co = 0; co < kMaxCo; ++co)
        if (co < Cout) o[(size_t)co * V] = lg[co] * inv;
    } else {
#pragma unroll
      for (int co = 0; co < kMaxCo; ++co)
        if (co < Cout) o[(size_t)co * V] = lg[co];
    }
  }
}

// ------------------------------------------------------------------------------------------------- dense backward
// g: fp32 [N][Cout][V] (d loss / d logits).  dx bf16 [N*V][CIN]; dW/db through per-block partials.
// Rows of g that are entirely zero (unlabelled voxels under CrossEntropyLoss) are skipped for dW.
template <int CIN>
__global__ void __launch_bounds__(128)
head_dense_bwd_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ x, long long V, int N,
                      const float* __restrict__ W, int Cout, __nv_bfloat16* __restrict__ dx,
                      float* __restrict__ partial /*[grid][kMaxCo*CIN + kMaxCo]*/) {
  pdl_prologue();
  extern __shared__ float shm[];
  float* Ws = shm;                          // [kMaxCo][CIN]
  float* gs = Ws + kMaxCo * CIN;            // [128][kMaxCo]
  float* xs = gs + 128 * kMaxCo;            // [128][CIN]
  __shared__ int wcnt[4];
  for (int i = threadIdx.x; i < kMaxCo * CIN; i += blockDim.x) Ws[i] = (i < Cout * CIN) ? W[i] : 0.f;
  constexpr int PAIRS = kMaxCo * CIN / 128;
  float acc[PAIRS];
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) acc[j] = 0.f;
  float accb = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long NV = (long long)N * V;
  const long long nchunks = (NV + 127) / 128;
  for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    __syncthreads();  // previous chunk's gs/xs fully consumed
    const long long nv = chunk * 128 + threadIdx.x;
    const bool inb = nv < NV;
    float gr[kMaxCo];
    bool nz = false;
    if (inb) {
      const long long n = nv / V, v = nv % V;
      const float* gp = g + (size_t)n * Cout * V + v;
#pragma unroll
      for (int co = 0; co < kMaxCo; ++co) {
        gr[co] = (co < Cout) ? __ldg(gp + (size_t)co * V) : 0.f;
        nz |= (gr[co] != 0.f);
      }
    }
    // deterministic compaction of the non-zero rows (ballot + per-warp prefix)
    const unsigned bal = __ballot_sync(0xffffffffu, nz);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; ++w) off += wcnt[w];
    const int c = wcnt[0] + wcnt[1] + wcnt[2] + wcnt[3];
    if (inb) {
      f8 o[CIN / 8];
#pragma unroll
      for (int j = 0; j < CIN / 8; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) o[j].v[k] = 0.f;
      if (nz) {
        const int slot = off + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
        for (int co = 0; co < kMaxCo; ++co) gs[slot * kMaxCo + co] = gr[co];
#pragma unroll
        for (int j = 0; j < CIN / 8; ++j) {
          const f8 t = unpack8(ldg16(x + nv * CIN + j * 8));
#pragma unroll
          for (int k = 0; k < 8; ++k) xs[slot * CIN + j * 8 + k] = t.v[k];
        }
#pragma unroll
        for (int co = 0; co < kMaxCo; ++co) {
          if (co < Cout) {
            const float gc = gr[co];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) o[ci / 8].v[ci % 8] = fmaf(gc, Ws[co * CIN + ci], o[ci / 8].v[ci % 8]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CIN / 8; ++j) stg16(dx + nv * CIN + j * 8, pack8(o[j]));
    }
    __syncthreads();
    for (int k = 0; k < c; ++k) {
#pragma unroll
      for (int j = 0; j < PAIRS; ++j) {
        const int p = threadIdx.x + 128 * j;
        acc[j] = fmaf(gs[k * kMaxCo + p / CIN], xs[k * CIN + p % CIN], acc[j]);
      }
      if (threadIdx.x < kMaxCo) accb += gs[k * kMaxCo + threadIdx.x];
    }
  }
  float* dst = partial + (size_t)blockIdx.x * (kMaxCo * CIN + kMaxCo);
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) dst[threadIdx.x + 128 * j] = acc[j];
  if (threadIdx.x < kMaxCo) dst[kMaxCo * CIN + threadIdx.x] = accb;
}

__global__ void head_dense_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int Cin, int Cout,
                                               float* __restrict__ dW, float* __restrict__ db) {
  pdl_prologue();
  const int stride = kMaxCo * Cin + kMaxCo;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride) return;
  double acc = 0.0;
  for (int bidx = 0; bidx < nblocks; ++bidx) acc += (double)partial[(size_t)bidx * stride + i];
  if (i < kMaxCo * Cin) {
    if (dW && i < Cout * Cin) dW[i] = (float)acc;
  } else {
    const int co = i - kMaxCo * Cin;
    if (db && co < Cout) db[co] = (float)acc;
  }
}

}  // namespace b2

using namespace b2;

extern "C" long long b2_head_workspace_bytes(int Cin) {
  return (long long)kCeBlocks * (kMaxCo * Cin + kMaxCo + 1) * (long long)sizeof(float) + 64;
}

#define B2_HEAD_CHECK(who)                                                                                 \
  B2_REQUIRE(Cin == 32 || Cin == 64, who ": Cin=%d unsupported (32 or 64)", Cin);                          \
  B2_REQUIRE(Cout >= 1 && Cout <= kMaxCo, who ": Cout=%d unsupported (<= 64)", Cout)

// Fused final_conv + CrossEntropyLoss(ignore_index=-1) + argmax (+ backward when compute_grad).
// loss_out[0] = mean loss over labelled voxels (NaN if none), loss_out[1] = sum; count_out = #labelled.
static int head_ce_impl(const void* x, const long long* labels, long long NV, const float* W, const float* b, int Cin,
                        int Cout, float grad_scale, const float* grad_scale_dev, int compute_grad, int eval_softmax,
                        int* preds, void* dx, float* dW, float* db, float* loss_out, int* count_out, void* workspace,
                        long long workspace_bytes, const void* stat_r, long long* stat_acc, const float* x_scale_shift,
                        int skip_dx_memset, cudaStream_t stream) {
  B2_REQUIRE(x && labels && W && loss_out && count_out && workspace, "b2_head_ce: null pointer");
  B2_HEAD_CHECK("b2_head_ce");
  B2_REQUIRE(workspace_bytes >= b2_head_workspace_bytes(Cin), "b2_head_ce: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  B2_CHECK_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int), stream));
  if (compute_grad && dx && !skip_dx_memset) B2_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)NV * Cin * 2, stream));
  int cblocks = (int)((NV + 255) / 256);
  if (cblocks > num_sms() * 8) cblocks = num_sms() * 8;
  B2_LAUNCH(count_labelled_kernel, cblocks, 256, 0, stream, labels, NV, count_out);
  B2_CHECK_CUDA(cudaGetLastError());
  B2_REQUIRE(NV < (1LL << 31), "b2_head_ce: %lld voxels unsupported", NV);
  const size_t sh = (size_t)(kMaxCo * (Cin + 4) + kMaxCo + kCeGroup * (kMaxCo + 1) + 3 * kCeGroup * (Cin + 4) +
                             kCeGroup) * sizeof(float) +
                    (size_t)(2 * kCeChunk + kCeChunk / 32 + 1) * sizeof(int);
  long long per_block = (NV + kCeBlocks - 1) / kCeBlocks;
  per_block = (per_block + kCeThreads - 1) / kCeThreads * kCeThreads;
  auto* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto* dxb = reinterpret_cast<__nv_bfloat16*>(dx);
  if (Cin == 64) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(head_ce_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    B2_LAUNCH(head_ce_kernel<64>, kCeBlocks, kCeThreads, sh, stream, xb, labels, NV, W, b, Cout, count_out, grad_scale,
                                                              grad_scale_dev, compute_grad, eval_softmax, preds, dxb,
                                                              partial, per_block,
                                                              reinterpret_cast<const __nv_bfloat16*>(stat_r), stat_acc,
                                                              x_scale_shift);
  } else {
    B2_CHECK_CUDA(cudaFuncSetAttribute(head_ce_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    B2_LAUNCH(head_ce_kernel<32>, kCeBlocks, kCeThreads, sh, stream, xb, labels, NV, W, b, Cout, count_out, grad_scale,
                                                              grad_scale_dev, compute_grad, eval_softmax, preds, dxb,
                                                              partial, per_block,
                                                              reinterpret_cast<const __nv_bfloat16*>(stat_r), stat_acc,
                                                              x_scale_shift);
  }
  B2_CHECK_CUDA(cudaGetLastError());
  const int stride = kMaxCo * Cin + kMaxCo + 1;
  B2_LAUNCH(head_ce_finalize_kernel, (stride + 31) / 32, 256, 0, stream, partial, kCeBlocks, Cin, Cout, count_out,
                                                                   compute_grad ? dW : nullptr,
                                                                   compute_grad ? db : nullptr, loss_out);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_head_ce(const void* x, const long long* labels, long long NV, const float* W, const float* b, int Cin,
                          int Cout, float grad_scale, const float* grad_scale_dev, int compute_grad, int eval_softmax,
                          int* preds, void* dx, float* dW, float* db, float* loss_out, int* count_out,
                          void* workspace, long long workspace_bytes, const float* x_scale_shift, cudaStream_t stream) {
  return head_ce_impl(x, labels, NV, W, b, Cin, Cout, grad_scale, grad_scale_dev, compute_grad, eval_softmax, preds, dx,
                      dW, db, loss_out, count_out, workspace, workspace_bytes, nullptr, nullptr, x_scale_shift, 0, stream);
}

// Same with compute_grad and dx: dX is the gradient at the last GroupNorm output, so the kernel also accumulates that
// layer's GroupNorm-backward statistics (sum dX, sum dX*r; r = its saved relu(conv), dense bf16 [NV][Cin]) into
// stat_acc int64 [Cin][4] (see the statistics accumulators above b2_conv3d_igemm_stats in the header).
// skip_dx_memset: dX is NOT cleared — only the rows of labelled voxels are written; the consumer
// (b2_relu_gn_bwd_acc with dy_row_labels = labels) treats every other row as zero without reading it.
extern "C" int b2_head_ce_bstats(const void* x, const long long* labels, long long NV, const float* W, const float* b,
                                 int Cin, int Cout, float grad_scale, const float* grad_scale_dev, int* preds,
                                 void* dx, float* dW, float* db, float* loss_out, int* count_out, void* workspace,
                                 long long workspace_bytes, const void* r, long long* stat_acc,
                                 const float* x_scale_shift, int skip_dx_memset, cudaStream_t stream) {
  B2_REQUIRE(dx && r && stat_acc, "b2_head_ce_bstats: null pointer");
  return head_ce_impl(x, labels, NV, W, b, Cin, Cout, grad_scale, grad_scale_dev, 1, 0, preds, dx, dW, db, loss_out,
                      count_out, workspace, workspace_bytes, r, stat_acc, x_scale_shift, skip_dx_memset, stream);
}

extern "C" int b2_head_gather(const void* x, const long long* index, long long nidx, const float* W, const float* b,
                              int Cin, int Cout, int softmax, float* scores, int* preds, const float* x_scale_shift,
                              cudaStream_t stream) {
  B2_REQUIRE(x && W && scores, "b2_head_gather: null pointer");
  B2_HEAD_CHECK("b2_head_gather");
  if (nidx <= 0) return B2_OK;
  B2_REQUIRE(index, "b2_head_gather: null index");
  const size_t sh = (size_t)(kMaxCo * Cin + kMaxCo) * sizeof(float);
  int blocks = (int)((nidx + 7) / 8);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  auto* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  if (Cin == 64)
    B2_LAUNCH(head_gather_kernel<64>, blocks, 256, sh, stream, xb, index, nidx, W, b, Cout, softmax, scores, preds,
              x_scale_shift);
  else
    B2_LAUNCH(head_gather_kernel<32>, blocks, 256, sh, stream, xb, index, nidx, W, b, Cout, softmax, scores, preds,
              x_scale_shift);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_head_dense_fwd(const void* x, int N, long long V, const float* W, const float* b, int Cin, int Cout,
                                 int softmax, float* out, cudaStream_t stream) {
  B2_REQUIRE(x && W && out, "b2_head_dense_fwd: null pointer");
  B2_HEAD_CHECK("b2_head_dense_fwd");
  long long blocks = ((long long)N * V + 127) / 128;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  auto* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  if (Cin == 64)
    B2_LAUNCH(head_dense_fwd_kernel<64>, (unsigned)blocks, 128, 0, stream, xb, V, N, W, b, Cout, softmax, out);
  else
    B2_LAUNCH(head_dense_fwd_kernel<32>, (unsigned)blocks, 128, 0, stream, xb, V, N, W, b, Cout, softmax, out);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_head_dense_bwd(const float* g, const void* x, int N, long long V, const float* W, int Cin, int Cout,
                                 void* dx, float* dW, float* db, void* workspace, long long workspace_bytes,
                                 cudaStream_t stream) {
  B2_REQUIRE(g && x && W && dx && workspace, "b2_head_dense_bwd: null pointer");
  B2_HEAD_CHECK("b2_head_dense_bwd");
  B2_REQUIRE(workspace_bytes >= b2_head_workspace_bytes(Cin), "b2_head_dense_bwd: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  const size_t sh = (size_t)(kMaxCo * Cin + 128 * kMaxCo + 128 * Cin) * sizeof(float);
  auto* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto* dxb = reinterpret_cast<__nv_bfloat16*>(dx);
  if (Cin == 64) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(head_dense_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    B2_LAUNCH(head_dense_bwd_kernel<64>, kCeBlocks, 128, sh, stream, g, xb, V, N, W, Cout, dxb, partial);
  } else {
    B2_CHECK_CUDA(cudaFuncSetAttribute(head_dense_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    B2_LAUNCH(head_dense_bwd_kernel<32>, kCeBlocks, 128, sh, stream, g, xb, V, N, W, Cout, dxb, partial);
  }
  B2_CHECK_CUDA(cudaGetLastError());
  const int stride = kMaxCo * Cin + kMaxCo;
  B2_LAUNCH(head_dense_bwd_finalize_kernel, (stride + 127) / 128, 128, 0, stream, partial, kCeBlocks, Cin, Cout, dW, db);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
