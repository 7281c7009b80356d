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
// partial layout per block: [Cin][kMaxCo] dW (transposed, bank-conflict-free) | [kMaxCo] db | loss
template <int CIN>
__global__ void __launch_bounds__(kCeThreads)
head_ce_kernel(const __nv_bfloat16* __restrict__ x, const long long* __restrict__ labels, long long NV,
               const float* __restrict__ W, const float* __restrict__ b, int Cout, const int* __restrict__ count,
               float grad_scale, const float* __restrict__ grad_scale_dev, int compute_grad, int eval_softmax,
               int* __restrict__ preds, __nv_bfloat16* __restrict__ dx, float* __restrict__ partial) {
  extern __shared__ float shm[];
  float* Wt = shm;                       // [CIN][kMaxCo]
  float* Ws = Wt + CIN * kMaxCo;         // [kMaxCo][CIN]
  float* bs = Ws + kMaxCo * CIN;         // [kMaxCo]
  float* red = bs + kMaxCo;              // [kMaxCo*CIN + kMaxCo + 1]
  load_head_weights(W, b, CIN, Cout, Wt, Ws, bs);
  for (int i = threadIdx.x; i < kMaxCo * CIN + kMaxCo + 1; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int cnt = *count;
  const float gsc = grad_scale_dev ? grad_scale * (*grad_scale_dev) : grad_scale;
  const float gs = (cnt > 0) ? gsc / (float)cnt : 0.f;
  float accW0[CIN], accW1[CIN];
  float accb0 = 0.f, accb1 = 0.f, loss = 0.f;
#pragma unroll
  for (int i = 0; i < CIN; ++i) { accW0[i] = 0.f; accW1[i] = 0.f; }

  const long long gw = (long long)blockIdx.x * nwarp + warp, nw = (long long)gridDim.x * nwarp;
  for (long long base = gw * 32; base < NV; base += nw * 32) {
    const long long v = base + lane;
    const long long lab = (v < NV) ? labels[v] : -1;
    unsigned any = __ballot_sync(0xffffffffu, lab >= 0);
    while (any) {
      const int src = __ffs(any) - 1;
      any &= any - 1;
      const long long vv = base + src;
      const int label = (int)__shfl_sync(0xffffffffu, (int)lab, src);
      uint32_t xp = 0;
      if (lane < CIN / 2) xp = __ldg(reinterpret_cast<const uint32_t*>(x + vv * CIN) + lane);
      WarpHead h = warp_logits<CIN>(xp, Wt, bs, lane);
      float p0, p1, lse;
      warp_softmax(h.l0, h.l1, Cout, lane, p0, p1, lse);
      const int am = warp_argmax(h.l0, h.l1, Cout, lane);
      float lv;
      if (eval_softmax) {  // reference val phase: CrossEntropyLoss applied to Softmax outputs (training.py:189,205-208)
        float q0, q1, lse2;
        warp_softmax(p0, p1, Cout, lane, q0, q1, lse2);
        const float pl = __shfl_sync(0xffffffffu, (label < 32) ? p0 : p1, label & 31);
        lv = lse2 - pl;
      } else {
        const float ll = __shfl_sync(0xffffffffu, (label < 32) ? h.l0 : h.l1, label & 31);
        lv = lse - ll;
      }
      if (lane == 0) {
        loss += lv;
        if (preds) preds[vv] = am;
      }
      if (compute_grad) {
        const float d0 = (lane < Cout) ? (p0 - ((lane == label) ? 1.f : 0.f)) * gs : 0.f;
        const float d1 = (lane + 32 < Cout) ? (p1 - ((lane + 32 == label) ? 1.f : 0.f)) * gs : 0.f;
        accb0 += d0;
        accb1 += d1;
#pragma unroll
        for (int ci = 0; ci < CIN; ci += 2) {
          const uint32_t p = __shfl_sync(0xffffffffu, xp, ci >> 1);
          const float xa = __uint_as_float(p << 16), xb = __uint_as_float(p & 0xffff0000u);
          accW0[ci] = fmaf(d0, xa, accW0[ci]);
          accW1[ci] = fmaf(d1, xa, accW1[ci]);
          accW0[ci + 1] = fmaf(d0, xb, accW0[ci + 1]);
          accW1[ci + 1] = fmaf(d1, xb, accW1[ci + 1]);
        }
        if (dx) {
          float g0 = 0.f, g1 = 0.f;  // dX for channels 2*lane, 2*lane+1
          for (int co = 0; co < Cout; ++co) {
            const float dc = __shfl_sync(0xffffffffu, (co < 32) ? d0 : d1, co & 31);
            if (lane < CIN / 2) {
              const float2 w2 = *reinterpret_cast<const float2*>(Ws + co * CIN + 2 * lane);
              g0 = fmaf(dc, w2.x, g0);
              g1 = fmaf(dc, w2.y, g1);
            }
          }
          if (lane < CIN / 2) {
            __nv_bfloat162 o = __floats2bfloat162_rn(g0, g1);
            reinterpret_cast<__nv_bfloat162*>(dx + vv * CIN)[lane] = o;
          }
        }
      }
    }
  }
  // deterministic block reduction: warps add their registers into smem one after another
  for (int w = 0; w < nwarp; ++w) {
    if (warp == w) {
      if (compute_grad) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {   // [ci][co] layout: lanes hit consecutive banks
          red[ci * kMaxCo + lane] += accW0[ci];
          red[ci * kMaxCo + lane + 32] += accW1[ci];
        }
        red[kMaxCo * CIN + lane] += accb0;
        red[kMaxCo * CIN + lane + 32] += accb1;
      }
      if (lane == 0) red[kMaxCo * CIN + kMaxCo] += loss;
    }
    __syncthreads();
  }
  float* dst = partial + (size_t)blockIdx.x * (kMaxCo * CIN + kMaxCo + 1);
  for (int i = threadIdx.x; i < kMaxCo * CIN + kMaxCo + 1; i += blockDim.x) dst[i] = red[i];
}

__global__ void head_ce_finalize_kernel(const float* __restrict__ partial, int nblocks, int Cin, int Cout,
                                        const int* __restrict__ count, float* __restrict__ dW, float* __restrict__ db,
                                        float* __restrict__ loss_out /*[2]: mean loss, sum*/) {
  const int stride = kMaxCo * Cin + kMaxCo + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride) return;
  double acc = 0.0;
  for (int bidx = 0; bidx < nblocks; ++bidx) acc += (double)partial[(size_t)bidx * stride + i];
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
                   float* __restrict__ scores /*[nidx][Cout]*/, int* __restrict__ preds) {
  extern __shared__ float shm[];
  float* Wt = shm;
  float* bs = Wt + CIN * kMaxCo;
  load_head_weights(W, b, CIN, Cout, Wt, nullptr, bs);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (long long k = (long long)blockIdx.x * nwarp + warp; k < nidx; k += (long long)gridDim.x * nwarp) {
    const long long vv = index[k];
    uint32_t xp = 0;
    if (lane < CIN / 2) xp = __ldg(reinterpret_cast<const uint32_t*>(x + vv * CIN) + lane);
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
extern "C" int b2_head_ce(const void* x, const long long* labels, long long NV, const float* W, const float* b, int Cin,
                          int Cout, float grad_scale, const float* grad_scale_dev, int compute_grad, int eval_softmax,
                          int* preds, void* dx, float* dW, float* db, float* loss_out, int* count_out,
                          void* workspace, long long workspace_bytes, cudaStream_t stream) {
  B2_REQUIRE(x && labels && W && loss_out && count_out && workspace, "b2_head_ce: null pointer");
  B2_HEAD_CHECK("b2_head_ce");
  B2_REQUIRE(workspace_bytes >= b2_head_workspace_bytes(Cin), "b2_head_ce: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  B2_CHECK_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int), stream));
  if (compute_grad && dx) B2_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)NV * Cin * 2, stream));
  int cblocks = (int)((NV + 255) / 256);
  if (cblocks > num_sms() * 8) cblocks = num_sms() * 8;
  count_labelled_kernel<<<cblocks, 256, 0, stream>>>(labels, NV, count_out);
  B2_CHECK_CUDA(cudaGetLastError());
  const size_t sh = (size_t)(2 * kMaxCo * Cin + kMaxCo + kMaxCo * Cin + kMaxCo + 1) * sizeof(float);
  auto* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto* dxb = reinterpret_cast<__nv_bfloat16*>(dx);
  if (Cin == 64) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(head_ce_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    head_ce_kernel<64><<<kCeBlocks, kCeThreads, sh, stream>>>(xb, labels, NV, W, b, Cout, count_out, grad_scale,
                                                              grad_scale_dev, compute_grad, eval_softmax, preds, dxb,
                                                              partial);
  } else {
    B2_CHECK_CUDA(cudaFuncSetAttribute(head_ce_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    head_ce_kernel<32><<<kCeBlocks, kCeThreads, sh, stream>>>(xb, labels, NV, W, b, Cout, count_out, grad_scale,
                                                              grad_scale_dev, compute_grad, eval_softmax, preds, dxb,
                                                              partial);
  }
  B2_CHECK_CUDA(cudaGetLastError());
  const int stride = kMaxCo * Cin + kMaxCo + 1;
  head_ce_finalize_kernel<<<(stride + 127) / 128, 128, 0, stream>>>(partial, kCeBlocks, Cin, Cout, count_out,
                                                                   compute_grad ? dW : nullptr,
                                                                   compute_grad ? db : nullptr, loss_out);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_head_gather(const void* x, const long long* index, long long nidx, const float* W, const float* b,
                              int Cin, int Cout, int softmax, float* scores, int* preds, cudaStream_t stream) {
  B2_REQUIRE(x && W && scores, "b2_head_gather: null pointer");
  B2_HEAD_CHECK("b2_head_gather");
  if (nidx <= 0) return B2_OK;
  B2_REQUIRE(index, "b2_head_gather: null index");
  const size_t sh = (size_t)(kMaxCo * Cin + kMaxCo) * sizeof(float);
  int blocks = (int)((nidx + 7) / 8);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  auto* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  if (Cin == 64)
    head_gather_kernel<64><<<blocks, 256, sh, stream>>>(xb, index, nidx, W, b, Cout, softmax, scores, preds);
  else
    head_gather_kernel<32><<<blocks, 256, sh, stream>>>(xb, index, nidx, W, b, Cout, softmax, scores, preds);
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
    head_dense_fwd_kernel<64><<<(unsigned)blocks, 128, 0, stream>>>(xb, V, N, W, b, Cout, softmax, out);
  else
    head_dense_fwd_kernel<32><<<(unsigned)blocks, 128, 0, stream>>>(xb, V, N, W, b, Cout, softmax, out);
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
    head_dense_bwd_kernel<64><<<kCeBlocks, 128, sh, stream>>>(g, xb, V, N, W, Cout, dxb, partial);
  } else {
    B2_CHECK_CUDA(cudaFuncSetAttribute(head_dense_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    head_dense_bwd_kernel<32><<<kCeBlocks, 128, sh, stream>>>(g, xb, V, N, W, Cout, dxb, partial);
  }
  B2_CHECK_CUDA(cudaGetLastError());
  const int stride = kMaxCo * Cin + kMaxCo;
  head_dense_bwd_finalize_kernel<<<(stride + 127) / 128, 128, 0, stream>>>(partial, kCeBlocks, Cin, Cout, dW, db);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
