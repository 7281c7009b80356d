// SGD with momentum (torch.optim.SGD semantics: v <- mu*v + g ; p <- p - lr*v ; dampening 0, no nesterov, wd 0)
// as one multi-tensor launch, plus the bf16 weight re-packs the tcgen05 conv kernels read:
//   fprop pack  Wf[tap][co][ci]      = W[co][ci][tap]
//   dgrad pack  Wd[tap][ci][co]      = W[co][ci][26 - tap]      (flipped taps, transposed channels)
#include "common.h"
#include "ptx.cuh"

namespace b2 {

static constexpr int kMaxTensors = 64;
static constexpr int kChunk = 4096;  // elements per block

struct SgdArgs {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* v[kMaxTensors];
  long long n[kMaxTensors];
  int first_block[kMaxTensors + 1];
  int count;
  float lr, momentum, grad_scale;
};

__global__ void __launch_bounds__(256) sgd_multi_kernel(const SgdArgs a) {
  int t = 0;
  while (t + 1 < a.count && (int)blockIdx.x >= a.first_block[t + 1]) ++t;
  const long long base = (long long)(blockIdx.x - a.first_block[t]) * kChunk;
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ v = a.v[t];
  const long long n = a.n[t];
  for (int i = threadIdx.x; i < kChunk; i += 256) {
    const long long j = base + i;
    if (j < n) {
      const float vv = a.momentum * v[j] + g[j] * a.grad_scale;
      v[j] = vv;
      p[j] -= a.lr * vv;
    }
  }
}

__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                    int Cout, int Cin) {
  const long long total = 27LL * Cout * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i enumerates the fprop pack [tap][co][ci] (coalesced writes)
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int co = (int)(r % Cout);
    const int tap = (int)(r / Cout);
    const float val = w[((long long)co * Cin + ci) * 27 + tap];
    const __nv_bfloat16 b = __float2bfloat16_rn(val);
    if (wf) wf[i] = b;
    if (wd) wd[((long long)(26 - tap) * Cin + ci) * Cout + co] = b;
  }
}

}  // namespace b2

using namespace b2;

// params/grads/moms: arrays of `count` device pointers (host arrays); numels: element counts.
extern "C" int b2_sgd_step(float* const* params, const float* const* grads, float* const* moms,
                           const long long* numels, int count, float lr, float momentum, float grad_scale,
                           cudaStream_t stream) {
  B2_REQUIRE(params && grads && moms && numels && count >= 0, "b2_sgd_step: null pointer");
  int done = 0;
  while (done < count) {
    SgdArgs a;
    int k = 0, blocks = 0;
    while (done + k < count && k < kMaxTensors) {
      const int idx = done + k;
      B2_REQUIRE(params[idx] && grads[idx] && moms[idx], "b2_sgd_step: null tensor %d", idx);
      a.p[k] = params[idx];
      a.g[k] = grads[idx];
      a.v[k] = moms[idx];
      a.n[k] = numels[idx];
      a.first_block[k] = blocks;
      blocks += (int)((numels[idx] + kChunk - 1) / kChunk);
      ++k;
    }
    a.first_block[k] = blocks;
    a.count = k;
    a.lr = lr;
    a.momentum = momentum;
    a.grad_scale = grad_scale;
    if (blocks > 0) {
      sgd_multi_kernel<<<blocks, 256, 0, stream>>>(a);
      B2_CHECK_CUDA(cudaGetLastError());
    }
    done += k;
  }
  return B2_OK;
}

// w: fp32 [Cout][Cin][3][3][3]; wf/wd: bf16 packs (either may be NULL)
extern "C" int b2_pack_conv_weights(const float* w, void* wf, void* wd, int Cout, int Cin, cudaStream_t stream) {
  B2_REQUIRE(w && (wf || wd), "b2_pack_conv_weights: null pointer");
  const long long total = 27LL * Cout * Cin;
  long long blocks = (total + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(wf),
                                                            reinterpret_cast<__nv_bfloat16*>(wd), Cout, Cin);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
