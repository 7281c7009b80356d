// Host-side helpers shared by the C-ABI translation units: error slot, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b2 {

// error codes returned by every exported function (0 = ok)
enum : int { B2_OK = 0, B2_ERR_ARG = -1, B2_ERR_CUDA = -2, B2_ERR_UNSUPPORTED = -3, B2_ERR_DRIVER = -4 };

void set_error(const char* fmt, ...);   // thread-local message, read back with b2_last_error()
int check_cuda(cudaError_t e, const char* what);

#define B2_CHECK_CUDA(expr)                                   \
  do {                                                        \
    int _rc = ::b2::check_cuda((expr), #expr);                \
    if (_rc != 0) return _rc;                                 \
  } while (0)

#define B2_REQUIRE(cond, ...)                                 \
  do {                                                        \
    if (!(cond)) {                                            \
      ::b2::set_error(__VA_ARGS__);                           \
      return ::b2::B2_ERR_ARG;                                \
    }                                                         \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box, int swizzle_bytes);

int num_sms();

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A training step is ~130 short kernels in one stream; the dependency gap between two of them (drain, launch, first
// wave) costs microseconds each.  Every kernel of this library starts with pdl_prologue(): griddepcontrol.wait (all
// memory operations of the preceding grid are complete and visible — nothing is read or written before it) followed
// by griddepcontrol.launch_dependents (the NEXT kernel's blocks may be scheduled as soon as every block of this grid
// has started, so its launch latency and prologue overlap this grid's tail).  Launches go through launch_pdl(), which
// sets cudaLaunchAttributeProgrammaticStreamSerialization (captured as programmatic edges in CUDA graphs).
// MEASURED (round 1, 1 x B200, CUDA-graph replay of the training step): 7.37 ms/step with the attribute, 7.19 ms
// without — the early-scheduled dependent blocks cost more than the launch gaps they hide.  The attribute is therefore
// OFF by default (plain stream order; griddepcontrol.* are no-ops then) and B2_PDL=1 in the environment turns it on.
bool pdl_enabled();

#ifdef __CUDACC__
// ---- order-independent fixed-point statistics accumulators ---------------------------------------------------------
// GroupNorm needs per-channel sums over the whole volume; the kernels that PRODUCE a tensor (conv epilogues, pooling /
// upsampling backward, the head) add their per-block fp32 partial sums into a two-limb fixed-point accumulator
// (int64 integer part + int64 fraction in units of 2^-32) with integer atomics.  Integer addition is associative, so
// the total is bit-identical whatever the arrival order (no float atomics, run-to-run deterministic).  An fp32 value
// splits into rintf(p) and a fraction that is kept to 2^-32: exact for |p| >= 2^-9, within 2^-33 absolute otherwise.
// A one-block kernel (gn_finalize_acc / gn_bwd_finalize_acc, norm.cu) turns the accumulators into mean, rstd and the
// backward coefficients: no per-block partial buffers, no "last block" pass, no statistics pass over the tensor.
// Layout per channel: [4] = {sum_hi, sum_lo, sq_hi, sq_lo}.
// Non-finite contributions (a diverged run) cannot be represented in fixed point: they POISON the slot instead (the
// fraction limb is raised to >= 2^60, far above anything finite contributions can accumulate: < 2^31 each), and
// stat_read() then returns NaN, so the finalize kernels propagate NaN like a floating-point reduction would.
static constexpr long long kStatPoison = 1ll << 60;
__device__ __forceinline__ void stat_atomic_add(long long* slot, float p) {
  if (!isfinite(p)) {
    atomicMax(slot + 1, kStatPoison);
    return;
  }
  const float h = rintf(p);
  const long long hi = (long long)h;
  const long long lo = __float2ll_rn((p - h) * 4294967296.f);
  if (hi != 0) atomicAdd(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)hi);
  if (lo != 0) atomicAdd(reinterpret_cast<unsigned long long*>(slot + 1), (unsigned long long)lo);
}
__device__ __forceinline__ double stat_read(const long long* slot) {
  const long long lo = __ldcg(slot + 1);
  if (lo >= (kStatPoison >> 1)) return __longlong_as_double(0x7ff8000000000000ll);
  return (double)__ldcg(slot) + (double)lo * (1.0 / 4294967296.0);
}

// pdl_wait: nothing of the preceding grid is read or written before it.  pdl_trigger: the NEXT kernel may be staged.
// Round 1 triggered at the very start of every block: the dependent grid was scheduled (and sat waiting on SM slots)
// for the whole duration of the primary — measured slower.  Round 2 triggers late: generic kernels only from the blocks
// of (roughly) their LAST resident wave (earlier blocks count as triggered when they exit), the persistent convolution
// kernels when a CTA's producers have issued their last load.  MEASURED (round 2, 1 x B200, two A/B pairs of 40-step
// graph replays): 6.80 / 6.82 ms with the attribute against 6.68 / 6.70 ms without — still a loss, so B2_PDL stays
// opt-in.  No-ops unless the launch carries the PDL attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_wait();
  const unsigned lin = blockIdx.x + blockIdx.y * gridDim.x, total = gridDim.x * gridDim.y;
  if (lin + 592u >= total) pdl_trigger();   // 592 = 148 SMs x 4 resident blocks: the last wave of a multi-wave grid
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define B2_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (void)::b2::launch_pdl(::b2::pdl_enabled(), kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
// Selective programmatic dependent launch (round 2): ONLY around the one-block statistics finalizes.  The finalize is
// scheduled while its producer drains and the apply kernel behind it while the (one-block) finalize runs, so both
// launch latencies leave the critical path.  MEASURED (1 x B200, 40-step graph replays, two A/B pairs): +0.5 % and
// -0.7 %, i.e. within run-to-run noise -> opt-in (B2_PDL_FINALIZE=1).
bool pdl_finalize_enabled();
#define B2_LAUNCH_DEP(kernel, grid, block, smem, stream, ...)                                                        \
  (void)::b2::launch_pdl(::b2::pdl_enabled() || ::b2::pdl_finalize_enabled(), kernel, dim3(grid), dim3(block),     \
                         (size_t)(smem), stream, __VA_ARGS__)
#endif

}  // namespace b2
