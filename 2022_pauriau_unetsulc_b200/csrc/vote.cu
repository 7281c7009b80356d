// Post-inference integer pass ("cutting", reference pattern_class.py:229-231): per elementary fold, histogram of
// voxel arg-max labels -> top-1 / top-2 -> cut iff count(top-2) > threshold -> one label per (sub)fold.
// Bit-exact against oracle/cutting_ref.py on identical fp32 scores (integer counts; float work is comparisons only).
// Also: per-class TP/FP/FN counters for esi_score (reference training.py:222-225).
#include "common.h"

namespace b2 {

// one thread per voxel: arg-max over C scores (ties -> lowest index), histogram into hist[fold][C]
__global__ void __launch_bounds__(256)
vote_argmax_hist_kernel(const float* __restrict__ scores, const int* __restrict__ fold, long long n, int C,
                        int* __restrict__ lab, int* __restrict__ hist) {
  pdl_prologue();
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    const float* s = scores + v * C;
    float best = s[0];
    int bi = 0;
    for (int c = 1; c < C; ++c) {
      const float x = s[c];
      if (x > best) { best = x; bi = c; }
    }
    lab[v] = bi;
    atomicAdd(&hist[(long long)fold[v] * C + bi], 1);
  }
}

// one thread per (fold, threshold): decision[th][fold] = {l1, l2, cut}
__global__ void __launch_bounds__(128)
vote_decide_kernel(const int* __restrict__ hist, int F, int C, const int* __restrict__ thresholds, int T,
                   int4* __restrict__ decision) {
  pdl_prologue();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int* h = hist + (long long)f * C;
  int l1 = 0, n1 = h[0];
  for (int c = 1; c < C; ++c)
    if (h[c] > n1) { n1 = h[c]; l1 = c; }
  int l2 = -1, n2 = -1;
  for (int c = 0; c < C; ++c)
    if (c != l1 && h[c] > n2) { n2 = h[c]; l2 = c; }
  for (int t = 0; t < T; ++t) {
    const int cut = (l2 >= 0 && n2 > thresholds[t]) ? 1 : 0;
    decision[(long long)t * F + f] = make_int4(l1, l2, cut, n2);
  }
}

__global__ void __launch_bounds__(256)
vote_assign_kernel(const float* __restrict__ scores, const int* __restrict__ fold, long long n, int C, int F, int T,
                   const int4* __restrict__ decision, int* __restrict__ out /*[T][n]*/) {
  pdl_prologue();
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    const int f = fold[v];
    for (int t = 0; t < T; ++t) {
      const int4 d = decision[(long long)t * F + f];
      int o = d.x;
      if (d.z) {
        const float s1 = scores[v * C + d.x], s2 = scores[v * C + d.y];
        o = (s1 >= s2) ? d.x : d.y;
      }
      out[(long long)t * n + v] = o;
    }
  }
}

__global__ void __launch_bounds__(256)
esi_counts_kernel(const int* __restrict__ y_true, const int* __restrict__ y_pred, long long n, int C,
                  unsigned long long* __restrict__ counts /*[3][C]: TP, FP, FN*/) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = y_true[i], p = y_pred[i];
    if (t == p) {
      if (t >= 0 && t < C) atomicAdd(&counts[t], 1ull);
    } else {
      if (p >= 0 && p < C) atomicAdd(&counts[C + p], 1ull);
      if (t >= 0 && t < C) atomicAdd(&counts[2 * C + t], 1ull);
    }
  }
}

// Per-step epoch metrics of the training loop (reference training.py:215-225), kept on the device: TP/FP/FN counters
// from the int64 label volume and the int32 predictions of the head kernel (-1 where unlabelled), plus the running
// loss sum.  One launch per step, capturable in the step's CUDA graph; the host reads the totals once per phase.
__global__ void __launch_bounds__(256)
step_metrics_kernel(const long long* __restrict__ labels, const int* __restrict__ preds, long long n, int C,
                    unsigned long long* __restrict__ counts, const float* __restrict__ loss, double loss_weight,
                    double* __restrict__ loss_acc) {
  pdl_prologue();
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss != nullptr) {   // the only writer; launches are stream-ordered
    loss_acc[0] += (double)loss[0] * loss_weight;
    loss_acc[1] += loss_weight;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = labels[i];
    if (t < 0) continue;                       // unlabelled voxel: the reference masks it out (labels != -1)
    const int p = preds[i];
    if (t == (long long)p) {
      if (t < C) atomicAdd(&counts[t], 1ull);
    } else {
      if (p >= 0 && p < C) atomicAdd(&counts[C + p], 1ull);
      if (t < C) atomicAdd(&counts[2 * C + t], 1ull);
    }
  }
}

}  // namespace b2

using namespace b2;

extern "C" long long b2_fold_vote_workspace_bytes(long long n, int C, int F, int T) {
  return ((long long)F * C + n) * (long long)sizeof(int) + (long long)T * F * (long long)sizeof(int4) + 64;
}

// scores fp32 [n][C]; fold int32 [n] dense ids in [0,F); thresholds: DEVICE int32 [T]; out int32 [T][n]
extern "C" int b2_fold_vote(const float* scores, const int* fold, long long n, int C, int F, const int* thresholds,
                            int T, int* out, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  if (n == 0) return B2_OK;
  B2_REQUIRE(scores && fold && thresholds && out && workspace, "b2_fold_vote: null pointer");
  B2_REQUIRE(n > 0 && C > 0 && F > 0 && T > 0, "b2_fold_vote: bad sizes");
  B2_REQUIRE(workspace_bytes >= b2_fold_vote_workspace_bytes(n, C, F, T), "b2_fold_vote: workspace too small");
  // layout: decision (16B aligned) | hist | lab
  int4* decision = reinterpret_cast<int4*>(workspace);
  int* hist = reinterpret_cast<int*>(decision + (size_t)T * F);
  int* lab = hist + (size_t)F * C;
  B2_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)F * C * sizeof(int), stream));
  int blocks = (int)((n + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  B2_LAUNCH(vote_argmax_hist_kernel, blocks, 256, 0, stream, scores, fold, n, C, lab, hist);
  B2_CHECK_CUDA(cudaGetLastError());
  B2_LAUNCH(vote_decide_kernel, (F + 127) / 128, 128, 0, stream, hist, F, C, thresholds, T, decision);
  B2_CHECK_CUDA(cudaGetLastError());
  B2_LAUNCH(vote_assign_kernel, blocks, 256, 0, stream, scores, fold, n, C, F, T, decision, out);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

// counts: uint64 [3][C] (TP, FP, FN), accumulated (caller zeroes)
extern "C" int b2_esi_counts(const int* y_true, const int* y_pred, long long n, int C, unsigned long long* counts,
                             cudaStream_t stream) {
  if (n == 0) return B2_OK;
  B2_REQUIRE(y_true && y_pred && counts && C > 0, "b2_esi_counts: null pointer");
  int blocks = (int)((n + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  B2_LAUNCH(esi_counts_kernel, blocks, 256, 0, stream, y_true, y_pred, n, C, counts);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}

// labels int64 [n] (-1 = unlabelled), preds int32 [n]; counts uint64 [3][C] accumulated; loss (may be NULL): fp32 device
// scalar of this step, loss_acc double [2] += {loss * loss_weight, loss_weight}
extern "C" int b2_step_metrics(const long long* labels, const int* preds, long long n, int C,
                               unsigned long long* counts, const float* loss, double loss_weight, double* loss_acc,
                               cudaStream_t stream) {
  B2_REQUIRE(labels && preds && counts && C > 0 && n >= 0, "b2_step_metrics: null pointer");
  B2_REQUIRE(loss == nullptr || loss_acc != nullptr, "b2_step_metrics: loss without loss_acc");
  int blocks = (int)((n + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (blocks < 1) blocks = 1;
  B2_LAUNCH(step_metrics_kernel, blocks, 256, 0, stream, labels, preds, n, C, counts, loss, loss_weight, loss_acc);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
