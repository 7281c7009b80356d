// test_thresholds() voxel matching on the device (reference pattern_class.py:205-228): the voxels of
// the cut graph and of the not-cut graph are the same set listed in different orders; the reference sorts both lists
// by native (x, y, z) with pandas and zips them to attach the not-cut graph's vertex id (= elementary fold) to every
// voxel of the cut graph.  Here: one 64-bit key per voxel (21 bits per biased coordinate), a stable LSD radix sort of
// (key, index) pairs for each list (cub::DeviceRadixSort — the CUDA toolkit's sorting primitive), and a gather that
// pairs equal ranks.  Stable = ties keep list order, like the reference's lexicographic sort.
#include "common.h"
#include <cub/device/device_radix_sort.cuh>

namespace b2 {

static constexpr int kCoordBias = 1 << 20;   // coordinates in (-2^20, 2^20)

__global__ void __launch_bounds__(256)
match_keys_kernel(const int* __restrict__ pts, int n, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long x = (unsigned long long)((pts[3 * i] + kCoordBias) & 0x1fffff);
  const unsigned long long y = (unsigned long long)((pts[3 * i + 1] + kCoordBias) & 0x1fffff);
  const unsigned long long z = (unsigned long long)((pts[3 * i + 2] + kCoordBias) & 0x1fffff);
  keys[i] = (x << 42) | (y << 21) | z;
  idx[i] = i;
}

__global__ void __launch_bounds__(256)
match_gather_kernel(const int* __restrict__ order_a, const int* __restrict__ order_b, const int* __restrict__ val_b,
                    int n, int* __restrict__ out_a) {
  pdl_prologue();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out_a[order_a[k]] = val_b[order_b[k]];
}

static size_t sort_temp_bytes(int n) {
  size_t bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr,
                                                  (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, n,
                                                  0, 63);
  if (e != cudaSuccess) {            // no device (build container): conservative bound, re-checked at call time
    (void)cudaGetLastError();
    bytes = (size_t)n * 16 + (1u << 20);
  }
  return bytes;
}
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace b2

using namespace b2;

extern "C" long long b2_match_voxels_workspace_bytes(int n) {
  if (n < 0) return -1;
  const size_t m = (size_t)(n > 0 ? n : 1);
  return (long long)(2 * align256(m * 8) + 3 * align256(m * 4) + align256(sort_temp_bytes((int)m)) + 256);
}

// pts_a, pts_b: int32 [n][3] native voxel coordinates of the same voxel set in two orders; val_b int32 [n];
// out_a[i] = val_b[j] where voxel i of list a and voxel j of list b have equal rank in the (x, y, z)-sorted lists.
extern "C" int b2_match_voxels(const int* pts_a, const int* pts_b, const int* val_b, int n, int* out_a,
                               void* workspace, long long workspace_bytes, cudaStream_t stream) {
  if (n == 0) return B2_OK;
  B2_REQUIRE(pts_a && pts_b && val_b && out_a && workspace && n > 0, "b2_match_voxels: null pointer");
  const size_t temp = sort_temp_bytes(n);
  const size_t k8 = align256((size_t)n * 8), k4 = align256((size_t)n * 4);
  B2_REQUIRE((size_t)workspace_bytes >= 2 * k8 + 3 * k4 + align256(temp), "b2_match_voxels: workspace too small");
  uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(w);
  unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(w + k8);
  int* idx_in = reinterpret_cast<int*>(w + 2 * k8);
  int* order_a = reinterpret_cast<int*>(w + 2 * k8 + k4);
  int* order_b = reinterpret_cast<int*>(w + 2 * k8 + 2 * k4);
  void* tmp = w + 2 * k8 + 3 * k4;
  size_t tb = temp;
  const int blocks = (n + 255) / 256;
  for (int pass = 0; pass < 2; ++pass) {
    B2_LAUNCH(match_keys_kernel, blocks, 256, 0, stream, pass == 0 ? pts_a : pts_b, n, keys_in, idx_in);
    B2_CHECK_CUDA(cudaGetLastError());
    B2_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const unsigned long long*)keys_in, keys_out,
                                                  (const int*)idx_in, pass == 0 ? order_a : order_b, n, 0, 63,
                                                  stream));
  }
  B2_LAUNCH(match_gather_kernel, blocks, 256, 0, stream, static_cast<const int*>(order_a),
            static_cast<const int*>(order_b), val_b, n, out_a);
  B2_CHECK_CUDA(cudaGetLastError());
  return B2_OK;
}
