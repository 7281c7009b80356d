// SulciDataset.__getitem__ on the device (reference dataset.py:66-88): the point list of a subject
// ("bucket", after the host-side rotation augmentation and integer cast) is scattered into the dense network input
// x fp32 [D,H,W] (1 at the skeleton voxels) and the label volume int64 [D,H,W] (background elsewhere).
// The reference builds both on the CPU with index_put: when several points fall on one voxel (rotation + truncation)
// the LAST point of the list wins.  Here: pass 1 records per voxel the largest point index (integer atomicMax:
// order-independent), pass 2 lets exactly that point write — bit-identical volumes, and only 16 bytes per point cross
// PCIe instead of 12 bytes per voxel.
#include "common.h"

namespace b2 {

__global__ void __launch_bounds__(256)
scatter_fill_kernel(float* __restrict__ x, long long* __restrict__ labels, int* __restrict__ winner, long long V,
                    long long background) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
    x[i] = 0.f;
    labels[i] = background;
    winner[i] = -1;
  }
}

__device__ __forceinline__ long long point_voxel(const int* __restrict__ pts, int i, int D, int H, int W) {
  const int a = pts[3 * i], b = pts[3 * i + 1], c = pts[3 * i + 2];
  if ((unsigned)a >= (unsigned)D || (unsigned)b >= (unsigned)H || (unsigned)c >= (unsigned)W) return -1;
  return ((long long)a * H + b) * W + c;
}

__global__ void __launch_bounds__(256)
scatter_claim_kernel(const int* __restrict__ pts, int n, int D, int H, int W, int* __restrict__ winner) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long v = point_voxel(pts, i, D, H, W);
  if (v >= 0) atomicMax(winner + v, i);
}

__global__ void __launch_bounds__(256)
scatter_write_kernel(const int* __restrict__ pts, const int* __restrict__ point_labels, int n, int D, int H, int W,
                     const int* __restrict__ winner, float* __restrict__ x, long long* __restrict__ labels) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long v = point_voxel(pts, i, D, H, W);
  if (v >= 0 && winner[v] == i) {
    x[v] = 1.f;
    labels[v] = (long long)point_labels[i];
  }
}

}  // namespace b2

using namespace b2;

extern "C" long long b2_scatter_volume_workspace_bytes(int D, int H, int W) {
  return (long long)D * H * W * (long long)sizeof(int);
}

// pts int32 [n][3] (voxel coordinates, first index = slowest dimension of the volume), point_labels int32 [n],
// x fp32 [D][H][W], labels int64 [D][H][W]; points outside the volume are ignored (the host wrapper rejects them like
// index_put would).  workspace: b2_scatter_volume_workspace_bytes.
extern "C" int b2_scatter_volume(const int* pts, const int* point_labels, int n, int D, int H, int W, float* x,
                                 long long* labels, long long background, void* workspace, long long workspace_bytes,
                                 cudaStream_t stream) {
  B2_REQUIRE(x && labels && workspace && (n == 0 || (pts && point_labels)), "b2_scatter_volume: null pointer");
  B2_REQUIRE(D > 0 && H > 0 && W > 0 && n >= 0, "b2_scatter_volume: bad shape");
  B2_REQUIRE(workspace_bytes >= b2_scatter_volume_workspace_bytes(D, H, W), "b2_scatter_volume: workspace too small");
  const long long V = (long long)D * H * W;
  int* winner = reinterpret_cast<int*>(workspace);
  long long fb = (V + 255) / 256;
  if (fb > num_sms() * 16) fb = num_sms() * 16;
  B2_LAUNCH(scatter_fill_kernel, (unsigned)fb, 256, 0, stream, x, labels, winner, V, background);
  B2_CHECK_CUDA(cudaGetLastError());
  if (n > 0) {
    B2_LAUNCH(scatter_claim_kernel, (n + 255) / 256, 256, 0, stream, pts, n, D, H, W, winner);
    B2_CHECK_CUDA(cudaGetLastError());
    B2_LAUNCH(scatter_write_kernel, (n + 255) / 256, 256, 0, stream, pts, point_labels, n, D, H, W,
              static_cast<const int*>(winner), x, labels);
    B2_CHECK_CUDA(cudaGetLastError());
  }
  return B2_OK;
}
