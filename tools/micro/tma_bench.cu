// Microbenchmark: per-SM TMA load throughput for 2-D vs 5-D boxes made of 128-byte rows (round-1 design question:
// why is conv3d_igemm_kernel stage time ~900 cycles regardless of stage bytes?).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../2022_pauriau_unetsulc_b200/csrc tma_bench.cu ../../2022_pauriau_unetsulc_b200/csrc/common.cu -o tma_bench
#include "common.h"
#include "ptx.cuh"
#include <vector>
using namespace b2;

struct P { int mode; int iters; int stages; int box_bytes; int bw, bh, bd; int W, H, D; int tiles_w, tiles_h, tiles_d; int rows2d; int nboxes; };

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, P p, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * p.box_bytes * p.nboxes);
  if (threadIdx.x == 0) { for (int s = 0; s < p.stages; ++s) mbar_init(&full[s], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t phase_done = 0; int sd = 0;
    // keep `stages` loads in flight
    for (int it = 0; it < p.iters + p.stages; ++it) {
      if (it >= p.stages) {  // wait oldest
        mbar_wait(&full[sd], phase_done);
        if (++sd == p.stages) { sd = 0; phase_done ^= 1; }
        ++done;
      }
      if (issued < p.iters) {
        int s = issued % p.stages;
        mbar_arrive_expect_tx(&full[s], p.box_bytes * p.nboxes);
        long long t = (long long)blockIdx.x * p.iters + issued;
        for (int b = 0; b < p.nboxes; ++b) {
          uint8_t* dst = smem + ((size_t)s * p.nboxes + b) * p.box_bytes;
          if (p.mode == 2) {
            int r0 = (int)((t * 131 + b * 977) % (p.rows2d - 256));
            tma_load_2d(dst, &tm, &full[s], 0, r0);
          } else {
            long long mt = t * 7 + b;
            int tw = (int)(mt % p.tiles_w); mt /= p.tiles_w;
            int th = (int)(mt % p.tiles_h); mt /= p.tiles_h;
            int td = (int)(mt % p.tiles_d);
            tma_load_5d(dst, &tm, &full[s], 0, tw * p.bw - 1, th * p.bh, td * p.bd + 1, 0);
          }
        }
        ++issued;
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int W = 96, H = 112, D = 96, C = 64;
  size_t n = (size_t)W * H * D * C;
  __nv_bfloat16* x; cudaMalloc(&x, n * 2); cudaMemset(x, 0, n * 2);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { const char* name; int mode; int bw, bh, bd; int nboxes; int ld; };
  Cfg cfgs[] = {{"2D {64,128}", 2, 0, 0, 0, 1, 64}, {"2D {64,128} x2 boxes", 2, 0, 0, 0, 2, 64},
                {"5D {64,32,4,1}", 5, 32, 4, 1, 1, 64}, {"5D {64,8,4,4}", 5, 8, 4, 4, 1, 64}, {"5D {64,128,1,1}", 5, 128, 1, 1, 1, 64},
                {"5D {64,32,4,1} x2", 5, 32, 4, 1, 2, 64}, {"5D {64,16,8,1}", 5, 16, 8, 1, 1, 64}, {"5D {64,4,4,8}", 5, 4, 4, 8, 1, 64}};
  for (auto& c : cfgs) {
    CUtensorMap tm;
    P p{}; p.mode = c.mode; p.iters = 2000; p.stages = 8; p.box_bytes = 16384; p.nboxes = c.nboxes;
    if (c.nboxes * p.stages * 16384 > 200 * 1024) p.stages = 6;
    if (c.mode == 2) {
      uint64_t dims[2] = {64, (uint64_t)W * H * D}; uint64_t str[1] = {128}; uint32_t box[2] = {64, 128};
      if (encode_tmap_bf16(&tm, x, 2, dims, str, box, 128)) { printf("encode fail\n"); return 1; }
      p.rows2d = W * H * D;
    } else {
      uint64_t dims[5] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)D, 1};
      uint64_t str[4] = {128, 128ull * W, 128ull * W * H, 128ull * W * H * D};
      uint32_t box[5] = {64, (uint32_t)c.bw, (uint32_t)c.bh, (uint32_t)c.bd, 1};
      if (encode_tmap_bf16(&tm, x, 5, dims, str, box, 128)) { printf("encode fail\n"); return 1; }
      p.bw = c.bw; p.bh = c.bh; p.bd = c.bd; p.W = W; p.H = H; p.D = D;
      p.tiles_w = W / c.bw ? W / c.bw : 1; p.tiles_h = H / c.bh; p.tiles_d = (D - 2) / c.bd;
    }
    size_t sh = (size_t)p.stages * p.box_bytes * p.nboxes + 1024 + 256;
    for (int grid : {1, 148}) {
      k<<<grid, 64, sh>>>(tm, p, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : h) avg += v; avg /= grid;
      double per = avg / p.iters;
      printf("%-24s grid %3d stages %d: %.0f cycles per iteration (%d x 16 KB) = %.1f B/clk/SM, %.2f cyc per 128B row\n",
             c.name, grid, p.stages, per, c.nboxes, 16384.0 * c.nboxes / per, per / (128.0 * c.nboxes));
    }
  }
  return 0;
}
