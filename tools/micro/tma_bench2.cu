// Microbenchmark 2: what limits per-SM TMA throughput?  Variants: rows per box, L2 promotion, number of issuing
// warps (independent rings), hot (L2-resident) vs streaming addresses, and an LDGSTS (cp.async 16 B) producer.
#include "common.h"
#include "ptx.cuh"
#include <vector>
using namespace b2;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct P { int iters, stages, box_bytes, rows, hot, nwarps; long long total_rows; int five; };

__global__ void __launch_bounds__(256, 1) k_tma(const __grid_constant__ CUtensorMap tm, P p, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.nwarps * p.stages * p.box_bytes);
  uint64_t* full = bars + warp * p.stages;
  if (lane == 0 && warp < p.nwarps) { for (int s = 0; s < p.stages; ++s) mbar_init(&full[s], 1); fence_mbar_init(); }
  __syncthreads();
  if (lane == 0 && warp < p.nwarps) {
    uint8_t* base = smem + (size_t)warp * p.stages * p.box_bytes;
    long long t0 = clock64();
    int issued = 0; uint32_t ph = 0; int sd = 0;
    for (int it = 0; it < p.iters + p.stages; ++it) {
      if (it >= p.stages) { mbar_wait(&full[sd], ph); if (++sd == p.stages) { sd = 0; ph ^= 1; } }
      if (issued < p.iters) {
        int s = issued % p.stages;
        mbar_arrive_expect_tx(&full[s], p.box_bytes);
        long long t = ((long long)blockIdx.x * p.nwarps + warp) * p.iters + issued;
        long long r0 = p.hot ? (long long)(blockIdx.x * 8 + warp) * 256 : (t * p.rows) % (p.total_rows - p.rows);
        if (p.five) {
          long long v = p.hot ? (long long)(blockIdx.x * 8 + warp) * 4 : t;
          int tw = (int)(v % 3); int th = (int)((v / 3) % 28); int td = (int)((v / 84) % 90);
          tma_load_5d(base + (size_t)s * p.box_bytes, &tm, &full[s], 0, tw * 32 - 1, th * 4 + 1, td + 1, 0);
        } else {
          tma_load_2d(base + (size_t)s * p.box_bytes, &tm, &full[s], 0, (int)r0);
        }
        ++issued;
      }
    }
    if (warp == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

// LDGSTS producer: 128 threads, each 16 B per op; 16 KB tile = 8 ops per thread; ring of `stages` tiles
__global__ void __launch_bounds__(128, 1) k_ldgsts(const uint8_t* __restrict__ x, P p, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  long long t0 = clock64();
  for (int it = 0; it < p.iters; ++it) {
    int s = it % p.stages;
    long long t = (long long)blockIdx.x * p.iters + it;
    long long r0 = p.hot ? (long long)blockIdx.x * 256 : (t * 128) % (p.total_rows - 128);
    const uint8_t* src = x + r0 * 128;
    uint8_t* dst = smem + (size_t)s * 16384;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int e = j * 128 + threadIdx.x;  // 16-byte element index within the tile
      int row = e >> 3, ch = e & 7;
      uint32_t d = smem_u32(dst + row * 128 + ((ch ^ (row & 7)) << 4));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (size_t)e * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (it >= p.stages - 1) asm volatile("cp.async.wait_group %0;" ::"n"(7) : "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const long long rows = 96LL * 112 * 96;
  uint8_t* x; cudaMalloc(&x, rows * 128); cudaMemset(x, 0, rows * 128);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  struct Cfg { const char* name; int box_rows; int stages; int nwarps; int hot; CUtensorMapL2promotion l2; };
  Cfg cfgs[] = {
      {"rows128 st8 w1 stream L2_256", 128, 8, 1, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"rows128 st8 w1 stream L2_128", 128, 8, 1, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"rows128 st8 w1 stream L2_none", 128, 8, 1, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"rows128 st8 w1 HOT", 128, 8, 1, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"rows256 st4 w1 stream", 256, 4, 1, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"rows64 st16 w1 stream", 64, 16, 1, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"rows128 st4 w2 stream", 128, 4, 2, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"rows128 st3 w4 stream", 128, 3, 4, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"rows64 st3 w8 stream", 64, 3, 8, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"rows128 st3 w4 HOT", 128, 3, 4, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
  };
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode fail %d\n", (int)r); return 1; }
    P p{}; p.iters = 2000; p.stages = c.stages; p.box_bytes = c.box_rows * 128; p.rows = c.box_rows; p.hot = c.hot;
    p.nwarps = c.nwarps; p.total_rows = rows;
    size_t sh = (size_t)c.nwarps * c.stages * p.box_bytes + 2048;
    for (int grid : {1, 148}) {
      k_tma<<<grid, 256, sh>>>(tm, p, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : h) avg += v; avg /= grid;
      double bytes = (double)p.iters * p.box_bytes * c.nwarps;
      printf("TMA %-32s grid %3d: %.1f B/clk/SM (%.2f cyc per 128B row)\n", c.name, grid, bytes / avg, avg / (bytes / 128));
    }
  }
  {
    struct C5 { const char* name; int stages, nwarps, hot; };
    C5 c5[] = {{"5D{64,32,4,1} st8 w1", 8, 1, 0}, {"5D{64,32,4,1} st4 w2", 4, 2, 0}, {"5D{64,32,4,1} st3 w4", 3, 4, 0},
               {"5D{64,32,4,1} st3 w4 HOT", 3, 4, 1}, {"5D{64,32,4,1} st1 w8", 1, 8, 0}};
    for (auto& c : c5) {
      CUtensorMap tm;
      cuuint64_t dims[5] = {64, 96, 112, 96, 1}; cuuint64_t str[4] = {128, 128ull * 96, 128ull * 96 * 112, 128ull * 96 * 112 * 96};
      cuuint32_t box[5] = {64, 32, 4, 1, 1}; cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode5 fail %d\n", (int)r); return 1; }
      P p{}; p.iters = 2000; p.stages = c.stages; p.box_bytes = 16384; p.rows = 128; p.hot = c.hot; p.nwarps = c.nwarps;
      p.total_rows = rows; p.five = 1;
      size_t sh = (size_t)c.nwarps * c.stages * p.box_bytes + 2048;
      for (int grid : {1, 148}) {
        k_tma<<<grid, 256, sh>>>(tm, p, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto v : h) avg += v; avg /= grid;
        double bytes = (double)p.iters * p.box_bytes * c.nwarps;
        printf("TMA %-32s grid %3d: %.1f B/clk/SM (%.2f cyc per 128B row)\n", c.name, grid, bytes / avg, avg / (bytes / 128));
      }
    }
  }
  for (int hot : {0, 1})
    for (int grid : {1, 148}) {
      P p{}; p.iters = 4000; p.stages = 8; p.hot = hot; p.total_rows = rows;
      k_ldgsts<<<grid, 128, 8 * 16384 + 1024>>>(x, p, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ldgsts: %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : h) avg += v; avg /= grid;
      printf("LDGSTS 128 threads x16B st8 %s grid %3d: %.1f B/clk/SM\n", hot ? "HOT" : "stream", grid, 4000.0 * 16384 / avg);
    }
  return 0;
}
