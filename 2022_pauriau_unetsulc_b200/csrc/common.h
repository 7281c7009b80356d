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

}  // namespace b2
