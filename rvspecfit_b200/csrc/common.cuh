// Shared device/host helpers for librvs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rvs_b200.h"

#define RVS_C_KMS 299792.458  // reference spec_fit.py:23

namespace rvs {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// Raise (never lower) the dynamic shared memory limit of a kernel to at least `smem`
// bytes.  The limit is per-function state of the device: concurrent host threads that set
// it to the size of their own launch lower it under each other's feet ("too many
// resources requested for launch"); here it only grows, under a lock.
cudaError_t ensure_dyn_smem_ptr(const void *kern, size_t smem);
template <typename K>
inline cudaError_t ensure_dyn_smem(K kern, size_t smem) {
  return ensure_dyn_smem_ptr(reinterpret_cast<const void *>(kern), smem);
}

// Optional per-stage timing (rvs_profile_enable): CUDA events bracket each kernel
// of the fused evaluation on its launching stream; rvs_profile_read sums them.
enum Stage { ST_LOCATE = 0, ST_PREP, ST_CHUNK, ST_GRAM, ST_SOLVE, ST_RESID, ST_COUNT };
bool prof_on();
void prof_begin(int stage, cudaStream_t st);
void prof_end(int stage, cudaStream_t st);

#define RVS_CUDA_OK(call)                                                         \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      rvs::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call,                   \
                     cudaGetErrorString(e__));                                    \
      return RVS_E_CUDA;                                                          \
    }                                                                             \
  } while (0)

#define RVS_LAUNCH_OK()                                                           \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      rvs::set_error("%s:%d launch: %s", __FILE__, __LINE__,                      \
                     cudaGetErrorString(e__));                                    \
      return RVS_E_CUDA;                                                          \
    }                                                                             \
    rvs::count_launch();                                                          \
  } while (0)

#define RVS_REQUIRE(cond, code, ...)                                              \
  do {                                                                            \
    if (!(cond)) {                                                                \
      rvs::set_error(__VA_ARGS__);                                                \
      return code;                                                                \
    }                                                                             \
  } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// read-only, L1-allocating global loads
template <typename T>
__device__ __forceinline__ T ldg(const T *p) {
  return __ldg(p);
}

// streaming 16-byte load that does not pollute L1 (template grid rows are read
// once per CTA)
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

}  // namespace rvs
