// Shared device/host helpers for the DG-SCT adapter engine (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dgsct.h"

namespace dgsct {

typedef __nv_bfloat16 bf16;

// thread-local text of the last CUDA error (exposed through dgsct_last_cuda_error)
void set_cuda_error(cudaError_t e, const char* where);

#define DGSCT_CUDA_CHECK(expr)                                   \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) {                                     \
      ::dgsct::set_cuda_error(_e, #expr);                        \
      return DGSCT_ERR_CUDA;                                     \
    }                                                            \
  } while (0)

#define DGSCT_LAUNCH_CHECK(name)                                 \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) {                                     \
      ::dgsct::set_cuda_error(_e, name);                         \
      return DGSCT_ERR_CUDA;                                     \
    }                                                            \
  } while (0)

// cudaFuncSetAttribute is per DEVICE: one flag per device ordinal (nn.DataParallel drives several devices from one process).
// Returns true the first time it is called for the current device with this flag array.
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;   // unknown: set it again (harmless)
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};
// DGSCT_OK when the current device is sm_100 (compute capability 10.x), DGSCT_ERR_ARCH otherwise (cached per device)
int check_arch();
// sticky pipeline-timeout flag in host-mapped memory: kernels raise it, every entry point reads it without synchronising
int* timeout_flag_device_ptr();
int timeout_flag_host_value(int reset);

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- dtype-erased element access (activations are fp32 or bf16 at run time) ---------------------
__device__ __forceinline__ float ld_act(const void* p, size_t i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_act(void* p, size_t i, float v, int is_bf16) {
  if (is_bf16) reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reductions; `red` is >= 32 floats of shared memory; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

}  // namespace dgsct
