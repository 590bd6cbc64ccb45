// Shared helpers for the san_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define SAN_OK 0
#define SAN_ERR_ARG -1
#define SAN_ERR_CUDA -2
#define SAN_ERR_UNSUPPORTED -3

void san_set_error(const char* fmt, ...);

#define SAN_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      san_set_error(__VA_ARGS__);                \
      return SAN_ERR_ARG;                        \
    }                                            \
  } while (0)

#define SAN_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      san_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call,                 \
                    cudaGetErrorString(e__));                                  \
      return SAN_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

#include <atomic>
extern std::atomic<long long> g_san_launches;
// one per kernel launch: counts it and surfaces launch-configuration errors
#define SAN_LAUNCH_CHECK()         \
  do {                             \
    g_san_launches.fetch_add(1);   \
    SAN_CUDA(cudaGetLastError());  \
  } while (0)

static inline int san_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Number of SMs of the current device (cached; 148 on B200).
int san_num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in every thread. `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : 0.0;
  r = warp_sum_d(r);
  return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
