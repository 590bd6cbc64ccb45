// Small kernels of the GAN branch:
//   * spectral-norm weight preprocessing of gan.py:24 (torch.nn.utils.spectral_norm, one power iteration per
//     training forward, eps 1e-12, dim 0):  v = normalize(W^T u), u = normalize(W v), sigma = u . (W v),
//     W_sn = W / sigma, and its backward  dW = (G - <G, W_sn> u v^T) / sigma   (u, v are constants);
//   * the point-wise mean losses of the GAN step: F.l1_loss (model.py:138-139) and the hinge / linear terms of
//     gan.loss_gan (gan.py:131-137).
// All are HBM-bound passes over <= 2.4 M weights or one [N,1,H,W] image batch; reductions in fp64.
#include <cmath>

#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

// out[c] = sum_r W[r][c] * u[r]   (thread per column, coalesced over c)
__global__ void __launch_bounds__(256) sn_matvec_t_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                                          float* __restrict__ out, int rows, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double acc = 0.0;
  for (int r = 0; r < rows; ++r) acc += (double)w[(size_t)r * cols + c] * (double)u[r];
  out[c] = (float)acc;
}

// out[r] = sum_c W[r][c] * v[c]   (warp per row)
__global__ void __launch_bounds__(256) sn_matvec_kernel(const float* __restrict__ w, const float* __restrict__ v,
                                                        float* __restrict__ out, int rows, int cols) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* wr = w + (size_t)r * cols;
  double acc = 0.0;
  for (int c = lane; c < cols; c += 32) acc += (double)wr[c] * (double)v[c];
  acc = warp_sum_d(acc);
  if (lane == 0) out[r] = (float)acc;
}

// Single block.  normalise != 0: vec = x / max(||x||, eps) (F.normalize).  sigma (optional) = vec . x
__global__ void __launch_bounds__(256) sn_finish_kernel(const float* __restrict__ x, float* __restrict__ vec, int n,
                                                        float eps, int normalise, float* __restrict__ sigma) {
  __shared__ double red[32];
  if (normalise) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i] * (double)x[i];
    s = block_sum_d(s, red);
    const float nrm = fmaxf((float)sqrt(s), eps);
    for (int i = threadIdx.x; i < n; i += blockDim.x) vec[i] = x[i] / nrm;
    __syncthreads();
  }
  if (sigma) {
    double d = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d += (double)vec[i] * (double)x[i];
    d = block_sum_d(d, red);
    if (threadIdx.x == 0) sigma[0] = (float)d;
  }
}

__global__ void __launch_bounds__(256) sn_scale_kernel(const float* __restrict__ w, const float* __restrict__ sigma,
                                                       float* __restrict__ out, long long n) {
  const float s = sigma[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = w[i] / s;
}

__global__ void __launch_bounds__(256) dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                  double* __restrict__ acc) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += (double)a[i] * (double)b[i];
  s = block_sum_d(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

// dw[r][c] = (g[r][c] - dot * u[r] * v[c]) / sigma
__global__ void __launch_bounds__(256) sn_bwd_kernel(const float* __restrict__ g, const float* __restrict__ u,
                                                     const float* __restrict__ v, const float* __restrict__ sigma,
                                                     const double* __restrict__ dot, float* __restrict__ dw, int rows,
                                                     int cols) {
  const float s = sigma[0], d = (float)dot[0];
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    dw[i] = (g[i] - d * u[r] * v[c]) / s;
  }
}

// ---- point-wise mean losses ----------------------------------------------------------------------
// mode 0: |x - y|;  mode 1: max(sign * x, -1)  (torch.clamp(min=-1));  mode 2: sign * x
__device__ __forceinline__ float pair_term(float x, float y, int mode, float sign) {
  if (mode == 0) return fabsf(x - y);
  const float t = sign * x;
  return mode == 1 ? fmaxf(t, -1.f) : t;
}

__global__ void __launch_bounds__(256) pair_loss_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            long long n, int mode, float sign, double* __restrict__ acc) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += (double)pair_term(x[i], mode == 0 ? y[i] : 0.f, mode, sign);
  s = block_sum_d(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

__global__ void mean_finalize_kernel(const double* acc, float* out, double inv_n) { out[0] = (float)(acc[0] * inv_n); }

__global__ void __launch_bounds__(256) pair_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ gout, long long n, int mode,
                                                            float sign, float* __restrict__ dx, float* __restrict__ dy) {
  const float g = gout[0] / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d;
    if (mode == 0) {
      const float t = x[i] - y[i];
      d = t > 0.f ? g : (t < 0.f ? -g : 0.f);            // sign(), 0 at 0 like torch
      if (dy) dy[i] = -d;
    } else if (mode == 1) {
      d = (sign * x[i] >= -1.f) ? g * sign : 0.f;        // clamp passes the gradient where input >= min
    } else {
      d = g * sign;
    }
    if (dx) dx[i] = d;
  }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)san_num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

int san_sn_sigma(const float* w, float* u, float* v, float* tmp, float* sigma, int rows, int cols, float eps,
                 int power_iteration, void* stream) {
  SAN_CHECK_ARG(w && u && v && tmp && sigma && rows > 0 && cols > 0, "san_sn_sigma: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (power_iteration) {
    sn_matvec_t_kernel<<<san_cdiv(cols, 256), 256, 0, st>>>(w, u, tmp, rows, cols);
    SAN_LAUNCH_CHECK();
    sn_finish_kernel<<<1, 256, 0, st>>>(tmp, v, cols, eps, 1, nullptr);
    SAN_LAUNCH_CHECK();
  }
  sn_matvec_kernel<<<san_cdiv(rows, 8), 256, 0, st>>>(w, v, tmp, rows, cols);
  SAN_LAUNCH_CHECK();
  sn_finish_kernel<<<1, 256, 0, st>>>(tmp, u, rows, eps, power_iteration ? 1 : 0, sigma);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_sn_scale(const float* w, const float* sigma, float* out, long long n, void* stream) {
  SAN_CHECK_ARG(w && sigma && out && n > 0, "san_sn_scale: bad args");
  sn_scale_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(w, sigma, out, n);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_sn_bwd(const float* g, const float* w_sn, const float* u, const float* v, const float* sigma, double* scratch,
               float* dw, int rows, int cols, void* stream) {
  SAN_CHECK_ARG(g && w_sn && u && v && sigma && scratch && dw && rows > 0 && cols > 0, "san_sn_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)rows * cols;
  SAN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  dot_kernel<<<grid_for(n), 256, 0, st>>>(g, w_sn, n, scratch);
  SAN_LAUNCH_CHECK();
  sn_bwd_kernel<<<grid_for(n), 256, 0, st>>>(g, u, v, sigma, scratch, dw, rows, cols);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_pair_loss_fwd(const float* x, const float* y, long long n, int mode, float sign, float* out, double* scratch,
                      void* stream) {
  SAN_CHECK_ARG(x && out && scratch && n > 0 && mode >= 0 && mode <= 2 && (mode != 0 || y), "san_pair_loss_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  SAN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  pair_loss_fwd_kernel<<<grid_for(n), 256, 0, st>>>(x, y, n, mode, sign, scratch);
  SAN_LAUNCH_CHECK();
  mean_finalize_kernel<<<1, 1, 0, st>>>(scratch, out, 1.0 / (double)n);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_pair_loss_bwd(const float* x, const float* y, const float* gout, long long n, int mode, float sign, float* dx,
                      float* dy, void* stream) {
  SAN_CHECK_ARG(x && gout && (dx || dy) && n > 0 && mode >= 0 && mode <= 2 && (mode != 0 || y),
                "san_pair_loss_bwd: bad args");
  pair_loss_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, gout, n, mode, sign, dx, mode == 0 ? dy : nullptr);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
