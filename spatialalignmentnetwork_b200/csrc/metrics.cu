// Evaluation metrics of CSModel.test() as device reductions.  The reference moves every
// image to the host and calls numpy / skimage per image (metrics.py:12-69, model.py:265-286):
//   * san_error_sums : sum (a-b)^2, sum |a-b|, sum a^2 over the whole batch in fp64 -> MSE, MAE, NMSE and
//     PSNR = 10 log10(1 / MSE) (metrics.py:23-38; compare_psnr with data_range = 1 over the 4-D batch);
//   * san_mi_metric  : np.histogram2d(x, y, bins, range) per image + the plug-in mutual information
//     sum xlogy(Pxy, Pxy) - xlogy(Pxy, Px Py)  (metrics.py:54-68), one CTA per image, shared-memory histogram.
// (SSIM is 1 - ssimloss, csrc/losses.cu.)
#include <cmath>

#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

__global__ void __launch_bounds__(256) error_sums_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         long long n, double* __restrict__ out) {
  __shared__ double red[32];
  double s2 = 0.0, s1 = 0.0, sa = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = a[i], d = x - (double)b[i];
    s2 += d * d; s1 += fabs(d); sa += x * x;
  }
  s2 = block_sum_d(s2, red);
  s1 = block_sum_d(s1, red);
  sa = block_sum_d(sa, red);
  if (threadIdx.x == 0) { atomicAdd(out + 0, s2); atomicAdd(out + 1, s1); atomicAdd(out + 2, sa); }
}

constexpr int MIM_MAXB = 64;

// numpy histogramdd bin of value x for `bins` equal bins on [lo, hi]: edges e_k = linspace(lo, hi, bins+1),
// bin = searchsorted(edges, x, 'right') - 1, the right-most edge belongs to the last bin, outside -> dropped.
__device__ __forceinline__ int np_bin(float xf, double lo, double hi, int bins) {
  const double x = (double)xf;
  if (!(x >= lo && x <= hi)) return -1;
  if (x == hi) return bins - 1;
  const double step = (hi - lo) / bins;
  int k = (int)floor((x - lo) / step);
  if (k < 0) k = 0;
  if (k > bins - 1) k = bins - 1;
  // correct the division's rounding against the edges themselves (lo + k*step, as np.linspace builds them)
  while (k > 0 && x < lo + k * step) --k;
  while (k < bins - 1 && x >= lo + (k + 1) * step) ++k;
  return k;
}

__global__ void __launch_bounds__(512) mi_metric_kernel(const float* __restrict__ X, const float* __restrict__ Y, int P,
                                                        int bins, double lo, double hi, double* __restrict__ out) {
  __shared__ unsigned int hist[MIM_MAXB * MIM_MAXB];
  __shared__ double px[MIM_MAXB], py[MIM_MAXB];
  __shared__ double red[32];
  const int n = blockIdx.x;
  const float* x = X + (size_t)n * P;
  const float* y = Y + (size_t)n * P;
  const int nb2 = bins * bins;
  for (int i = threadIdx.x; i < nb2; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const int bx = np_bin(x[i], lo, hi, bins), by = np_bin(y[i], lo, hi, bins);
    if (bx >= 0 && by >= 0) atomicAdd(&hist[bx * bins + by], 1u);
  }
  __syncthreads();
  double tot = 0.0;
  for (int i = threadIdx.x; i < nb2; i += blockDim.x) tot += (double)hist[i];
  tot = block_sum_d(tot, red);
  const double inv = 1.0 / (tot + 1e-10);
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    double r = 0.0, c = 0.0;
    for (int j = 0; j < bins; ++j) { r += (double)hist[i * bins + j] * inv; c += (double)hist[j * bins + i] * inv; }
    px[i] = r; py[i] = c;
  }
  __syncthreads();
  double mi = 0.0;
  for (int i = threadIdx.x; i < nb2; i += blockDim.x) {
    const double p = (double)hist[i] * inv;
    if (p > 0.0) mi += p * log(p) - p * log(px[i / bins] * py[i % bins]);   // xlogy(0, .) = 0
  }
  mi = block_sum_d(mi, red);
  if (threadIdx.x == 0) out[n] = mi;
}

}  // namespace

extern "C" {

int san_error_sums(const float* a, const float* b, long long n, double* out3, void* stream) {
  SAN_CHECK_ARG(a && b && out3 && n > 0, "san_error_sums: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  SAN_CUDA(cudaMemsetAsync(out3, 0, 3 * sizeof(double), st));
  long long g = (n + 255) / 256;
  const long long cap = (long long)san_num_sms() * 8;
  if (g > cap) g = cap;
  error_sums_kernel<<<(int)g, 256, 0, st>>>(a, b, n, out3);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_mi_metric(const float* x, const float* y, int N, int P, int bins, float minv, float maxv, double* out,
                  void* stream) {
  SAN_CHECK_ARG(x && y && out && N > 0 && P > 0, "san_mi_metric: bad args");
  SAN_CHECK_ARG(bins > 0 && bins <= MIM_MAXB && maxv > minv, "san_mi_metric: 1..%d bins, maxv > minv", MIM_MAXB);
  mi_metric_kernel<<<N, 512, 0, (cudaStream_t)stream>>>(x, y, P, bins, (double)minv, (double)maxv, out);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
