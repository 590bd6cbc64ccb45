// On-device misalignment augmentation of the auxiliary modality (reference augment.py:7-66, applied per
// training step in train.py:207-212):
//   grid = affine_grid(theta, align_corners=False) + bicubic_upsample(ctrl 9x9 -> HxW, align_corners=False)
//   out  = grid_sample(img, grid, bilinear, padding_mode='reflection', align_corners=False)
// The reference runs affine_grid, interpolate(bicubic), permute, add and two grid_sample calls (real, imag);
// here: one kernel builds the grid (2 floats per pixel written), one gathers all planes of a pixel with the
// coordinates / weights computed once.  HBM-bound: (4 C k) read + (4 C k) written + 8 B grid per pixel,
// k = 2 for complex64.
#include <cmath>

#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

// torch upsample_bicubic2d: cubic convolution, A = -0.75
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
  const float A = -0.75f;
  float x = t + 1.f;
  c[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  c[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  c[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

// grid[n,h,w,:] = theta[n] (x_w, y_h, 1) + bicubic(ctrl[n,:,G,G])(h, w)
__global__ void __launch_bounds__(256) augment_grid_kernel(const float* __restrict__ theta, const float* __restrict__ ctrl,
                                                           int G, float* __restrict__ grid, int N, int H, int W) {
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int n = (int)(i / ((long long)W * H));
    const float x = (2.f * w + 1.f) / W - 1.f, y = (2.f * h + 1.f) / H - 1.f;
    const float* t = theta + n * 6;
    float gx = t[0] * x + t[1] * y + t[2];
    float gy = t[3] * x + t[4] * y + t[5];
    if (ctrl) {
      // source index of the bicubic up-sampling, align_corners=False (not clamped for the cubic filter)
      const float sy = (float)G / H * (h + 0.5f) - 0.5f, sx = (float)G / W * (w + 0.5f) - 0.5f;
      const int iy = (int)floorf(sy), ix = (int)floorf(sx);
      float cy[4], cx[4];
      cubic_coeffs(sy - iy, cy);
      cubic_coeffs(sx - ix, cx);
      const float* c0 = ctrl + (size_t)n * 2 * G * G;
      float ax = 0.f, ay = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int yy = min(max(iy - 1 + a, 0), G - 1);
        float rx = 0.f, ry = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xx = min(max(ix - 1 + b, 0), G - 1);
          rx += cx[b] * c0[yy * G + xx];
          ry += cx[b] * c0[G * G + yy * G + xx];
        }
        ax += cy[a] * rx;
        ay += cy[a] * ry;
      }
      gx += ax;
      gy += ay;
    }
    reinterpret_cast<float2*>(grid)[i] = make_float2(gx, gy);
  }
}

// torch grid_sampler reflect_coordinates(in, twice_low = -1, twice_high = 2*size - 1) followed by clip_coordinates
__device__ __forceinline__ float reflect_clip(float in, int size) {
  const float mn = -0.5f, span = (float)size;
  in = fabsf(in - mn);
  const float extra = fmodf(in, span);
  const int flips = (int)floorf(in / span);
  const float r = (flips & 1) ? (span - extra + mn) : (extra + mn);
  return fminf((float)(size - 1), fmaxf(r, 0.f));
}

// out[n, c, ho, wo, k] = bilinear(img[n, c, :, :, k]) at the reflected grid position; k < K interleaved components
__global__ void __launch_bounds__(256) warp_reflect_kernel(const float* __restrict__ img, const float* __restrict__ grid,
                                                           float* __restrict__ out, int N, int C, int H, int W, int Ho,
                                                           int Wo, int K) {
  const long long total = (long long)N * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ((long long)Ho * Wo));
    const long long p = i - (long long)n * Ho * Wo;
    const float2 g = reinterpret_cast<const float2*>(grid)[i];
    const float ix = reflect_clip(((g.x + 1.f) * W - 1.f) * 0.5f, W);
    const float iy = reflect_clip(((g.y + 1.f) * H - 1.f) * 0.5f, H);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    // after the clip x0, y0 are inside; x1 / y1 may be one past the edge (weight then 0 in exact arithmetic)
    const bool okx = x1 < W, oky = y1 < H;
    const float w00 = wx0 * wy0, w01 = okx ? wx1 * wy0 : 0.f, w10 = oky ? wx0 * wy1 : 0.f,
                w11 = (okx && oky) ? wx1 * wy1 : 0.f;
    const int xb = okx ? x1 : x0, yb = oky ? y1 : y0;
    for (int c = 0; c < C; ++c) {
      const float* src = img + ((size_t)(n * C + c) * H * W) * K;
      float* dst = out + ((size_t)(n * C + c) * Ho * Wo + p) * K;
      for (int k = 0; k < K; ++k) {
        const float v = w00 * src[((size_t)y0 * W + x0) * K + k] + w01 * src[((size_t)y0 * W + xb) * K + k] +
                        w10 * src[((size_t)yb * W + x0) * K + k] + w11 * src[((size_t)yb * W + xb) * K + k];
        dst[k] = v;
      }
    }
  }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)san_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

int san_augment_grid(const float* theta, const float* ctrl, int G, float* grid, int N, int H, int W, void* stream) {
  SAN_CHECK_ARG(theta && grid && N > 0 && H > 0 && W > 0 && (!ctrl || G >= 1), "san_augment_grid: bad args");
  augment_grid_kernel<<<grid_for((long long)N * H * W), 256, 0, (cudaStream_t)stream>>>(theta, ctrl, G, grid, N, H, W);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_warp_reflect(const float* img, const float* grid, float* out, int N, int C, int H, int W, int Ho, int Wo,
                     int interleave, void* stream) {
  SAN_CHECK_ARG(img && grid && out && N > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && interleave >= 1 &&
                    interleave <= 2,
                "san_warp_reflect: bad args");
  warp_reflect_kernel<<<grid_for((long long)N * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(img, grid, out, N, C, H, W, Ho,
                                                                                         Wo, interleave);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
