// Spatial-alignment kernels: identity grid + displacement (reference cross.py:23-30),
// bilinear grid_sample with zero padding / align_corners=False (cross.py:32-38) forward
// and backward, and the displacement smoothness loss (model.py:21-28).
#include "san_common.cuh"
#include "../../include/san_b200.h"

int san_finalize_scalar(const double* acc, float* out, double scale, cudaStream_t st);

namespace {

// net output x [N,2,H,W] -> grid [N,H,W,2] = identity + offset
__global__ void grid_from_offset_kernel(const float* __restrict__ x, float* __restrict__ grid, int H, int W,
                                        long long total) {
  const long long HW = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    const int h = (int)(p / W), w = (int)(p - (long long)h * W);
    const float gx = (2.f * w + 1.f) / W - 1.f;
    const float gy = (2.f * h + 1.f) / H - 1.f;
    const float ox = x[(n * 2) * HW + p], oy = x[(n * 2 + 1) * HW + p];
    ((float2*)grid)[i] = make_float2(gx + ox, gy + oy);
  }
}

// g [N,H,W,2] -> dx [N,2,H,W]
__global__ void grid_to_nchw_kernel(const float* __restrict__ g, float* __restrict__ dx, long long HW,
                                    long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    const float2 v = ((const float2*)g)[i];
    dx[(n * 2) * HW + p] = v.x;
    dx[(n * 2 + 1) * HW + p] = v.y;
  }
}

__device__ __forceinline__ bool inb(int x, int y, int W, int H) { return x >= 0 && x < W && y >= 0 && y < H; }

// img [N,C,H,W], grid [N,Ho,Wo,2] -> out [N,C,Ho,Wo]
__global__ void warp_fwd_kernel(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                                int C, int H, int W, int Ho, int Wo, long long total) {
  const long long HWo = (long long)Ho * Wo, HW = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HWo, p = i - n * HWo;
    const float2 gq = ((const float2*)grid)[i];
    const float ix = ((gq.x + 1.f) * W - 1.f) / 2.f;
    const float iy = ((gq.y + 1.f) * H - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float nw = ((fx + 1.f) - ix) * ((fy + 1.f) - iy);
    const float ne = (ix - fx) * ((fy + 1.f) - iy);
    const float sw = ((fx + 1.f) - ix) * (iy - fy);
    const float se = (ix - fx) * (iy - fy);
    const bool bnw = inb(x0, y0, W, H), bne = inb(x1, y0, W, H), bsw = inb(x0, y1, W, H), bse = inb(x1, y1, W, H);
    for (int c = 0; c < C; ++c) {
      const float* im = img + (n * C + c) * HW;
      float v = 0.f;
      if (bnw) v += im[(long long)y0 * W + x0] * nw;
      if (bne) v += im[(long long)y0 * W + x1] * ne;
      if (bsw) v += im[(long long)y1 * W + x0] * sw;
      if (bse) v += im[(long long)y1 * W + x1] * se;
      out[(n * C + c) * HWo + p] = v;
    }
  }
}

// dimg must be zero-initialised (scatter-add); dgrid [N,Ho,Wo,2]
__global__ void warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ img,
                                const float* __restrict__ grid, float* __restrict__ dimg, float* __restrict__ dgrid,
                                int C, int H, int W, int Ho, int Wo, long long total) {
  const long long HWo = (long long)Ho * Wo, HW = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HWo, p = i - n * HWo;
    const float2 gq = ((const float2*)grid)[i];
    const float ix = ((gq.x + 1.f) * W - 1.f) / 2.f;
    const float iy = ((gq.y + 1.f) * H - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float ax = (fx + 1.f) - ix, bx = ix - fx, ay = (fy + 1.f) - iy, by = iy - fy;
    const float nw = ax * ay, ne = bx * ay, sw = ax * by, se = bx * by;
    const bool bnw = inb(x0, y0, W, H), bne = inb(x1, y0, W, H), bsw = inb(x0, y1, W, H), bse = inb(x1, y1, W, H);
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float go = gout[(n * C + c) * HWo + p];
      const float* im = img + (n * C + c) * HW;
      float* di = dimg ? dimg + (n * C + c) * HW : nullptr;
      if (bnw) {
        const float v = im[(long long)y0 * W + x0];
        gix -= v * ay * go; giy -= v * ax * go;
        if (di) atomicAdd(di + (long long)y0 * W + x0, nw * go);
      }
      if (bne) {
        const float v = im[(long long)y0 * W + x1];
        gix += v * ay * go; giy -= v * bx * go;
        if (di) atomicAdd(di + (long long)y0 * W + x1, ne * go);
      }
      if (bsw) {
        const float v = im[(long long)y1 * W + x0];
        gix -= v * by * go; giy += v * ax * go;
        if (di) atomicAdd(di + (long long)y1 * W + x0, sw * go);
      }
      if (bse) {
        const float v = im[(long long)y1 * W + x1];
        gix += v * by * go; giy += v * bx * go;
        if (di) atomicAdd(di + (long long)y1 * W + x1, se * go);
      }
    }
    if (dgrid) ((float2*)dgrid)[i] = make_float2(gix * (W * 0.5f), giy * (H * 0.5f));
  }
}

// s[n,h,w,c] at n*sn + h*sh + w*sw + c*sc.  acc[0] += sum dW^2, acc[1] += sum dH^2
__global__ void grad_loss_fwd_kernel(const float* __restrict__ s, long long sn, long long sh, long long sw,
                                     long long sc, int H, int W, long long total, double* acc) {
  __shared__ double red[32];
  double ax = 0.0, ay = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 1);
    long long t = i >> 1;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const long long n = t / H;
    const float* q = s + n * sn + h * sh + w * sw + c * sc;
    const float v = q[0];
    if (w + 1 < W) { const float d = q[sw] - v; ax += (double)(d * d); }
    if (h + 1 < H) { const float d = q[sh] - v; ay += (double)(d * d); }
  }
  ax = block_sum_d(ax, red);
  ay = block_sum_d(ay, red);
  if (threadIdx.x == 0) { atomicAdd(acc, ax); atomicAdd(acc + 1, ay); }
}

__global__ void grad_loss_finalize_kernel(const double* acc, float* out, double cx, double cy) {
  out[0] = (float)((acc[0] / cx + acc[1] / cy) * 0.5);
}

// ds contiguous [N,H,W,2]
__global__ void grad_loss_bwd_kernel(const float* __restrict__ s, long long sn, long long sh, long long sw,
                                     long long sc, const float* __restrict__ gout, float* __restrict__ ds, int H,
                                     int W, long long total, float kx, float ky) {
  const float g = gout[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 1);
    long long t = i >> 1;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const long long n = t / H;
    const float* q = s + n * sn + h * sh + w * sw + c * sc;
    const float v = q[0];
    float dx = 0.f, dy = 0.f;
    if (w > 0) dx += v - q[-sw];
    if (w + 1 < W) dx -= q[sw] - v;
    if (h > 0) dy += v - q[-sh];
    if (h + 1 < H) dy -= q[sh] - v;
    ds[i] = g * (kx * dx + ky * dy);
  }
}

inline int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)san_num_sms() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

extern "C" {

int san_grid_from_offset(const float* x_nchw, float* grid, int N, int H, int W, void* stream) {
  SAN_CHECK_ARG(x_nchw && grid && N > 0 && H > 0 && W > 0, "san_grid_from_offset: bad args");
  const long long total = (long long)N * H * W;
  grid_from_offset_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x_nchw, grid, H, W, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_grid_to_nchw(const float* g_nhwc, float* dx_nchw, int N, int H, int W, void* stream) {
  SAN_CHECK_ARG(g_nhwc && dx_nchw && N > 0 && H > 0 && W > 0, "san_grid_to_nchw: bad args");
  const long long total = (long long)N * H * W;
  grid_to_nchw_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(g_nhwc, dx_nchw, (long long)H * W, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_warp_fwd(const float* img, const float* grid, float* out, int N, int C, int H, int W, int Ho, int Wo,
                 void* stream) {
  SAN_CHECK_ARG(img && grid && out && N > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "san_warp_fwd: bad args");
  const long long total = (long long)N * Ho * Wo;
  warp_fwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(img, grid, out, C, H, W, Ho, Wo, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_warp_bwd(const float* gout, const float* img, const float* grid, float* dimg, float* dgrid, int N, int C,
                 int H, int W, int Ho, int Wo, void* stream) {
  SAN_CHECK_ARG(gout && img && grid && (dimg || dgrid) && N > 0 && C > 0, "san_warp_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (dimg) SAN_CUDA(cudaMemsetAsync(dimg, 0, sizeof(float) * (size_t)N * C * H * W, st));
  const long long total = (long long)N * Ho * Wo;
  warp_bwd_kernel<<<ew_grid(total), 256, 0, st>>>(gout, img, grid, dimg, dgrid, C, H, W, Ho, Wo, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_grad_loss_fwd(const float* s, long long sn, long long sh, long long sw, long long sc, int N, int H, int W,
                      float* out, double* scratch, void* stream) {
  SAN_CHECK_ARG(s && out && scratch && N > 0 && H > 1 && W > 1, "san_grad_loss_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  SAN_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st));
  const long long total = (long long)N * H * W * 2;
  grad_loss_fwd_kernel<<<ew_grid(total), 256, 0, st>>>(s, sn, sh, sw, sc, H, W, total, scratch);
  SAN_LAUNCH_CHECK();
  grad_loss_finalize_kernel<<<1, 1, 0, st>>>(scratch, out, (double)N * H * (W - 1) * 2, (double)N * (H - 1) * W * 2);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_grad_loss_bwd(const float* s, long long sn, long long sh, long long sw, long long sc, int N, int H, int W,
                      const float* gout, float* ds, void* stream) {
  SAN_CHECK_ARG(s && gout && ds && N > 0 && H > 1 && W > 1, "san_grad_loss_bwd: bad args");
  const long long total = (long long)N * H * W * 2;
  const float kx = (float)(1.0 / ((double)N * H * (W - 1) * 2));
  const float ky = (float)(1.0 / ((double)N * (H - 1) * W * 2));
  grad_loss_bwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(s, sn, sh, sw, sc, gout, ds, H, W, total, kx, ky);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
