// Windowed similarity losses as single-pass tile kernels with warp/block reductions:
//   SSIM  (reference ssimloss.py:11-40, 7x7 uniform window, valid)
//   LNCC  (reference lnccloss.py:7-56, 9x9 box sums, zero padding 4)
//   MI    (reference miloss.py:26-57, 64-bin Gaussian Parzen joint histogram)
// plus the small single-channel filter used by the multi-scale variants
// (miloss.py:13-24).  Forward kernels reduce to a double accumulator; backward
// kernels are analytic (no autograd graph of box filters).
#include <cmath>

#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

__global__ void affine_scalar_kernel(const double* acc, float* out, double offset, double scale) {
  out[0] = (float)(offset + scale * acc[0]);
}

// ---- window functors: value and partial derivatives w.r.t. the five window sums
// (sx, sy, sxx, syy, sxy) -------------------------------------------------------
struct SsimF {
  static constexpr int WIN = 7, PAD = 0;
  __device__ static float eval(const float s[5], float d[5], bool want_d) {
    const float inv = 1.f / 49.f, cn = 49.f / 48.f, C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float ux = s[0] * inv, uy = s[1] * inv, uxx = s[2] * inv, uyy = s[3] * inv, uxy = s[4] * inv;
    const float vx = cn * (uxx - ux * ux), vy = cn * (uyy - uy * uy), vxy = cn * (uxy - ux * uy);
    const float A1 = 2.f * ux * uy + C1, A2 = 2.f * vxy + C2;
    const float B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
    const float D = B1 * B2;
    const float S = (A1 * A2) / D;
    if (want_d) {
      const float dA1 = A2 / D, dA2 = A1 / D, dB1 = -S / B1, dB2 = -S / B2;
      d[0] = inv * (dA1 * 2.f * uy + dA2 * (-2.f * cn * uy) + dB1 * 2.f * ux + dB2 * (-2.f * cn * ux));
      d[1] = inv * (dA1 * 2.f * ux + dA2 * (-2.f * cn * ux) + dB1 * 2.f * uy + dB2 * (-2.f * cn * uy));
      d[2] = inv * dB2 * cn;
      d[3] = inv * dB2 * cn;
      d[4] = inv * dA2 * 2.f * cn;
    }
    return S;
  }
};

struct LnccF {
  static constexpr int WIN = 9, PAD = 4;
  __device__ static float eval(const float s[5], float d[5], bool want_d) {
    const float ws = 81.f;
    const float Is = s[0], Js = s[1], I2 = s[2], J2 = s[3], IJ = s[4];
    const float uI = Is / ws, uJ = Js / ws;
    const float cross = IJ - uJ * Is - uI * Js + uI * uJ * ws;
    const float Iv = I2 - 2.f * uI * Is + uI * uI * ws;
    const float Jv = J2 - 2.f * uJ * Js + uJ * uJ * ws;
    const float D = Iv * Jv + 1e-5f;
    const float cc = cross * cross / D;
    if (want_d) {
      const float dc = 2.f * cross / D;
      const float dIv = -cc * Jv / D, dJv = -cc * Iv / D;
      d[0] = dc * (-Js / ws) + dIv * (-2.f * Is / ws);
      d[1] = dc * (-Is / ws) + dJv * (-2.f * Js / ws);
      d[2] = dIv;
      d[3] = dJv;
      d[4] = dc;
    }
    return cc;
  }
};

// ---- forward: block = 32x32 window positions ------------------------------------
constexpr int FQ = 32;
template <class F>
__global__ void __launch_bounds__(256) window_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, int H,
                                                         int W, int Hq, int Wq, double* acc) {
  constexpr int WIN = F::WIN, PAD = F::PAD, IT = FQ + WIN - 1;
  __shared__ float xs[IT * IT], ys[IT * IT];
  __shared__ float hs[5][IT * FQ];
  __shared__ double red[32];
  const long long n = blockIdx.z;
  const int q0y = blockIdx.y * FQ, q0x = blockIdx.x * FQ;
  const float* Xn = X + n * (long long)H * W;
  const float* Yn = Y + n * (long long)H * W;
  for (int i = threadIdx.x; i < IT * IT; i += blockDim.x) {
    const int r = i / IT, c = i - r * IT;
    const int gy = q0y + r - PAD, gx = q0x + c - PAD;
    float a = 0.f, b = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) { a = Xn[(long long)gy * W + gx]; b = Yn[(long long)gy * W + gx]; }
    xs[i] = a; ys[i] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IT * FQ; i += blockDim.x) {
    const int r = i / FQ, c = i - r * FQ;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float a = xs[r * IT + c + k], b = ys[r * IT + c + k];
      s0 += a; s1 += b; s2 += a * a; s3 += b * b; s4 += a * b;
    }
    hs[0][i] = s0; hs[1][i] = s1; hs[2][i] = s2; hs[3][i] = s3; hs[4][i] = s4;
  }
  __syncthreads();
  double local = 0.0;
  for (int i = threadIdx.x; i < FQ * FQ; i += blockDim.x) {
    const int r = i / FQ, c = i - r * FQ;
    if (q0y + r >= Hq || q0x + c >= Wq) continue;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
#pragma unroll
      for (int m = 0; m < 5; ++m) s[m] += hs[m][(r + k) * FQ + c];
    }
    float d[5];
    local += (double)F::eval(s, d, false);
  }
  local = block_sum_d(local, red);
  if (threadIdx.x == 0) atomicAdd(acc, local);
}

// ---- backward: block = 16x16 input pixels -----------------------------------------
constexpr int BP = 16;
template <class F>
__global__ void __launch_bounds__(256) window_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                         const float* __restrict__ gout, float gscale,
                                                         float* __restrict__ dX, float* __restrict__ dY, int H, int W,
                                                         int Hq, int Wq) {
  constexpr int WIN = F::WIN, PAD = F::PAD;
  constexpr int QT = BP + WIN - 1;        // window positions touching the tile
  constexpr int IT = BP + 2 * (WIN - 1);  // inputs needed by those windows
  __shared__ float xs[IT * IT], ys[IT * IT];
  __shared__ float hs[5][IT * QT];        // horizontal sums, later horizontal sums of D
  __shared__ float D[5][QT * QT];
  const long long n = blockIdx.z;
  const int p0y = blockIdx.y * BP, p0x = blockIdx.x * BP;
  const int qb_y = p0y + PAD - WIN + 1, qb_x = p0x + PAD - WIN + 1;  // first window position
  const int ib_y = qb_y - PAD, ib_x = qb_x - PAD;                    // first input row/col
  const float* Xn = X + n * (long long)H * W;
  const float* Yn = Y + n * (long long)H * W;
  for (int i = threadIdx.x; i < IT * IT; i += blockDim.x) {
    const int r = i / IT, c = i - r * IT;
    const int gy = ib_y + r, gx = ib_x + c;
    float a = 0.f, b = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) { a = Xn[(long long)gy * W + gx]; b = Yn[(long long)gy * W + gx]; }
    xs[i] = a; ys[i] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IT * QT; i += blockDim.x) {
    const int r = i / QT, c = i - r * QT;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float a = xs[r * IT + c + k], b = ys[r * IT + c + k];
      s0 += a; s1 += b; s2 += a * a; s3 += b * b; s4 += a * b;
    }
    hs[0][i] = s0; hs[1][i] = s1; hs[2][i] = s2; hs[3][i] = s3; hs[4][i] = s4;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < QT * QT; i += blockDim.x) {
    const int r = i / QT, c = i - r * QT;
    const int qy = qb_y + r, qx = qb_x + c;
    float d[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (qy >= 0 && qy < Hq && qx >= 0 && qx < Wq) {
      float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
#pragma unroll
        for (int m = 0; m < 5; ++m) s[m] += hs[m][(r + k) * QT + c];
      }
      F::eval(s, d, true);
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) D[m][i] = d[m];
  }
  __syncthreads();
  // horizontal sums of D over the WIN windows containing each pixel column
  for (int i = threadIdx.x; i < QT * BP; i += blockDim.x) {
    const int r = i / BP, c = i - r * BP;
    float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
#pragma unroll
      for (int m = 0; m < 5; ++m) t[m] += D[m][r * QT + c + k];
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) hs[m][i] = t[m];
  }
  __syncthreads();
  const float g = gout[0] * gscale;
  for (int i = threadIdx.x; i < BP * BP; i += blockDim.x) {
    const int r = i / BP, c = i - r * BP;
    const int py = p0y + r, px = p0x + c;
    if (py >= H || px >= W) continue;
    float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
#pragma unroll
      for (int m = 0; m < 5; ++m) t[m] += hs[m][(r + k) * BP + c];
    }
    const float xv = xs[(r + WIN - 1) * IT + c + WIN - 1], yv = ys[(r + WIN - 1) * IT + c + WIN - 1];
    const long long o = n * (long long)H * W + (long long)py * W + px;
    if (dX) dX[o] = g * (t[0] + 2.f * xv * t[2] + yv * t[4]);
    if (dY) dY[o] = g * (t[1] + 2.f * yv * t[3] + xv * t[4]);
  }
}

template <class F>
int window_fwd(const float* X, const float* Y, int N, int H, int W, double offset, double sign, float* out,
               double* scratch, cudaStream_t st) {
  const int Hq = H + 2 * F::PAD - F::WIN + 1, Wq = W + 2 * F::PAD - F::WIN + 1;
  SAN_CHECK_ARG(Hq > 0 && Wq > 0, "window loss: image %dx%d smaller than window", H, W);
  SAN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  dim3 grid(san_cdiv(Wq, FQ), san_cdiv(Hq, FQ), N);
  window_fwd_kernel<F><<<grid, 256, 0, st>>>(X, Y, H, W, Hq, Wq, scratch);
  SAN_LAUNCH_CHECK();
  affine_scalar_kernel<<<1, 1, 0, st>>>(scratch, out, offset, sign / ((double)N * Hq * Wq));
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

template <class F>
int window_bwd(const float* X, const float* Y, const float* gout, int N, int H, int W, float* dX, float* dY,
               cudaStream_t st) {
  const int Hq = H + 2 * F::PAD - F::WIN + 1, Wq = W + 2 * F::PAD - F::WIN + 1;
  SAN_CHECK_ARG(Hq > 0 && Wq > 0, "window loss: image %dx%d smaller than window", H, W);
  dim3 grid(san_cdiv(W, BP), san_cdiv(H, BP), N);
  const float gscale = (float)(-1.0 / ((double)N * Hq * Wq));
  window_bwd_kernel<F><<<grid, 256, 0, st>>>(X, Y, gout, gscale, dX, dY, H, W, Hq, Wq);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

// ---- mutual information -------------------------------------------------------------
constexpr int MI_B = 64;    // bins (fixed by the kernels' register tiling)
constexpr int MI_CH = 64;   // pixels staged per iteration

// joint[n][b1][b2] += sum_px pI[b1,px] pJ[b2,px];  mI[n][b] += sum_px pI[b,px]; same for mJ.
__global__ void __launch_bounds__(256) mi_hist_fwd_kernel(const float* __restrict__ I, const float* __restrict__ J,
                                                          float* __restrict__ joint, float* __restrict__ mI,
                                                          float* __restrict__ mJ, int P, float minv, float binw,
                                                          float inv2s2, float norm) {
  __shared__ __align__(16) float pI[MI_CH][MI_B];
  __shared__ __align__(16) float pJ[MI_CH][MI_B];
  const long long n = blockIdx.x;
  const float* In = I + n * (long long)P;
  const float* Jn = J + n * (long long)P;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  float acc[4][4];
  float ma[4] = {0.f, 0.f, 0.f, 0.f}, mb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int base = blockIdx.y * MI_CH; base < P; base += gridDim.y * MI_CH) {
    __syncthreads();
    // fill: thread handles pixel (tid % 64), bins [16*(tid/64), +16)
    {
      const int px = threadIdx.x & (MI_CH - 1), part = threadIdx.x >> 6;
      const bool ok = base + px < P;
      const float vi = ok ? In[base + px] : 0.f, vj = ok ? Jn[base + px] : 0.f;
#pragma unroll 8
      for (int b = part * 16; b < part * 16 + 16; ++b) {
        const float bin = minv + binw * b;
        const float di = vi - bin, dj = vj - bin;
        pI[px][b] = ok ? __expf(-di * di * inv2s2) * norm : 0.f;
        pJ[px][b] = ok ? __expf(-dj * dj * inv2s2) * norm : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < MI_CH; ++px) {
      const float4 a = *(const float4*)&pI[px][ti * 4];
      const float4 b = *(const float4*)&pJ[px][tj * 4];
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ma[i] += av[i];
        mb[i] += bv[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }
  float* jn = joint + n * MI_B * MI_B;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(jn + (ti * 4 + i) * MI_B + tj * 4 + j, acc[i][j]);
  if (tj == 0)
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(mI + n * MI_B + ti * 4 + i, ma[i]);
  if (ti == 0)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(mJ + n * MI_B + tj * 4 + j, mb[j]);
}

// dI[px] = sum_b1 (gmI[b1] + sum_b2 G[b1,b2] pJ[b2,px]) * dpI[b1,px]/dv ; symmetric for dJ
__global__ void __launch_bounds__(256) mi_hist_bwd_kernel(const float* __restrict__ I, const float* __restrict__ J,
                                                          const float* __restrict__ gjoint,
                                                          const float* __restrict__ gmI, const float* __restrict__ gmJ,
                                                          float* __restrict__ dI, float* __restrict__ dJ, int P,
                                                          float minv, float binw, float inv2s2, float norm) {
  __shared__ float G[MI_B][MI_B + 1];
  __shared__ float gi[MI_B], gj[MI_B];
  const long long n = blockIdx.x;
  for (int i = threadIdx.x; i < MI_B * MI_B; i += blockDim.x) G[i / MI_B][i % MI_B] = gjoint[n * MI_B * MI_B + i];
  if (threadIdx.x < MI_B) { gi[threadIdx.x] = gmI[n * MI_B + threadIdx.x]; gj[threadIdx.x] = gmJ[n * MI_B + threadIdx.x]; }
  __syncthreads();
  const float is2 = 2.f * inv2s2;  // 1/sigma^2
  for (int px = blockIdx.y * blockDim.x + threadIdx.x; px < P; px += gridDim.y * blockDim.x) {
    const float vi = I[n * (long long)P + px], vj = J[n * (long long)P + px];
    float p[MI_B];
    if (dI) {
#pragma unroll
      for (int b = 0; b < MI_B; ++b) { const float d = vj - (minv + binw * b); p[b] = __expf(-d * d * inv2s2) * norm; }
      float out = 0.f;
      for (int b1 = 0; b1 < MI_B; ++b1) {
        float t = gi[b1];
#pragma unroll
        for (int b2 = 0; b2 < MI_B; ++b2) t = fmaf(G[b1][b2], p[b2], t);
        const float d = vi - (minv + binw * b1);
        out += t * (__expf(-d * d * inv2s2) * norm) * (-d * is2);
      }
      dI[n * (long long)P + px] = out;
    }
    if (dJ) {
#pragma unroll
      for (int b = 0; b < MI_B; ++b) { const float d = vi - (minv + binw * b); p[b] = __expf(-d * d * inv2s2) * norm; }
      float out = 0.f;
      for (int b2 = 0; b2 < MI_B; ++b2) {
        float t = gj[b2];
#pragma unroll
        for (int b1 = 0; b1 < MI_B; ++b1) t = fmaf(G[b1][b2], p[b1], t);
        const float d = vj - (minv + binw * b2);
        out += t * (__expf(-d * d * inv2s2) * norm) * (-d * is2);
      }
      dJ[n * (long long)P + px] = out;
    }
  }
}

// single-channel KxK correlation with zero padding K/2 (multi-scale pyramid smoothing)
__global__ void filter2d_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int H,
                                int W, int K, long long total) {
  extern __shared__ float wk[];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) wk[i] = w[i];
  __syncthreads();
  const int pad = K / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % W);
    const long long t = i / W;
    const int r = (int)(t % H);
    const long long pl = t / H;
    const float* p = x + pl * (long long)H * W;
    float s = 0.f;
    for (int a = 0; a < K; ++a) {
      const int yy = r + a - pad;
      if (yy < 0 || yy >= H) continue;
      for (int b = 0; b < K; ++b) {
        const int xx = c + b - pad;
        if (xx < 0 || xx >= W) continue;
        s = fmaf(__ldg(p + (long long)yy * W + xx), wk[a * K + b], s);
      }
    }
    y[i] = s;
  }
}

}  // namespace

extern "C" {

int san_ssim_loss_fwd(const float* X, const float* Y, int N, int H, int W, float* out, double* scratch, void* stream) {
  SAN_CHECK_ARG(X && Y && out && scratch && N > 0, "san_ssim_loss_fwd: bad args");
  return window_fwd<SsimF>(X, Y, N, H, W, 1.0, -1.0, out, scratch, (cudaStream_t)stream);
}
int san_ssim_loss_bwd(const float* X, const float* Y, const float* gout, int N, int H, int W, float* dX, float* dY,
                      void* stream) {
  SAN_CHECK_ARG(X && Y && gout && (dX || dY) && N > 0, "san_ssim_loss_bwd: bad args");
  return window_bwd<SsimF>(X, Y, gout, N, H, W, dX, dY, (cudaStream_t)stream);
}
int san_lncc_loss_fwd(const float* I, const float* J, int N, int H, int W, float* out, double* scratch, void* stream) {
  SAN_CHECK_ARG(I && J && out && scratch && N > 0, "san_lncc_loss_fwd: bad args");
  return window_fwd<LnccF>(I, J, N, H, W, 0.0, -1.0, out, scratch, (cudaStream_t)stream);
}
int san_lncc_loss_bwd(const float* I, const float* J, const float* gout, int N, int H, int W, float* dI, float* dJ,
                      void* stream) {
  SAN_CHECK_ARG(I && J && gout && (dI || dJ) && N > 0, "san_lncc_loss_bwd: bad args");
  return window_bwd<LnccF>(I, J, gout, N, H, W, dI, dJ, (cudaStream_t)stream);
}

int san_mi_hist_fwd(const float* I, const float* J, float* joint, float* mI, float* mJ, int N, int P, int bins,
                    float sigma, float minv, float maxv, void* stream) {
  SAN_CHECK_ARG(I && J && joint && mI && mJ && N > 0 && P > 0, "san_mi_hist_fwd: bad args");
  SAN_CHECK_ARG(bins == MI_B, "san_mi_hist_fwd: only %d bins supported (got %d)", MI_B, bins);
  cudaStream_t st = (cudaStream_t)stream;
  SAN_CUDA(cudaMemsetAsync(joint, 0, sizeof(float) * (size_t)N * MI_B * MI_B, st));
  SAN_CUDA(cudaMemsetAsync(mI, 0, sizeof(float) * (size_t)N * MI_B, st));
  SAN_CUDA(cudaMemsetAsync(mJ, 0, sizeof(float) * (size_t)N * MI_B, st));
  int chunks = san_cdiv(P, MI_CH);
  if (chunks > 32) chunks = 32;
  dim3 grid(N, chunks);
  mi_hist_fwd_kernel<<<grid, 256, 0, st>>>(I, J, joint, mI, mJ, P, minv, (maxv - minv) / (bins - 1),
                                           1.f / (2.f * sigma * sigma), 1.f / (sqrtf(2.f * (float)M_PI) * sigma));
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_mi_hist_bwd(const float* I, const float* J, const float* gjoint, const float* gmI, const float* gmJ, float* dI,
                    float* dJ, int N, int P, int bins, float sigma, float minv, float maxv, void* stream) {
  SAN_CHECK_ARG(I && J && gjoint && gmI && gmJ && (dI || dJ) && N > 0 && P > 0, "san_mi_hist_bwd: bad args");
  SAN_CHECK_ARG(bins == MI_B, "san_mi_hist_bwd: only %d bins supported (got %d)", MI_B, bins);
  int chunks = san_cdiv(P, 256);
  if (chunks > 64) chunks = 64;
  dim3 grid(N, chunks);
  mi_hist_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(I, J, gjoint, gmI, gmJ, dI, dJ, P, minv,
                                                             (maxv - minv) / (bins - 1), 1.f / (2.f * sigma * sigma),
                                                             1.f / (sqrtf(2.f * (float)M_PI) * sigma));
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_filter2d(const float* x, const float* w, float* y, long long planes, int H, int W, int K, void* stream) {
  SAN_CHECK_ARG(x && w && y && planes > 0 && H > 0 && W > 0 && K > 0 && K <= 63 && (K & 1), "san_filter2d: bad args");
  const long long total = planes * H * W;
  long long g = (total + 255) / 256;
  const long long cap = (long long)san_num_sms() * 16;
  if (g > cap) g = cap;
  filter2d_kernel<<<(int)g, 256, sizeof(float) * K * K, (cudaStream_t)stream>>>(x, w, y, H, W, K, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
