// Per-plane normalisation / activation / resampling kernels for the U-Nets:
// InstanceNorm2d + LeakyReLU (reference varnet.py:139-146,176-182,235), BatchNorm2d +
// LeakyReLU (reference unet.py:119-140), the per-sample group norm of NormUnet
// (varnet.py:257-273), avg_pool2d(2) (varnet.py:98, unet.py:137), nearest x2
// up-sampling (unet.py:130) and the pixel shuffles used by the 2x2 stride-2
// transposed convolution (varnet.py:176-179).
//
// A "plane" is one (n, c) image of P = H*W contiguous floats.  Normalisation is
// expressed in CENTRED form  out = lrelu(a[plane] * (y - mu[plane]) + b[plane])  (so a
// large mean/std ratio costs no precision, like PyTorch's (x - mean) * rstd) with the
// coefficients produced by tiny finalize kernels from two-pass plane statistics, so a
// conv output is read twice (stats; L2-resident second pass) and written once.
#include <cooperative_groups.h>

#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

// ---- statistics -------------------------------------------------------------
// in_a / in_b (optional): the InstanceNorm coefficients of in_finalize_fwd_kernel written by the same thread (one launch
// instead of two per normalised tensor); same float arithmetic on the same rounded m2.
__global__ void __launch_bounds__(256) plane_stats_kernel(const float* __restrict__ x, float* __restrict__ mean,
                                                          float* __restrict__ m2, int P, float* __restrict__ in_a,
                                                          float* __restrict__ in_b, float eps) {
  // ONE pass: shifted-data sums in fp64 (pivot = first element of the plane, so the subtraction
  // Q - S^2/P cancels at most a few bits): mean = K + S/P, m2 = sum (x - mean)^2 = Q - S^2/P.
  __shared__ double red[32];
  const float* p = x + (long long)blockIdx.x * P;
  const double K = (double)__ldg(p);
  double s = 0.0, q = 0.0;
  if ((P & 3) == 0) {
    const float4* p4 = (const float4*)p;
    const int n4 = P >> 2;
    for (int i = threadIdx.x; i < n4; i += 2 * blockDim.x) {
      const int i2 = i + blockDim.x;
      const float4 va = __ldg(p4 + i);
      float4 vb = make_float4((float)K, (float)K, (float)K, (float)K);
      if (i2 < n4) vb = __ldg(p4 + i2);
      const double d0 = (double)va.x - K, d1 = (double)va.y - K, d2 = (double)va.z - K, d3 = (double)va.w - K;
      const double e0 = (double)vb.x - K, e1 = (double)vb.y - K, e2 = (double)vb.z - K, e3 = (double)vb.w - K;
      s += ((d0 + d1) + (d2 + d3)) + ((e0 + e1) + (e2 + e3));
      q += ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) + ((e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3));
    }
  } else {
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      const double d = (double)__ldg(p + i) - K;
      s += d;
      q += d * d;
    }
  }
  // (the float-K padding of the vector path is exact only when K is a float, which it is: it was loaded as one)
  const double S = block_sum_d(s, red);
  const double Q = block_sum_d(q, red);
  if (threadIdx.x == 0) {
    double v = Q - S * S / (double)P;
    if (v < 0.0) v = 0.0;
    mean[blockIdx.x] = (float)(K + S / (double)P);
    m2[blockIdx.x] = (float)v;
    if (in_a) {
      in_a[blockIdx.x] = 1.f / sqrtf((float)v / (float)P + eps);
      in_b[blockIdx.x] = 0.f;
    }
  }
}

__global__ void in_finalize_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ m2,
                                       float* __restrict__ a, float* __restrict__ b, int planes, int P, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= planes) return;
  const float var = m2[i] / (float)P;
  const float rstd = 1.f / sqrtf(var + eps);
  a[i] = rstd;
  b[i] = 0.f;  // out = rstd * (y - mean)
}

// InstanceNorm statistics from the per-(n, c) sums the conv epilogue accumulated (san_tc_conv_stats): sums[plane*group + k][2]
// = (sum, sum of squares) of sub-plane k of a plane (group = 4 for the pixel-shuffled transposed conv, whose four
// sub-planes normalise together; 1 otherwise), P elements per sub-plane.  Same outputs as plane_stats_kernel + finalise.
__global__ void in_stats_from_sums_kernel(const double* __restrict__ sums, float* __restrict__ mean, float* __restrict__ m2,
                                          float* __restrict__ a, float* __restrict__ b, int planes, int group, int P, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= planes) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < group; ++k) {
    s1 += sums[2 * ((long long)i * group + k)];
    s2 += sums[2 * ((long long)i * group + k) + 1];
  }
  const double cnt = (double)group * P;
  const double mu = s1 / cnt;
  double M2 = s2 - s1 * mu;
  if (M2 < 0.0) M2 = 0.0;
  mean[i] = (float)mu;
  m2[i] = (float)M2;
  a[i] = (float)(1.0 / sqrt(M2 / cnt + (double)eps));
  b[i] = 0.f;
}

// one thread per channel
__global__ void bn_finalize_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ m2,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* running_mean, float* running_var, float* __restrict__ mu_out,
                                       float* __restrict__ a, float* __restrict__ b, float* __restrict__ sa, int N,
                                       int C, int P, float eps, float momentum, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mu, rstd;
  if (training) {
    double sm = 0.0;
    for (int n = 0; n < N; ++n) sm += (double)mean[n * C + c];
    mu = sm / N;
    double M2 = 0.0;
    for (int n = 0; n < N; ++n) {
      const double d = (double)mean[n * C + c] - mu;
      M2 += (double)m2[n * C + c] + d * d * P;
    }
    const double cnt = (double)N * P;
    const double var = M2 / cnt;
    rstd = 1.0 / sqrt(var + (double)eps);
    if (running_mean) {
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
      const double unb = cnt > 1 ? M2 / (cnt - 1.0) : var;
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
  } else {
    mu = running_mean[c];
    rstd = 1.0 / sqrt((double)running_var[c] + (double)eps);
  }
  // out = gamma*rstd * (y - mu) + beta ; xhat = rstd * (y - mu)
  const float fa = (float)(gamma[c] * rstd), fb = beta[c], fsa = (float)rstd, fmu = (float)mu;
  for (int n = 0; n < N; ++n) {
    mu_out[n * C + c] = fmu; a[n * C + c] = fa; b[n * C + c] = fb; sa[n * C + c] = fsa;
  }
}

// ---- forward apply ------------------------------------------------------------
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// grid (planes, chunks)
__global__ void __launch_bounds__(256) affine_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ mu,
                                                             const float* __restrict__ a,
                                                             const float* __restrict__ b, float slope,
                                                             float* __restrict__ out, int P) {
  const long long base = (long long)blockIdx.x * P;
  const float cm = mu ? mu[blockIdx.x] : 0.f;
  const float ca = a[blockIdx.x], cb = b ? b[blockIdx.x] : 0.f;
  if ((P & 3) == 0) {
    const float4* y4 = (const float4*)(y + base);
    float4* o4 = (float4*)(out + base);
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P / 4; i += gridDim.y * blockDim.x) {
      float4 v = y4[i];
      v.x = lrelu(fmaf(ca, v.x - cm, cb), slope); v.y = lrelu(fmaf(ca, v.y - cm, cb), slope);
      v.z = lrelu(fmaf(ca, v.z - cm, cb), slope); v.w = lrelu(fmaf(ca, v.w - cm, cb), slope);
      o4[i] = v;
    }
  } else {
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P; i += gridDim.y * blockDim.x)
      out[base + i] = lrelu(fmaf(ca, y[base + i] - cm, cb), slope);
  }
}

// ---- backward -------------------------------------------------------------------
// s1 = sum g', s2 = sum g' * sa*(y - mu), g' = g * lrelu'(a*(y - mu) + b); one block per plane
__global__ void __launch_bounds__(256) act_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                             const float* __restrict__ mu,
                                                             const float* __restrict__ a, const float* __restrict__ b,
                                                             const float* __restrict__ sa,
                                                             float slope, float* __restrict__ s1, float* __restrict__ s2,
                                                             int P) {
  __shared__ double red[32];
  const long long base = (long long)blockIdx.x * P;
  const float cm = mu ? mu[blockIdx.x] : 0.f;
  const float ca = a[blockIdx.x], cb = b ? b[blockIdx.x] : 0.f;
  const float csa = sa ? sa[blockIdx.x] : 1.f;
  float t1 = 0.f, t2 = 0.f;
  if ((P & 3) == 0) {
    // 16 B loads, two independent float4 pairs in flight per thread
    const float4* y4 = (const float4*)(y + base);
    const float4* g4 = (const float4*)(g + base);
    const int n4 = P >> 2;
    for (int i = threadIdx.x; i < n4; i += 2 * blockDim.x) {
      const int i2 = i + blockDim.x;
      const bool has2 = i2 < n4;
      const float4 ya = __ldg(y4 + i), ga = __ldg(g4 + i);
      const float4 yb = has2 ? __ldg(y4 + i2) : make_float4(cm, cm, cm, cm);
      const float4 gb = has2 ? __ldg(g4 + i2) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
      const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        t1 += gg;
        t2 += gg * (csa * yc);
      }
    }
  } else {
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      const float yc = y[base + i] - cm;
      float gg = g[base + i];
      if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
      t1 += gg;
      t2 += gg * (csa * yc);
    }
  }
  const double r1 = block_sum_d((double)t1, red);
  const double r2 = block_sum_d((double)t2, red);
  if (threadIdx.x == 0) { s1[blockIdx.x] = (float)r1; s2[blockIdx.x] = (float)r2; }
}

// ---- the same two passes reading the incoming gradient IN PLACE from the data gradient of the consuming
// conv (dx [N, Ctot, Hd, Wd]): channel offset c0 of a concatenated source, batch stride, and the adjoint of
// the resampling the consumer applied while staging (MODE 0 direct, 1 avg-pool 2x2, 2 pixel shuffle, 3 nearest
// x2).  Replaces slice copies + up2 / space_to_depth2 / pool2 kernels + their intermediate tensors.
struct GMap {
  const float* g;     // dx of the consumer
  long long g_bs;     // floats per image of dx (Ctot * Hd * Wd)
  int c0, Cy;         // first dx channel of this term, channels of the term (planes per image)
  int Hy, Wy;         // spatial size of one y (sub-)plane: MODE 2: y is [N, 4*Cy, Hy, Wy]
};
template <int MODE>
__device__ __forceinline__ float gmap_load(const float* __restrict__ gp, int e, int Hy, int Wy) {
  if (MODE == 0) return __ldg(gp + e);
  if (MODE == 1) {              // y plane Hy x Wy, dx plane Hy/2 x Wy/2
    const int h = e / Wy, w = e - h * Wy;
    return 0.25f * __ldg(gp + (h >> 1) * (Wy >> 1) + (w >> 1));
  }
  if (MODE == 2) {              // y planes 4 x (Hy x Wy), dx plane 2Hy x 2Wy
    const int hw = Hy * Wy;
    const int sub = e / hw, r = e - sub * hw;
    const int h = r / Wy, w = r - h * Wy;
    return __ldg(gp + (2 * h + (sub >> 1)) * (2 * Wy) + 2 * w + (sub & 1));
  }
  const int h = e / Wy, w = e - h * Wy;   // MODE 3: y plane Hy x Wy, dx plane 2Hy x 2Wy
  const float* q = gp + (2 * h) * (2 * Wy) + 2 * w;
  const float2 r0 = __ldg((const float2*)q), r1 = __ldg((const float2*)(q + 2 * Wy));
  return (r0.x + r0.y) + (r1.x + r1.y);
}
template <int MODE>
__device__ __forceinline__ const float* gmap_plane(const GMap& m, int plane) {
  const int n = plane / m.Cy, c = plane - n * m.Cy;
  const long long hwd = MODE == 0 ? (long long)m.Hy * m.Wy : MODE == 1 ? (long long)(m.Hy >> 1) * (m.Wy >> 1) : 4LL * m.Hy * m.Wy;
  return m.g + (long long)n * m.g_bs + (long long)(m.c0 + c) * hwd;
}

template <int MODE>
__global__ void __launch_bounds__(256) act_bwd_reduce_map_kernel(const GMap m, const float* __restrict__ y,
                                                                 const float* __restrict__ mu, const float* __restrict__ a,
                                                                 const float* __restrict__ b, const float* __restrict__ sa,
                                                                 float slope, float* __restrict__ s1, float* __restrict__ s2,
                                                                 int P) {
  __shared__ double red[32];
  const long long base = (long long)blockIdx.x * P;
  const float* gp = gmap_plane<MODE>(m, blockIdx.x);
  const float cm = mu ? mu[blockIdx.x] : 0.f;
  const float ca = a[blockIdx.x], cb = b ? b[blockIdx.x] : 0.f;
  const float csa = sa ? sa[blockIdx.x] : 1.f;
  float t1 = 0.f, t2 = 0.f;
  if (MODE == 0 && (P & 3) == 0) {
    const float4* y4 = (const float4*)(y + base);
    const float4* g4 = (const float4*)gp;
    const int n4 = P >> 2;
    for (int i = threadIdx.x; i < n4; i += 2 * blockDim.x) {
      const int i2 = i + blockDim.x;
      const bool has2 = i2 < n4;
      const float4 ya = __ldg(y4 + i), ga = __ldg(g4 + i);
      const float4 yb = has2 ? __ldg(y4 + i2) : make_float4(cm, cm, cm, cm);
      const float4 gb = has2 ? __ldg(g4 + i2) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
      const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        t1 += gg;
        t2 += gg * (csa * yc);
      }
    }
  } else if (MODE == 1 && (m.Wy & 3) == 0) {
    // avg-pool adjoint, vectorised: 4 consecutive y pixels of a row read 2 consecutive dx pixels
    const float4* y4 = (const float4*)(y + base);
    const int n4 = P >> 2, w4 = m.Wy >> 2, wd = m.Wy >> 1;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const int h = i / w4, wq = i - h * w4;
      const float4 yv4 = __ldg(y4 + i);
      const float2 g2 = __ldg((const float2*)(gp + (h >> 1) * wd + 2 * wq));
      const float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
      const float gv[4] = {0.25f * g2.x, 0.25f * g2.x, 0.25f * g2.y, 0.25f * g2.y};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        t1 += gg;
        t2 += gg * (csa * yc);
      }
    }
  } else if (MODE == 2 && (m.Wy & 3) == 0) {
    // pixel-shuffle adjoint, vectorised: 4 consecutive pixels of sub-plane (a, b) read every other one of 8
    // consecutive dx pixels of row 2h + a
    const float4* y4 = (const float4*)(y + base);
    const int n4 = P >> 2, w4 = m.Wy >> 2, q4 = m.Hy * w4, wd = 2 * m.Wy;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const int sub = i / q4, r = i - sub * q4;
      const int h = r / w4, wq = r - h * w4;
      const float4 yv4 = __ldg(y4 + i);
      const float4* gq = (const float4*)(gp + (2 * h + (sub >> 1)) * wd + 8 * wq);
      const float4 ga = __ldg(gq), gb = __ldg(gq + 1);
      const float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
      const float gv[4] = {(sub & 1) ? ga.y : ga.x, (sub & 1) ? ga.w : ga.z, (sub & 1) ? gb.y : gb.x, (sub & 1) ? gb.w : gb.z};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        t1 += gg;
        t2 += gg * (csa * yc);
      }
    }
  } else {
    for (int i = threadIdx.x; i < P; i += 2 * blockDim.x) {
      const int i2 = i + blockDim.x;
      const bool has2 = i2 < P;
      const float ya = __ldg(y + base + i), ga = gmap_load<MODE>(gp, i, m.Hy, m.Wy);
      const float yb = has2 ? __ldg(y + base + i2) : cm, gb = has2 ? gmap_load<MODE>(gp, i2, m.Hy, m.Wy) : 0.f;
      float yc = ya - cm, gg = ga;
      if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
      t1 += gg; t2 += gg * (csa * yc);
      yc = yb - cm; gg = gb;
      if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
      t1 += gg; t2 += gg * (csa * yc);
    }
  }
  const double r1 = block_sum_d((double)t1, red);
  const double r2 = block_sum_d((double)t2, red);
  if (threadIdx.x == 0) { s1[blockIdx.x] = (float)r1; s2[blockIdx.x] = (float)r2; }
}

// dy = p * g' + q * (y - mu) + r with g read through the map; grid (planes, chunks)
template <int MODE>
__global__ void __launch_bounds__(256) act_bwd_apply_map_kernel(const GMap m, const float* __restrict__ y,
                                                                const float* __restrict__ mu, const float* __restrict__ a,
                                                                const float* __restrict__ b, float slope,
                                                                const float* __restrict__ p, const float* __restrict__ q,
                                                                const float* __restrict__ r, float* __restrict__ dy, int P,
                                                                unsigned int* __restrict__ absmax) {
  const long long base = (long long)blockIdx.x * P;
  const float* gp = gmap_plane<MODE>(m, blockIdx.x);
  const float cm = mu ? mu[blockIdx.x] : 0.f;
  const float ca = a[blockIdx.x], cb = b ? b[blockIdx.x] : 0.f;
  const float cp = p[blockIdx.x], cq = q ? q[blockIdx.x] : 0.f, cr = r ? r[blockIdx.x] : 0.f;
  const int stride = gridDim.y * blockDim.x;
  float am = 0.f;       // max |dy| of this thread (optional output: the dynamic fp16-pair scale of the consumer, tc_common.cuh)
  if (MODE == 0 && (P & 3) == 0) {
    const float4* y4 = (const float4*)(y + base);
    const float4* g4 = (const float4*)gp;
    float4* d4 = (float4*)(dy + base);
    const int n4 = P >> 2;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
      const int i2 = i + stride;
      const bool has2 = i2 < n4;
      const float4 ya = __ldg(y4 + i), ga = __ldg(g4 + i);
      float4 yb = ya, gb = ga;
      if (has2) { yb = __ldg(y4 + i2); gb = __ldg(g4 + i2); }
      float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
      float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        yv[k] = fmaf(cp, gg, fmaf(cq, yc, cr));
        am = fmaxf(am, fabsf(yv[k]));       // (the duplicated tail values of !has2 are real elements: harmless)
      }
      d4[i] = make_float4(yv[0], yv[1], yv[2], yv[3]);
      if (has2) d4[i2] = make_float4(yv[4], yv[5], yv[6], yv[7]);
    }
  } else if (MODE == 1 && (m.Wy & 3) == 0) {
    const float4* y4 = (const float4*)(y + base);
    float4* d4 = (float4*)(dy + base);
    const int n4 = P >> 2, w4 = m.Wy >> 2, wd = m.Wy >> 1;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int h = i / w4, wq = i - h * w4;
      const float4 yv4 = __ldg(y4 + i);
      const float2 g2 = __ldg((const float2*)(gp + (h >> 1) * wd + 2 * wq));
      float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
      const float gv[4] = {0.25f * g2.x, 0.25f * g2.x, 0.25f * g2.y, 0.25f * g2.y};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        yv[k] = fmaf(cp, gg, fmaf(cq, yc, cr));
        am = fmaxf(am, fabsf(yv[k]));
      }
      d4[i] = make_float4(yv[0], yv[1], yv[2], yv[3]);
    }
  } else if (MODE == 2 && (m.Wy & 3) == 0) {
    const float4* y4 = (const float4*)(y + base);
    float4* d4 = (float4*)(dy + base);
    const int n4 = P >> 2, w4 = m.Wy >> 2, q4 = m.Hy * w4, wd = 2 * m.Wy;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int sub = i / q4, r = i - sub * q4;
      const int h = r / w4, wq = r - h * w4;
      const float4 yv4 = __ldg(y4 + i);
      const float4* gq = (const float4*)(gp + (2 * h + (sub >> 1)) * wd + 8 * wq);
      const float4 ga = __ldg(gq), gb = __ldg(gq + 1);
      float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
      const float gv[4] = {(sub & 1) ? ga.y : ga.x, (sub & 1) ? ga.w : ga.z, (sub & 1) ? gb.y : gb.x, (sub & 1) ? gb.w : gb.z};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        yv[k] = fmaf(cp, gg, fmaf(cq, yc, cr));
        am = fmaxf(am, fabsf(yv[k]));
      }
      d4[i] = make_float4(yv[0], yv[1], yv[2], yv[3]);
    }
  } else {
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P; i += 2 * stride) {
      const int i2 = i + stride;
      const bool has2 = i2 < P;
      const float ya = __ldg(y + base + i), ga = gmap_load<MODE>(gp, i, m.Hy, m.Wy);
      const float yb = has2 ? __ldg(y + base + i2) : cm, gb = has2 ? gmap_load<MODE>(gp, i2, m.Hy, m.Wy) : 0.f;
      float yc = ya - cm, gg = ga;
      if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
      float o = fmaf(cp, gg, fmaf(cq, yc, cr));
      dy[base + i] = o;
      am = fmaxf(am, fabsf(o));
      if (has2) {
        yc = yb - cm; gg = gb;
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        o = fmaf(cp, gg, fmaf(cq, yc, cr));
        dy[base + i2] = o;
        am = fmaxf(am, fabsf(o));
      }
    }
  }
  if (absmax) {      // one atomicMax per warp on the uint view (non-negative floats order like their bit patterns)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
    if ((threadIdx.x & 31) == 0 && am < 3.0e38f) atomicMax(absmax, __float_as_uint(am));
  }
}

// ---- InstanceNorm backward of one plane in ONE kernel: reduce -> coefficients -> apply (act_bwd_reduce_map +
// in_finalize_bwd + act_bwd_apply_map).  One CTA per (n, c) plane.  The second pass walks the plane in REVERSE order,
// most recently read data first, so that its reads of y and dx are served by the 126 MB L2 wherever the planes of the
// resident CTAs fit (always below 320x320; partly there): 3 HBM passes (y, dx in; dy out) instead of 5.
// 4 consecutive y pixels + their gradients through the map (vector paths need Wy % 4 == 0, else scalar fall-back).
template <int MODE>
__device__ __forceinline__ void fused_load4(const GMap& m, const float* __restrict__ yb, const float* __restrict__ gp, int i,
                                            float* yv, float* gv) {
  const float4 y4 = __ldg((const float4*)yb + i);
  yv[0] = y4.x; yv[1] = y4.y; yv[2] = y4.z; yv[3] = y4.w;
  if (MODE == 0) {
    const float4 g4 = __ldg((const float4*)gp + i);
    gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
  } else if (MODE == 1) {
    const int w4 = m.Wy >> 2, h = i / w4, wq = i - h * w4;
    const float2 g2 = __ldg((const float2*)(gp + (h >> 1) * (m.Wy >> 1) + 2 * wq));
    gv[0] = 0.25f * g2.x; gv[1] = gv[0]; gv[2] = 0.25f * g2.y; gv[3] = gv[2];
  } else if (MODE == 2) {
    const int w4 = m.Wy >> 2, q4 = m.Hy * w4;
    const int sub = i / q4, r = i - sub * q4, h = r / w4, wq = r - h * w4;
    const float4* gq = (const float4*)(gp + (2 * h + (sub >> 1)) * (2 * m.Wy) + 8 * wq);
    const float4 ga = __ldg(gq), gb = __ldg(gq + 1);
    gv[0] = (sub & 1) ? ga.y : ga.x; gv[1] = (sub & 1) ? ga.w : ga.z; gv[2] = (sub & 1) ? gb.y : gb.x; gv[3] = (sub & 1) ? gb.w : gb.z;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) gv[k] = gmap_load<3>(gp, 4 * i + k, m.Hy, m.Wy);
  }
}

// CLUSTER: the plane is split over the two CTAs of a thread-block cluster, which exchange their partial sums through
// distributed shared memory.  At 320 x 320 a plane with its gradient is 820 KB: one 1024-thread CTA per SM keeps 148 planes
// = 121 MB live, right at the L2 capacity, and half of the second pass came from HBM again (ncu: 1.47 GB read for 0.94 GB
// algorithmic); two CTAs per plane halve the live set (74 planes = 61 MB).
template <int MODE, bool CLUSTER>
__device__ __forceinline__ void in_bwd_fused_map_body(const GMap& m, const float* __restrict__ y, const float* __restrict__ mu,
                                                      const float* __restrict__ a, float slope, float* __restrict__ dy, int P,
                                                      unsigned int* __restrict__ absmax) {
  __shared__ double red[32];
  __shared__ double part[2];
  __shared__ float coef[3];
  const int plane = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int rank = CLUSTER ? (int)(blockIdx.x & 1) : 0;
  const long long base = (long long)plane * P;
  const float* gp = gmap_plane<MODE>(m, plane);
  const float* yb = y + base;
  const float cm = mu[plane], ca = a[plane];
  const bool vec = (P & 3) == 0 && (m.Wy & 3) == 0;
  const int n4 = P >> 2;
  // this CTA's share of the plane: float4 groups [q0, q1) (vector path) or elements [e0, e1)
  const int q0 = CLUSTER ? (int)((long long)n4 * rank / 2) : 0, q1 = CLUSTER ? (int)((long long)n4 * (rank + 1) / 2) : n4;
  const int e0 = CLUSTER ? (int)((long long)P * rank / 2) : 0, e1 = CLUSTER ? (int)((long long)P * (rank + 1) / 2) : P;
  float t1 = 0.f, t2 = 0.f;
  if (vec) {
    for (int i = q0 + threadIdx.x; i < q1; i += blockDim.x) {
      float yv[4], gv[4];
      fused_load4<MODE>(m, yb, gp, i, yv, gv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (ca * yc <= 0.f) gg *= slope;
        t1 += gg;
        t2 += gg * (ca * yc);
      }
    }
  } else {
    for (int i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
      const float yc = __ldg(yb + i) - cm;
      float gg = gmap_load<MODE>(gp, i, m.Hy, m.Wy);
      if (ca * yc <= 0.f) gg *= slope;
      t1 += gg;
      t2 += gg * (ca * yc);
    }
  }
  double r1 = block_sum_d((double)t1, red);
  double r2 = block_sum_d((double)t2, red);
  if (CLUSTER) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (threadIdx.x == 0) { part[0] = r1; part[1] = r2; }
    cluster.sync();
    if (threadIdx.x == 0) {
      const double* peer = cluster.map_shared_rank(part, rank ^ 1);
      const double p1 = peer[0], p2 = peer[1];
      // rank 0's partial first in both CTAs: the two halves form bit-identical coefficients
      r1 = rank == 0 ? r1 + p1 : p1 + r1;
      r2 = rank == 0 ? r2 + p2 : p2 + r2;
    }
    cluster.sync();                 // the peer has read `part` before this CTA may exit
  }
  if (threadIdx.x == 0) {           // in_finalize_bwd: dy = p g' + q (y - mu) + r
    const double A = ca, M = P;
    coef[0] = (float)A;
    coef[1] = (float)(-A * A * (double)(float)r2 / M);
    coef[2] = (float)(-A * (double)(float)r1 / M);
  }
  __syncthreads();
  const float cp = coef[0], cq = coef[1], cr = coef[2];
  float am = 0.f;
  if (vec) {
    float4* d4 = (float4*)(dy + base);
    for (int k0 = threadIdx.x; k0 < q1 - q0; k0 += blockDim.x) {
      const int i = q1 - 1 - k0;      // reverse: what the first pass read last is still in L2
      float yv[4], gv[4];
      fused_load4<MODE>(m, yb, gp, i, yv, gv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (ca * yc <= 0.f) gg *= slope;
        yv[k] = fmaf(cp, gg, fmaf(cq, yc, cr));
        am = fmaxf(am, fabsf(yv[k]));
      }
      d4[i] = make_float4(yv[0], yv[1], yv[2], yv[3]);
    }
  } else {
    for (int k0 = threadIdx.x; k0 < e1 - e0; k0 += blockDim.x) {
      const int i = e1 - 1 - k0;
      const float yc = __ldg(yb + i) - cm;
      float gg = gmap_load<MODE>(gp, i, m.Hy, m.Wy);
      if (ca * yc <= 0.f) gg *= slope;
      const float o = fmaf(cp, gg, fmaf(cq, yc, cr));
      dy[base + i] = o;
      am = fmaxf(am, fabsf(o));
    }
  }
  if (absmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
    if ((threadIdx.x & 31) == 0 && am < 3.0e38f) atomicMax(absmax, __float_as_uint(am));
  }
}

template <int MODE>
__global__ void __launch_bounds__(1024) in_bwd_fused_map_kernel(const GMap m, const float* __restrict__ y,
                                                                const float* __restrict__ mu, const float* __restrict__ a,
                                                                float slope, float* __restrict__ dy, int P,
                                                                unsigned int* __restrict__ absmax) {
  in_bwd_fused_map_body<MODE, false>(m, y, mu, a, slope, dy, P, absmax);
}
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024)
in_bwd_fused_map_cl2_kernel(const GMap m, const float* __restrict__ y, const float* __restrict__ mu,
                            const float* __restrict__ a, float slope, float* __restrict__ dy, int P,
                            unsigned int* __restrict__ absmax) {
  in_bwd_fused_map_body<MODE, true>(m, y, mu, a, slope, dy, P, absmax);
}

// dx = rstd * (g' - mean(g') - xhat * mean(g' xhat)), xhat = rstd*(y - mu)  ->  dy = p g' + q (y - mu) + r
__global__ void in_finalize_bwd_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                       const float* __restrict__ a, float* __restrict__ p,
                                       float* __restrict__ q, float* __restrict__ r, int planes, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= planes) return;
  const double A = a[i], M = P;
  p[i] = (float)A;
  q[i] = (float)(-A * A * (double)s2[i] / M);
  r[i] = (float)(-A * (double)s1[i] / M);
}

__global__ void bn_finalize_bwd_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                       const float* __restrict__ gamma, const float* __restrict__ sa,
                                       float* __restrict__ p, float* __restrict__ q,
                                       float* __restrict__ r, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       int N, int C, int P, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double S1 = 0.0, S2 = 0.0;
  for (int n = 0; n < N; ++n) { S1 += (double)s1[n * C + c]; S2 += (double)s2[n * C + c]; }
  dgamma[c] = (float)S2;
  dbeta[c] = (float)S1;
  const double rstd = sa[c];  // identical over n; row 0
  const double G = gamma[c], Mc = (double)N * P;
  float fp, fq, fr;
  if (training) {
    fp = (float)(G * rstd);
    fq = (float)(-G * rstd * rstd * S2 / Mc);   // multiplies (y - mu)
    fr = (float)(-G * rstd * S1 / Mc);
  } else {
    fp = (float)(G * rstd); fq = 0.f; fr = 0.f;
  }
  for (int n = 0; n < N; ++n) { p[n * C + c] = fp; q[n * C + c] = fq; r[n * C + c] = fr; }
}

// dy = p * g' + q * (y - mu) + r ; grid (planes, chunks)
__global__ void __launch_bounds__(256) act_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                            const float* __restrict__ mu,
                                                            const float* __restrict__ a, const float* __restrict__ b,
                                                            float slope, const float* __restrict__ p,
                                                            const float* __restrict__ q, const float* __restrict__ r,
                                                            float* __restrict__ dy, int P) {
  const long long base = (long long)blockIdx.x * P;
  const float cm = mu ? mu[blockIdx.x] : 0.f;
  const float ca = a[blockIdx.x], cb = b ? b[blockIdx.x] : 0.f;
  const float cp = p[blockIdx.x], cq = q ? q[blockIdx.x] : 0.f, cr = r ? r[blockIdx.x] : 0.f;
  if ((P & 3) == 0) {
    const float4* y4 = (const float4*)(y + base);
    const float4* g4 = (const float4*)(g + base);
    float4* d4 = (float4*)(dy + base);
    const int n4 = P >> 2, stride = gridDim.y * blockDim.x;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
      const int i2 = i + stride;
      const bool has2 = i2 < n4;
      const float4 ya = __ldg(y4 + i), ga = __ldg(g4 + i);
      float4 yb = ya, gb = ga;
      if (has2) { yb = __ldg(y4 + i2); gb = __ldg(g4 + i2); }
      float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
      float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yc = yv[k] - cm;
        float gg = gv[k];
        if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
        yv[k] = fmaf(cp, gg, fmaf(cq, yc, cr));
      }
      d4[i] = make_float4(yv[0], yv[1], yv[2], yv[3]);
      if (has2) d4[i2] = make_float4(yv[4], yv[5], yv[6], yv[7]);
    }
    return;
  }
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P; i += gridDim.y * blockDim.x) {
    const float yc = y[base + i] - cm;
    float gg = g[base + i];
    if (fmaf(ca, yc, cb) <= 0.f) gg *= slope;
    dy[base + i] = fmaf(cp, gg, fmaf(cq, yc, cr));
  }
}

// ---- resampling -----------------------------------------------------------------
// y[h,w] = scale * sum of the 2x2 block of x; x planes are H x W, y planes Ho x Wo
__global__ void pool2_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int Ho, int Wo,
                             float scale, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const long long t = i / Wo;
    const int ho = (int)(t % Ho);
    const long long pl = t / Ho;
    const float* p = x + (pl * H + 2 * ho) * W + 2 * wo;
    y[i] = scale * ((p[0] + p[1]) + (p[W] + p[W + 1]));
  }
}

// y[2h+a, 2w+b] = scale * x[h,w]; total = planes * 2H * 2W
__global__ void up2_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, float scale,
                           long long total) {
  const int Wo = 2 * W, Ho = 2 * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const long long t = i / Wo;
    const int ho = (int)(t % Ho);
    const long long pl = t / Ho;
    y[i] = scale * x[(pl * H + (ho >> 1)) * W + (wo >> 1)];
  }
}

// x [N, Co*4, H, W] (channel = co*4 + a*2 + b)  <->  y [N, Co, 2H, 2W]
template <bool TO_SPACE>
__global__ void shuffle2_kernel(const float* __restrict__ src, float* __restrict__ dst, int Co, int H, int W,
                                long long total) {
  const int Wo = 2 * W, Ho = 2 * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i enumerates the spatial-layout tensor [N, Co, 2H, 2W]
    const int wo = (int)(i % Wo);
    long long t = i / Wo;
    const int ho = (int)(t % Ho);
    t /= Ho;
    const int co = (int)(t % Co);
    const long long n = t / Co;
    const long long j = ((n * Co * 4 + co * 4 + (ho & 1) * 2 + (wo & 1)) * H + (ho >> 1)) * W + (wo >> 1);
    if (TO_SPACE) dst[i] = src[j]; else dst[j] = src[i];
  }
}

inline int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)san_num_sms() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}
inline int chunks_for(int planes, int P) {
  // enough blocks to fill the machine, at most one block per 1024 elements
  int per = (P + 1023) / 1024;
  long long want = ((long long)san_num_sms() * 8 + planes - 1) / planes;
  int c = (int)(want < per ? want : per);
  return c < 1 ? 1 : c;
}

}  // namespace

extern "C" {

int san_plane_stats(const float* x, float* mean, float* m2, int planes, int P, void* stream) {
  SAN_CHECK_ARG(x && mean && m2 && planes > 0 && P > 0, "san_plane_stats: bad args");
  plane_stats_kernel<<<planes, 256, 0, (cudaStream_t)stream>>>(x, mean, m2, P, nullptr, nullptr, 0.f);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_plane_stats_in(const float* x, float* mean, float* m2, float* a, float* b, int planes, int P, float eps,
                       void* stream) {
  SAN_CHECK_ARG(x && mean && m2 && a && b && planes > 0 && P > 0, "san_plane_stats_in: bad args");
  plane_stats_kernel<<<planes, 256, 0, (cudaStream_t)stream>>>(x, mean, m2, P, a, b, eps);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_in_stats_from_sums(const double* sums, float* mean, float* m2, float* a, float* b, int planes, int group, int P,
                           float eps, void* stream) {
  SAN_CHECK_ARG(sums && mean && m2 && a && b && planes > 0 && group > 0 && P > 0, "san_in_stats_from_sums: bad args");
  in_stats_from_sums_kernel<<<san_cdiv(planes, 256), 256, 0, (cudaStream_t)stream>>>(sums, mean, m2, a, b, planes, group, P, eps);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_in_finalize_fwd(const float* mean, const float* m2, float* a, float* b, int planes, int P, float eps,
                        void* stream) {
  SAN_CHECK_ARG(mean && m2 && a && b && planes > 0, "san_in_finalize_fwd: bad args");
  in_finalize_fwd_kernel<<<san_cdiv(planes, 256), 256, 0, (cudaStream_t)stream>>>(mean, m2, a, b, planes, P, eps);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_bn_finalize_fwd(const float* mean, const float* m2, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float* mu, float* a, float* b, float* sa, int N,
                        int C, int P, float eps, float momentum, int training, void* stream) {
  SAN_CHECK_ARG(gamma && beta && mu && a && b && sa && N > 0 && C > 0, "san_bn_finalize_fwd: bad args");
  SAN_CHECK_ARG(training ? (mean && m2) : (running_mean && running_var), "san_bn_finalize_fwd: missing statistics");
  bn_finalize_fwd_kernel<<<san_cdiv(C, 64), 64, 0, (cudaStream_t)stream>>>(mean, m2, gamma, beta, running_mean,
                                                                           running_var, mu, a, b, sa, N, C, P, eps,
                                                                           momentum, training);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_affine_act_fwd(const float* y, const float* mu, const float* a, const float* b, float slope, float* out,
                       int planes, int P, void* stream) {
  SAN_CHECK_ARG(y && a && out && planes > 0 && P > 0, "san_affine_act_fwd: bad args");
  dim3 grid(planes, chunks_for(planes, P));
  affine_act_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, mu, a, b, slope, out, P);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_act_bwd_reduce(const float* g, const float* y, const float* mu, const float* a, const float* b,
                       const float* sa, float slope, float* s1, float* s2, int planes, int P, void* stream) {
  SAN_CHECK_ARG(g && y && a && s1 && s2 && planes > 0 && P > 0, "san_act_bwd_reduce: bad args");
  act_bwd_reduce_kernel<<<planes, 256, 0, (cudaStream_t)stream>>>(g, y, mu, a, b, sa, slope, s1, s2, P);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_in_finalize_bwd(const float* s1, const float* s2, const float* a, float* p, float* q,
                        float* r, int planes, int P, void* stream) {
  SAN_CHECK_ARG(s1 && s2 && a && p && q && r && planes > 0, "san_in_finalize_bwd: bad args");
  in_finalize_bwd_kernel<<<san_cdiv(planes, 256), 256, 0, (cudaStream_t)stream>>>(s1, s2, a, p, q, r, planes, P);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_bn_finalize_bwd(const float* s1, const float* s2, const float* gamma, const float* sa,
                        float* p, float* q, float* r, float* dgamma, float* dbeta, int N, int C, int P, int training,
                        void* stream) {
  SAN_CHECK_ARG(s1 && s2 && gamma && sa && p && q && r && dgamma && dbeta, "san_bn_finalize_bwd: bad args");
  bn_finalize_bwd_kernel<<<san_cdiv(C, 64), 64, 0, (cudaStream_t)stream>>>(s1, s2, gamma, sa, p, q, r, dgamma,
                                                                           dbeta, N, C, P, training);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_act_bwd_apply(const float* g, const float* y, const float* mu, const float* a, const float* b, float slope,
                      const float* p, const float* q, const float* r, float* dy, int planes, int P, void* stream) {
  SAN_CHECK_ARG(g && y && a && p && dy && planes > 0 && P > 0, "san_act_bwd_apply: bad args");
  dim3 grid(planes, chunks_for(planes, P));
  act_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, y, mu, a, b, slope, p, q, r, dy, P);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

static int gmap_check(const float* g, int N, int Ctot, int c0, int Cy, int Hy, int Wy, int mode, const char* who) {
  SAN_CHECK_ARG(g && N > 0 && Ctot > 0 && c0 >= 0 && Cy > 0 && c0 + Cy <= Ctot && Hy > 0 && Wy > 0 && mode >= 0 && mode <= 3,
                "%s: bad gradient map", who);
  if (mode == 1) SAN_CHECK_ARG(Hy % 2 == 0 && Wy % 2 == 0, "%s: pooled source needs even size", who);
  return SAN_OK;
}
static GMap make_gmap(const float* g, int Ctot, int c0, int Cy, int Hy, int Wy, int mode) {
  const long long hwd = mode == 0 ? (long long)Hy * Wy : mode == 1 ? (long long)(Hy / 2) * (Wy / 2) : 4LL * Hy * Wy;
  return GMap{g, (long long)Ctot * hwd, c0, Cy, Hy, Wy};
}

int san_act_bwd_reduce_map(const float* g, int Ctot, int c0, int mode, const float* y, const float* mu, const float* a,
                           const float* b, const float* sa, float slope, float* s1, float* s2, int N, int Cy, int Hy,
                           int Wy, void* stream) {
  SAN_CHECK_ARG(y && a && s1 && s2, "san_act_bwd_reduce_map: bad args");
  int rc = gmap_check(g, N, Ctot, c0, Cy, Hy, Wy, mode, "san_act_bwd_reduce_map");
  if (rc != SAN_OK) return rc;
  const GMap m = make_gmap(g, Ctot, c0, Cy, Hy, Wy, mode);
  const int planes = N * Cy, P = (mode == 2 ? 4 : 1) * Hy * Wy;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case 0: act_bwd_reduce_map_kernel<0><<<planes, 256, 0, st>>>(m, y, mu, a, b, sa, slope, s1, s2, P); break;
    case 1: act_bwd_reduce_map_kernel<1><<<planes, 256, 0, st>>>(m, y, mu, a, b, sa, slope, s1, s2, P); break;
    case 2: act_bwd_reduce_map_kernel<2><<<planes, 256, 0, st>>>(m, y, mu, a, b, sa, slope, s1, s2, P); break;
    default: act_bwd_reduce_map_kernel<3><<<planes, 256, 0, st>>>(m, y, mu, a, b, sa, slope, s1, s2, P); break;
  }
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_act_bwd_apply_map(const float* g, int Ctot, int c0, int mode, const float* y, const float* mu, const float* a,
                          const float* b, float slope, const float* p, const float* q, const float* r, float* dy, int N,
                          int Cy, int Hy, int Wy, float* absmax, void* stream) {
  SAN_CHECK_ARG(y && a && p && dy, "san_act_bwd_apply_map: bad args");
  int rc = gmap_check(g, N, Ctot, c0, Cy, Hy, Wy, mode, "san_act_bwd_apply_map");
  if (rc != SAN_OK) return rc;
  const GMap m = make_gmap(g, Ctot, c0, Cy, Hy, Wy, mode);
  const int planes = N * Cy, P = (mode == 2 ? 4 : 1) * Hy * Wy;
  dim3 grid(planes, chunks_for(planes, P));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* am = (unsigned int*)absmax;
  if (am) SAN_CUDA(cudaMemsetAsync(am, 0, sizeof(float), st));
  switch (mode) {
    case 0: act_bwd_apply_map_kernel<0><<<grid, 256, 0, st>>>(m, y, mu, a, b, slope, p, q, r, dy, P, am); break;
    case 1: act_bwd_apply_map_kernel<1><<<grid, 256, 0, st>>>(m, y, mu, a, b, slope, p, q, r, dy, P, am); break;
    case 2: act_bwd_apply_map_kernel<2><<<grid, 256, 0, st>>>(m, y, mu, a, b, slope, p, q, r, dy, P, am); break;
    default: act_bwd_apply_map_kernel<3><<<grid, 256, 0, st>>>(m, y, mu, a, b, slope, p, q, r, dy, P, am); break;
  }
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_in_bwd_fused_map(const float* g, int Ctot, int c0, int mode, const float* y, const float* mu, const float* a,
                         float slope, float* dy, int N, int Cy, int Hy, int Wy, float* absmax, void* stream) {
  SAN_CHECK_ARG(y && mu && a && dy, "san_in_bwd_fused_map: bad args");
  int rc = gmap_check(g, N, Ctot, c0, Cy, Hy, Wy, mode, "san_in_bwd_fused_map");
  if (rc != SAN_OK) return rc;
  const GMap m = make_gmap(g, Ctot, c0, Cy, Hy, Wy, mode);
  const int planes = N * Cy, P = (mode == 2 ? 4 : 1) * Hy * Wy;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* am = (unsigned int*)absmax;
  if (am) SAN_CUDA(cudaMemsetAsync(am, 0, sizeof(float), st));
  // threads per plane: a quarter of the float4 groups, 128 .. 1024 (big planes: one 1024-thread CTA per SM keeps the
  // live planes of all SMs near the L2 capacity)
  int nt = ((P / 16) + 31) / 32 * 32;
  static const int nt_max = [] { const char* e = getenv("SAN_IN_BWD_NT"); const int v = e ? atoi(e) : 1024; return v < 128 ? 128 : (v > 1024 ? 1024 : v / 32 * 32); }();
  nt = nt < 128 ? 128 : (nt > nt_max ? nt_max : nt);
  // planes too large for 148 of them (+ their gradients) to stay in L2: two CTAs of a cluster per plane (SAN_IN_BWD_CL2=0: off)
  static const int cl2_env = [] { const char* e = getenv("SAN_IN_BWD_CL2"); return e ? atoi(e) : 1; }();
  if (cl2_env && (long long)P * 8 * san_num_sms() > 100ll * 1024 * 1024) {
    switch (mode) {
      case 0: in_bwd_fused_map_cl2_kernel<0><<<2 * planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
      case 1: in_bwd_fused_map_cl2_kernel<1><<<2 * planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
      case 2: in_bwd_fused_map_cl2_kernel<2><<<2 * planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
      default: in_bwd_fused_map_cl2_kernel<3><<<2 * planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
    }
    SAN_LAUNCH_CHECK();
    return SAN_OK;
  }
  switch (mode) {
    case 0: in_bwd_fused_map_kernel<0><<<planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
    case 1: in_bwd_fused_map_kernel<1><<<planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
    case 2: in_bwd_fused_map_kernel<2><<<planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
    default: in_bwd_fused_map_kernel<3><<<planes, nt, 0, st>>>(m, y, mu, a, slope, dy, P, am); break;
  }
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_pool2(const float* x, float* y, long long planes, int H, int W, float scale, void* stream) {
  SAN_CHECK_ARG(x && y && planes > 0 && H >= 2 && W >= 2, "san_pool2: bad args");
  const int Ho = H / 2, Wo = W / 2;
  const long long total = planes * Ho * Wo;
  pool2_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, y, H, W, Ho, Wo, scale, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_up2(const float* x, float* y, long long planes, int H, int W, float scale, void* stream) {
  SAN_CHECK_ARG(x && y && planes > 0 && H > 0 && W > 0, "san_up2: bad args");
  const long long total = planes * 4 * H * W;
  up2_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, y, H, W, scale, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_depth_to_space2(const float* x, float* y, int N, int Co, int H, int W, void* stream) {
  SAN_CHECK_ARG(x && y && N > 0 && Co > 0 && H > 0 && W > 0, "san_depth_to_space2: bad args");
  const long long total = (long long)N * Co * 4 * H * W;
  shuffle2_kernel<true><<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, y, Co, H, W, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_space_to_depth2(const float* y, float* x, int N, int Co, int H, int W, void* stream) {
  SAN_CHECK_ARG(x && y && N > 0 && Co > 0 && H > 0 && W > 0, "san_space_to_depth2: bad args");
  const long long total = (long long)N * Co * 4 * H * W;
  shuffle2_kernel<false><<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(y, x, Co, H, W, total);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
