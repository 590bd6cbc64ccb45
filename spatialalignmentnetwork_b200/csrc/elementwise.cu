// Element-wise / small-reduction kernels around the FFT data-consistency path:
// soft-DC backward, sensitivity-map gradients, root-sum-of-squares, sensitivity
// normalisation (reference varnet.py:419, 508-530; signal_utils.py:24-26).
#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

__global__ void finalize_scalar_kernel(const double* acc, float* out, double scale) {
  out[0] = (float)(acc[0] * scale);
}

// out[b,p] = sign * u[b,p] * conj(pl[n,p]),  n = b / C, pl planar [N,2,P]
__global__ void cmul_conj_planar_kernel(const float2* __restrict__ u, const float* __restrict__ pl,
                                        float2* __restrict__ out, int C, long long P, long long total,
                                        float sign) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / P, p = i - b * P;
    const long long n = b / C;
    const float2 x = make_float2(pl[(n * 2) * P + p], pl[(n * 2 + 1) * P + p]);
    float2 r = cmulc(u[i], x);
    out[i] = make_float2(sign * r.x, sign * r.y);
  }
}

// dk = G - where(m, G, 0) * w ;  dk0 = where(m, G, 0) * w ;  acc += -sum Re(conj(where(m, k-k0, 0)) * G)
__global__ void dc_bwd_kernel(const float2* __restrict__ G, const float2* __restrict__ k,
                              const float2* __restrict__ k0, const unsigned char* __restrict__ mask,
                              const float* __restrict__ dcw, float2* __restrict__ dk, float2* __restrict__ dk0,
                              double* acc, int W, long long total) {
  __shared__ double red[32];
  const float w = __ldg(dcw);
  double local = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % W);
    float2 g = G[i];
    if (mask[col]) {
      const float2 a = k[i], b = k0[i];
      const float dx = a.x - b.x, dy = a.y - b.y;
      local -= (double)(dx * g.x + dy * g.y);
      if (dk) dk[i] = make_float2(g.x - g.x * w, g.y - g.y * w);
      if (dk0) dk0[i] = make_float2(g.x * w, g.y * w);
    } else {
      if (dk) dk[i] = g;
      if (dk0) dk0[i] = make_float2(0.f, 0.f);
    }
  }
  local = block_sum_d(local, red);
  if (threadIdx.x == 0) atomicAdd(acc, local);
}

// x viewed as [N, C, P, T] (T = 2 for complex64, 1 for real) -> out [N, P]
__global__ void rss_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int C, long long P, int T,
                               long long NP) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < NP;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* q = x + ((n * C + c) * P + p) * T;
      for (int t = 0; t < T; ++t) s += q[t] * q[t];
    }
    out[i] = sqrtf(s);
  }
}

// dx[n,c,p,t] = g[n,p] * x / r   (0 where r == 0, the library sub-gradient)
__global__ void rss_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                               const float* __restrict__ r, float* __restrict__ dx, int C, long long P, int T,
                               long long NP) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < NP;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    const float rr = r[i];
    const float f = rr > 0.f ? g[i] / rr : 0.f;
    for (int c = 0; c < C; ++c) {
      const long long o = ((n * C + c) * P + p) * T;
      for (int t = 0; t < T; ++t) dx[o + t] = f * x[o + t];
    }
  }
}

// s planar [N*C, 2, P] -> S c64 [N, C, P] = s / (rss_c(s) + eps)
__global__ void sens_normalize_fwd_kernel(const float* __restrict__ s, float2* __restrict__ S, int C,
                                          long long P, long long NP, float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < NP;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    float t = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long b = n * C + c;
      const float re = s[(b * 2) * P + p], im = s[(b * 2 + 1) * P + p];
      t += re * re + im * im;
    }
    const float inv = 1.f / (sqrtf(t) + eps);
    for (int c = 0; c < C; ++c) {
      const long long b = n * C + c;
      S[b * P + p] = make_float2(s[(b * 2) * P + p] * inv, s[(b * 2 + 1) * P + p] * inv);
    }
  }
}

// ds_c = G_c/(r+eps) - s_c/(r (r+eps)^2) * sum_c' Re(conj(G_c') s_c')
__global__ void sens_normalize_bwd_kernel(const float2* __restrict__ G, const float* __restrict__ s,
                                          float* __restrict__ ds, int C, long long P, long long NP, float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < NP;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    float t = 0.f, dot = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long b = n * C + c;
      const float re = s[(b * 2) * P + p], im = s[(b * 2 + 1) * P + p];
      const float2 g = G[b * P + p];
      t += re * re + im * im;
      dot += g.x * re + g.y * im;
    }
    const float r = sqrtf(t);
    const float inv = 1.f / (r + eps);
    const float coef = r > 0.f ? dot * inv * inv / r : 0.f;
    for (int c = 0; c < C; ++c) {
      const long long b = n * C + c;
      const float re = s[(b * 2) * P + p], im = s[(b * 2 + 1) * P + p];
      const float2 g = G[b * P + p];
      ds[(b * 2) * P + p] = g.x * inv - coef * re;
      ds[(b * 2 + 1) * P + p] = g.y * inv - coef * im;
    }
  }
}

__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                             float a, float b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}

inline int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)san_num_sms() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

int san_finalize_scalar(const double* acc, float* out, double scale, cudaStream_t st) {
  finalize_scalar_kernel<<<1, 1, 0, st>>>(acc, out, scale);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

// max |x| of a tensor into a device scalar (the dynamic fp16-pair scale of gradient operands, tc_common.cuh):
// non-negative floats order like their bit patterns, so one atomicMax on the uint view per block suffices.
// NaN / inf elements are skipped (they would poison the scale; the values themselves still propagate).
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out) {
  float m = 0.f;
  const long long n4 = (((uintptr_t)x & 15) == 0) ? n >> 2 : 0;     // vector path only for 16 B aligned tensors
  const float4* x4 = (const float4*)x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    const float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));   // fmaxf drops NaN operands
    m = fmaxf(m, a);
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  if (!(m < 3.0e38f)) m = 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    atomicMax(out, __float_as_uint(m));
  }
}

extern "C" {

int san_cmul_conj_planar(const void* u, const float* planar, void* out, int N, int C, long long P, float sign,
                         void* stream) {
  SAN_CHECK_ARG(u && planar && out && N > 0 && C > 0 && P > 0, "san_cmul_conj_planar: bad args");
  const long long total = (long long)N * C * P;
  cmul_conj_planar_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>((const float2*)u, planar, (float2*)out, C, P,
                                                                            total, sign);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_dc_bwd(const void* G, const void* k, const void* k0, const unsigned char* mask, const float* dc_weight,
               void* dk, void* dk0, float* d_dc_weight, double* scratch, long long rows, int W, void* stream) {
  SAN_CHECK_ARG(G && k && k0 && mask && dc_weight && d_dc_weight && scratch && rows > 0 && W > 0, "san_dc_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  SAN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  const long long total = rows * W;
  dc_bwd_kernel<<<ew_grid(total), 256, 0, st>>>((const float2*)G, (const float2*)k, (const float2*)k0, mask, dc_weight,
                                                (float2*)dk, (float2*)dk0, scratch, W, total);
  SAN_LAUNCH_CHECK();
  return san_finalize_scalar(scratch, d_dc_weight, 1.0, st);
}

int san_rss_fwd(const float* x, float* out, int N, int C, long long P, int is_complex, void* stream) {
  SAN_CHECK_ARG(x && out && N > 0 && C > 0 && P > 0, "san_rss_fwd: bad args");
  const long long NP = (long long)N * P;
  rss_fwd_kernel<<<ew_grid(NP), 256, 0, (cudaStream_t)stream>>>(x, out, C, P, is_complex ? 2 : 1, NP);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_rss_bwd(const float* g, const float* x, const float* r, float* dx, int N, int C, long long P,
                int is_complex, void* stream) {
  SAN_CHECK_ARG(g && x && r && dx && N > 0 && C > 0 && P > 0, "san_rss_bwd: bad args");
  const long long NP = (long long)N * P;
  rss_bwd_kernel<<<ew_grid(NP), 256, 0, (cudaStream_t)stream>>>(g, x, r, dx, C, P, is_complex ? 2 : 1, NP);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_sens_normalize_fwd(const float* s_planar, void* S, int N, int C, long long P, float eps, void* stream) {
  SAN_CHECK_ARG(s_planar && S && N > 0 && C > 0 && P > 0, "san_sens_normalize_fwd: bad args");
  const long long NP = (long long)N * P;
  sens_normalize_fwd_kernel<<<ew_grid(NP), 256, 0, (cudaStream_t)stream>>>(s_planar, (float2*)S, C, P, NP, eps);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_sens_normalize_bwd(const void* G, const float* s_planar, float* ds_planar, int N, int C, long long P,
                           float eps, void* stream) {
  SAN_CHECK_ARG(G && s_planar && ds_planar && N > 0 && C > 0 && P > 0, "san_sens_normalize_bwd: bad args");
  const long long NP = (long long)N * P;
  sens_normalize_bwd_kernel<<<ew_grid(NP), 256, 0, (cudaStream_t)stream>>>((const float2*)G, s_planar, ds_planar, C, P,
                                                                           NP, eps);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_absmax(const float* x, long long n, float* out, void* stream) {
  SAN_CHECK_ARG(x && out && n > 0, "san_absmax: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  SAN_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  absmax_kernel<<<ew_grid(n >> 2 > 0 ? n >> 2 : 1), 256, 0, st>>>(x, n, (unsigned int*)out);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_axpby(const float* x, const float* y, float* out, float a, float b, long long n, void* stream) {
  SAN_CHECK_ARG(x && out && n > 0, "san_axpby: bad args");
  axpby_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, out, a, b, n);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
