// FP32 direct convolution kernels (NCHW) for the U-Net conv stacks:
// reference varnet.py:139-146 (3x3 no-bias), :75-80 (1x1 + bias), :176-179
// (ConvTranspose2d 2x2 s2, run as a 1x1 conv + pixel shuffle) and unet.py:119-140
// (3x3 / 1x1 + bias).  This is the exact-fp32 path: shared-memory tiled FFMA with
// register blocking (4 pixels x CPT output channels per thread).  The default path
// of the U-Nets is the tcgen05 BF16x3 implicit GEMM of conv_tc.cu / wgrad_tc.cu; these
// kernels remain as the SAN_TC=0 A/B path, as the on-device checker in the tests and
// for weight gradients of images narrower than 16 pixels.
//
//   forward : y[n,co,h,w] = bias[co] + sum_{ci,r,s} x[n,ci,h+r-p,w+s-p] * w[co,ci,r,s]
//   dgrad   : the same kernel on dY with weights packed flipped/transposed
//   wgrad   : dW[co,ci,r,s] = sum_{n,h,w} dY[n,co,h,w] * x[n,ci,h+r-p,w+s-p]
#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

constexpr int CI_T = 8;

// packed weights: P[ci][tap][co]  (co contiguous)
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KK,
                                    int dgrad) {
  const int total = Cout * Cin * KK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (!dgrad) {
      // out index: (ci*KK + tap)*Cout + co
      const int co = i % Cout, t = i / Cout, tap = t % KK, ci = t / KK;
      out[i] = w[(co * Cin + ci) * KK + tap];
    } else {
      // conv over dY: input channels = Cout, output channels = Cin
      // out index: (co*KK + tap)*Cin + ci
      const int ci = i % Cin, t = i / Cin, tap = t % KK, co = t / KK;
      out[i] = w[(co * Cin + ci) * KK + (KK - 1 - tap)];
    }
  }
}

// Block = 64*NG threads.  Pixel-thread pt = tid % 64 owns 4 consecutive pixels of one
// row of a (4*PXW) x PXH tile; cout-group g = tid / 64 owns CPT output channels.
template <int K, int CPT, int NG>
__global__ void __launch_bounds__(64 * NG) conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wp,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           int Cin, int H, int W, int Cout, long long x_bs,
                                                           long long y_bs, int PXW, int PXH, int tiles_w) {
  constexpr int KK = K * K, PAD = K / 2, CO_T = CPT * NG, NT = 64 * NG;
  extern __shared__ float smem[];
  const int TW = 4 * PXW, TH = PXH;
  const int IW = TW + K - 1, IH = TH + K - 1;
  const int IWP = IW | 1;  // odd row stride
  float* xs = smem;                    // [CI_T][IH][IWP]
  float* ws = smem + CI_T * IH * IWP;  // [CI_T][KK][CO_T]
  const int tid = threadIdx.x;
  const int pt = tid & 63, g = tid >> 6;
  const int tx = pt % PXW, ty = pt / PXW;
  const bool active = ty < PXH;
  const int tile = blockIdx.x;
  const int w0 = (tile % tiles_w) * TW, h0 = (tile / tiles_w) * TH;
  const int co0 = blockIdx.y * CO_T;
  const int n = blockIdx.z;
  const float* xn = x + n * x_bs;
  const long long HW = (long long)H * W;

  float acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CI_T) {
    const int nci = min(CI_T, Cin - ci0);
    for (int i = tid; i < CI_T * IH * IW; i += NT) {
      const int c = i % IW;
      const int t = i / IW;
      const int r = t % IH, ci = t / IH;
      const int gh = h0 + r - PAD, gw = w0 + c - PAD;
      float v = 0.f;
      if (ci < nci && gh >= 0 && gh < H && gw >= 0 && gw < W) v = xn[(ci0 + ci) * HW + (long long)gh * W + gw];
      xs[(ci * IH + r) * IWP + c] = v;
    }
    for (int i = tid; i < CI_T * KK * CO_T; i += NT) {
      const int co = i % CO_T;
      const int t = i / CO_T;  // ci*KK + tap
      const int ci = t / KK;
      float v = 0.f;
      if (ci < nci && co0 + co < Cout) v = wp[((long long)(ci0 * KK + t)) * Cout + co0 + co];
      ws[i] = v;
    }
    __syncthreads();
    if (active) {
      for (int ci = 0; ci < nci; ++ci) {
        float xin[K][4 + K - 1];
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
          for (int c = 0; c < 4 + K - 1; ++c) xin[r][c] = xs[(ci * IH + ty + r) * IWP + tx * 4 + c];
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
          for (int s = 0; s < K; ++s) {
            const float* wv = ws + (ci * KK + r * K + s) * CO_T + g * CPT;
            float wr[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) wr[j] = wv[j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(xin[r][s + i], wr[j], acc[i][j]);
          }
      }
    }
    __syncthreads();
  }
  if (!active) return;
  const int h = h0 + ty, wq = w0 + tx * 4;
  if (h >= H) return;
  float* yn = y + n * y_bs;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int co = co0 + g * CPT + j;
    if (co >= Cout) break;
    const float bv = bias ? bias[co] : 0.f;
    float* dst = yn + co * HW + (long long)h * W + wq;
    if (wq + 3 < W && ((W & 3) == 0)) {
      *(float4*)dst = make_float4(acc[0][j] + bv, acc[1][j] + bv, acc[2][j] + bv, acc[3][j] + bv);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (wq + i < W) dst[i] = acc[i][j] + bv;
    }
  }
}

// wgrad: block (256 threads) owns a CI_T x 32 (ci, co) tile of dW and walks pixel
// tiles (32 wide x 8 rows) assigned to its split; thread = (rq, ci, cog):
// rq in [0,4) picks rows {rq, rq+4}, ci in [0,8), cog in [0,8) owns 4 output channels.
constexpr int WG_TW = 32, WG_TH = 8, WG_CO = 32;
template <int K>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, float* __restrict__ dbias, int N,
                                                         int Cin, int H, int W, int Cout, long long x_bs,
                                                         long long dy_bs, int tiles_w, int tiles_h, int nsplit) {
  constexpr int KK = K * K, PAD = K / 2;
  constexpr int IW = WG_TW + K - 1, IH = WG_TH + K - 1, IWP = IW | 1;
  constexpr int DYS = WG_TW + 1;
  __shared__ float xs[CI_T * IH * IWP];
  __shared__ float ds[WG_CO * WG_TH * DYS];
  const int tid = threadIdx.x;
  const int rq = tid & 3, ci = (tid >> 2) & 7, cog = tid >> 5;
  const int n_ci_t = (Cin + CI_T - 1) / CI_T;
  const int ci0 = (blockIdx.x % n_ci_t) * CI_T, co0 = (blockIdx.x / n_ci_t) * WG_CO;
  const int split = blockIdx.y;
  const long long HW = (long long)H * W;
  const int ntile = N * tiles_h * tiles_w;

  float acc[4][KK];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < KK; ++t) acc[j][t] = 0.f;

  for (int tile = split; tile < ntile; tile += nsplit) {
    const int n = tile / (tiles_h * tiles_w);
    const int tt = tile - n * tiles_h * tiles_w;
    const int h0 = (tt / tiles_w) * WG_TH, w0 = (tt % tiles_w) * WG_TW;
    const float* xn = x + n * x_bs;
    const float* dn = dy + n * dy_bs;
    __syncthreads();
    for (int i = tid; i < CI_T * IH * IW; i += 256) {
      const int c = i % IW;
      const int t = i / IW;
      const int r = t % IH, cc = t / IH;
      const int gh = h0 + r - PAD, gw = w0 + c - PAD;
      float v = 0.f;
      if (ci0 + cc < Cin && gh >= 0 && gh < H && gw >= 0 && gw < W) v = xn[(ci0 + cc) * HW + (long long)gh * W + gw];
      xs[(cc * IH + r) * IWP + c] = v;
    }
    for (int i = tid; i < WG_CO * WG_TH * WG_TW; i += 256) {
      const int c = i % WG_TW;
      const int t = i / WG_TW;
      const int r = t % WG_TH, co = t / WG_TH;
      const int gh = h0 + r, gw = w0 + c;
      float v = 0.f;
      if (co0 + co < Cout && gh < H && gw < W) v = dn[(co0 + co) * HW + (long long)gh * W + gw];
      ds[(co * WG_TH + r) * DYS + c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = rq + rr * 4;
      float win[K][K];
#pragma unroll
      for (int r = 0; r < K; ++r)
#pragma unroll
        for (int s = 1; s < K; ++s) win[r][s] = xs[(ci * IH + row + r) * IWP + s - 1];
      for (int c = 0; c < WG_TW; ++c) {
#pragma unroll
        for (int r = 0; r < K; ++r) {
#pragma unroll
          for (int s = 0; s < K - 1; ++s) win[r][s] = win[r][s + 1];
          win[r][K - 1] = xs[(ci * IH + row + r) * IWP + c + K - 1];
        }
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = ds[((cog * 4 + j) * WG_TH + row) * DYS + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bacc[j] += d[j];
#pragma unroll
          for (int r = 0; r < K; ++r)
#pragma unroll
            for (int s = 0; s < K; ++s) acc[j][r * K + s] = fmaf(d[j], win[r][s], acc[j][r * K + s]);
        }
      }
    }
  }
  // reduce the 4 row-quarters (adjacent lanes), then one atomic per output
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int t = 0; t < KK; ++t) {
      float v = acc[j][t];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      acc[j][t] = v;
    }
    float b = bacc[j];
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    bacc[j] = b;
  }
  if (rq == 0 && ci0 + ci < Cin) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + cog * 4 + j;
      if (co >= Cout) break;
#pragma unroll
      for (int t = 0; t < KK; ++t) atomicAdd(dw + ((long long)co * Cin + ci0 + ci) * KK + t, acc[j][t]);
      if (dbias && ci0 + ci == 0) atomicAdd(dbias + co, bacc[j]);
    }
  }
}

template <int K, int CPT, int NG>
int launch_fwd(const float* x, const float* wp, const float* bias, float* y, int N, int Cin, int H, int W, int Cout,
               long long x_bs, long long y_bs, cudaStream_t st) {
  int PXW, PXH;
  if (W % 32 == 0 || W > 80) { PXW = 8; PXH = 8; }
  else if (W > 20) { PXW = 10; PXH = 6; }
  else if (W > 16) { PXW = 5; PXH = 12; }
  else if (W > 8) { PXW = 4; PXH = 16; }
  else { PXW = 2; PXH = 32; }
  const int TW = 4 * PXW, TH = PXH;
  const int tiles_w = san_cdiv(W, TW), tiles_h = san_cdiv(H, TH);
  const int IW = TW + K - 1, IH = TH + K - 1, IWP = IW | 1;
  const size_t smem = sizeof(float) * ((size_t)CI_T * IH * IWP + (size_t)CI_T * K * K * CPT * NG);
  SAN_CHECK_ARG(smem <= 48 * 1024, "conv: shared memory %zu too large", smem);
  SAN_CHECK_ARG(N <= 65535, "conv: batch %d too large for grid.z", N);
  dim3 grid(tiles_w * tiles_h, san_cdiv(Cout, CPT * NG), N);
  conv_fwd_kernel<K, CPT, NG><<<grid, 64 * NG, smem, st>>>(x, wp, bias, y, Cin, H, W, Cout, x_bs, y_bs, PXW, PXH, tiles_w);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

template <int K>
int dispatch_fwd(const float* x, const float* wp, const float* bias, float* y, int N, int Cin, int H, int W, int Cout,
                 long long x_bs, long long y_bs, cudaStream_t st) {
  // pick channels-per-thread and group count minimising padded output channels
  int cpt;
  if (Cout <= 2) cpt = 2;
  else if (Cout <= 4) cpt = 4;
  else if (Cout % 8 == 0) cpt = 8;
  else if (Cout % 6 == 0) cpt = 6;
  else cpt = 8;
  const int need = san_cdiv(Cout, cpt);
  static const int ngs[5] = {6, 4, 3, 2, 1};
  int best = 1, best_cost = 1 << 30;
  for (int i = 0; i < 5; ++i) {
    const int ng = ngs[i];
    if (cpt < 6 && ng > 1) continue;
    const int cost = san_cdiv(need, ng) * ng;
    if (cost < best_cost) { best_cost = cost; best = ng; }
  }
#define GO(C, G) return launch_fwd<K, C, G>(x, wp, bias, y, N, Cin, H, W, Cout, x_bs, y_bs, st)
  if (cpt == 2) GO(2, 1);
  if (cpt == 4) GO(4, 1);
  if (cpt == 6) {
    switch (best) { case 6: GO(6, 6); case 4: GO(6, 4); case 3: GO(6, 3); case 2: GO(6, 2); default: GO(6, 1); }
  }
  switch (best) { case 6: GO(8, 6); case 4: GO(8, 4); case 3: GO(8, 3); case 2: GO(8, 2); default: GO(8, 1); }
#undef GO
}

}  // namespace

extern "C" {

int san_conv_pack_weights(const float* w, float* packed, int Cout, int Cin, int K, int dgrad, void* stream) {
  SAN_CHECK_ARG(w && packed && Cout > 0 && Cin > 0 && K > 0, "san_conv_pack_weights: bad args");
  const int total = Cout * Cin * K * K;
  pack_weights_kernel<<<san_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, packed, Cout, Cin, K * K, dgrad);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_conv2d_fwd(const float* x, const float* w_packed, const float* bias, float* y, int N, int Cin, int H, int W,
                   int Cout, int K, long long x_bs, long long y_bs, void* stream) {
  SAN_CHECK_ARG(x && w_packed && y && N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "san_conv2d_fwd: bad args");
  if (x_bs <= 0) x_bs = (long long)Cin * H * W;
  if (y_bs <= 0) y_bs = (long long)Cout * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 3) return dispatch_fwd<3>(x, w_packed, bias, y, N, Cin, H, W, Cout, x_bs, y_bs, st);
  if (K == 1) return dispatch_fwd<1>(x, w_packed, bias, y, N, Cin, H, W, Cout, x_bs, y_bs, st);
  san_set_error("san_conv2d_fwd: kernel size %d unsupported (1 or 3)", K);
  return SAN_ERR_UNSUPPORTED;
}

int san_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias, int N, int Cin, int H, int W, int Cout,
                     int K, long long x_bs, long long dy_bs, void* stream) {
  SAN_CHECK_ARG(x && dy && dw && N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "san_conv2d_wgrad: bad args");
  SAN_CHECK_ARG(K == 1 || K == 3, "san_conv2d_wgrad: kernel size %d unsupported (1 or 3)", K);
  if (x_bs <= 0) x_bs = (long long)Cin * H * W;
  if (dy_bs <= 0) dy_bs = (long long)Cout * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  SAN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * K * K, st));
  if (dbias) SAN_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)Cout, st));
  const int tiles_w = san_cdiv(W, WG_TW), tiles_h = san_cdiv(H, WG_TH);
  const int ntile = N * tiles_w * tiles_h;
  const int nblk = san_cdiv(Cin, CI_T) * san_cdiv(Cout, WG_CO);
  int nsplit = san_cdiv((long long)san_num_sms() * 4, nblk);
  if (nsplit > ntile) nsplit = ntile;
  if (nsplit > 65535) nsplit = 65535;
  if (nsplit < 1) nsplit = 1;
  dim3 grid(nblk, nsplit);
  if (K == 3)
    conv_wgrad_kernel<3><<<grid, 256, 0, st>>>(x, dy, dw, dbias, N, Cin, H, W, Cout, x_bs, dy_bs, tiles_w, tiles_h, nsplit);
  else
    conv_wgrad_kernel<1><<<grid, 256, 0, st>>>(x, dy, dw, dbias, N, Cin, H, W, Cout, x_bs, dy_bs, tiles_w, tiles_h, nsplit);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
