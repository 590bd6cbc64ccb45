// 2-D batched orthonormal FFT kernels with fused prologues/epilogues.
//
// Replaces the torch.fft.fft2/ifft2(norm='ortho') call sites of the reference's
// VarNet data-consistency path (reference signal_utils.py:4-12, varnet.py:508-512,
// 525-530, 395-402, 486).  Separable: a row pass (FFT along W, contiguous) followed
// by a column pass (FFT along H on a tile of CT adjacent columns so global accesses
// stay coalesced).  Each pass stages its lines in shared memory and runs an
// autosort (Stockham) mixed-radix FFT there; radices 2/3/4/5 are specialised, any
// other prime factor (e.g. 23 for W=368) uses a direct DFT butterfly.  Element-wise
// work around the transform (coil-sensitivity multiply, coil reduction, ACS column
// mask, masked soft data-consistency + residual, root-sum-of-squares) is fused into
// the row-pass loader / column-pass storer so k-space makes one HBM round trip.
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include <cuda.h>      // CUtensorMap + enums only: cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint

#include "san_common.cuh"
#include "tc_common.cuh"
#include "fft_v2.cuh"
#include "../../include/san_b200.h"

namespace {

using namespace san_fft;

bool make_plan(int n, FftPlan* p) {
  p->n = n;
  p->ns = 0;
  int m = n;
  while (m % 4 == 0) { p->radix[p->ns++] = 4; m /= 4; }
  while (m % 2 == 0) { p->radix[p->ns++] = 2; m /= 2; }
  for (int f = 3; m > 1; f += 2) {
    while (m % f == 0) {
      if (f > 32 || p->ns >= 12) return false;
      p->radix[p->ns++] = f;
      m /= f;
    }
  }
  if (n == 1) { p->ns = 0; }
  int Ns = 1;
  for (int s = 0; s < p->ns; ++s) {
    int sh = 0;
    while ((1 << sh) < Ns) ++sh;
    p->shift[s] = ((1 << sh) == Ns) ? sh : -1;
    p->twstep[s] = n / (Ns * p->radix[s]);
    Ns *= p->radix[s];
  }
  return true;
}

// forward twiddles exp(-2*pi*i*j/n), generated in fp64 (one table per length and device)
std::mutex g_tw_mutex;
std::map<std::pair<int, int>, float2*> g_tw;

int get_twiddles(int n, const float2** out) {
  int dev = 0;
  SAN_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tw_mutex);
  auto it = g_tw.find({dev, n});
  if (it != g_tw.end()) { *out = it->second; return SAN_OK; }
  std::vector<float2> h(n);
  for (int j = 0; j < n; ++j) {
    double a = -2.0 * M_PI * (double)j / (double)n;
    h[j] = make_float2((float)cos(a), (float)sin(a));
  }
  float2* d = nullptr;
  SAN_CUDA(cudaMalloc(&d, sizeof(float2) * n));
  SAN_CUDA(cudaMemcpy(d, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
  g_tw[{dev, n}] = d;
  *out = d;
  return SAN_OK;
}

// twiddles are read from the block's shared-memory copy of the table
template <bool INV>
__device__ __forceinline__ float2 twid(const float2* __restrict__ tw, int idx) {
  float2 t = tw[idx];
  if (INV) t.y = -t.y;
  return t;
}

// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// One radix-R butterfly of a Stockham stage.  `in`/`out` are one line of length n;
// Ns = product of the radices of the previous stages; j in [0, n/R).
template <bool INV, int RT>
__device__ __forceinline__ void butterfly(const float2* __restrict__ in, float2* __restrict__ out,
                                          int n, int Rrt, int T, int Ns, int shift, int twstep, int j,
                                          const float2* __restrict__ tw) {
  const int R = RT ? RT : Rrt;   // RT = 0: generic odd radix known only at run time
  const int k = shift >= 0 ? (j & (Ns - 1)) : (j % Ns);
  const int jq = shift >= 0 ? (j >> shift) : (j / Ns);
  const int base = k * twstep;
  const int j0 = jq * Ns * R + k;
  if (R == 4) {
    float2 v0 = in[j], v1 = in[j + T], v2 = in[j + 2 * T], v3 = in[j + 3 * T];
    if (k) {
      v1 = cmul(v1, twid<INV>(tw, base));
      v2 = cmul(v2, twid<INV>(tw, 2 * base));
      v3 = cmul(v3, twid<INV>(tw, 3 * base));
    }
    float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
    float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    float2 a3 = mul_mi<INV>(make_float2(v1.x - v3.x, v1.y - v3.y));
    out[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
    out[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
    out[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
    out[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
  } else if (R == 2) {
    float2 v0 = in[j], v1 = in[j + T];
    if (k) v1 = cmul(v1, twid<INV>(tw, base));
    out[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
    out[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
  } else if (R == 5) {
    float2 v[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      v[r] = in[j + r * T];
      if (r && k) v[r] = cmul(v[r], twid<INV>(tw, r * base));
    }
    const int q = n / 5;
    const float2 w1 = twid<INV>(tw, q), w2 = twid<INV>(tw, 2 * q);
    // symmetric / antisymmetric pairs
    float2 s1 = make_float2(v[1].x + v[4].x, v[1].y + v[4].y);
    float2 d1 = make_float2(v[1].x - v[4].x, v[1].y - v[4].y);
    float2 s2 = make_float2(v[2].x + v[3].x, v[2].y + v[3].y);
    float2 d2 = make_float2(v[2].x - v[3].x, v[2].y - v[3].y);
    out[j0] = make_float2(v[0].x + s1.x + s2.x, v[0].y + s1.y + s2.y);
    // X1 = v0 + c1*s1 + c2*s2 + i*(sn1*d1 + sn2*d2)   with w1 = c1 + i*sn1 (sign carries INV)
    float2 p1 = make_float2(v[0].x + w1.x * s1.x + w2.x * s2.x, v[0].y + w1.x * s1.y + w2.x * s2.y);
    float2 q1 = make_float2(w1.y * d1.x + w2.y * d2.x, w1.y * d1.y + w2.y * d2.y);
    float2 p2 = make_float2(v[0].x + w2.x * s1.x + w1.x * s2.x, v[0].y + w2.x * s1.y + w1.x * s2.y);
    float2 q2 = make_float2(w2.y * d1.x - w1.y * d2.x, w2.y * d1.y - w1.y * d2.y);
    // i*q = (-q.y, q.x)
    out[j0 + Ns] = make_float2(p1.x - q1.y, p1.y + q1.x);
    out[j0 + 4 * Ns] = make_float2(p1.x + q1.y, p1.y - q1.x);
    out[j0 + 2 * Ns] = make_float2(p2.x - q2.y, p2.y + q2.x);
    out[j0 + 3 * Ns] = make_float2(p2.x + q2.y, p2.y - q2.x);
  } else {
    // generic direct DFT butterfly (R <= 32)
    float2 v[32];
    for (int r = 0; r < R; ++r) {
      v[r] = in[j + r * T];
      if (r && k) v[r] = cmul(v[r], twid<INV>(tw, r * base));
    }
    const int q = n / R;
    for (int o = 0; o < R; ++o) {
      float2 acc = v[0];
      int e = 0;
      for (int r = 1; r < R; ++r) {
        e += o;
        if (e >= R) e -= R;
        float2 w = twid<INV>(tw, e * q);
        acc.x += v[r].x * w.x - v[r].y * w.y;
        acc.y += v[r].x * w.y + v[r].y * w.x;
      }
      out[j0 + o * Ns] = acc;
    }
  }
}

template <bool INV, int RT>
__device__ __forceinline__ void fft_stage(const float2* a, float2* b, int nfft, int stride, int n, int R, int Ns, int shift,
                                          int twstep, const float2* __restrict__ tw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int T = n / R;
  for (int f = warp; f < nfft; f += nw) {
    const float2* in = a + f * stride;
    float2* out = b + f * stride;
    for (int j = lane; j < T; j += 32) butterfly<INV, RT>(in, out, n, R, T, Ns, shift, twstep, j, tw);
  }
}

// Runs `nfft` independent FFTs (line f at a + f*stride); returns the buffer holding the result.
// Each WARP owns whole lines (f = warp, warp + nwarps, ...), so the Stockham stages of a line only need
// __syncwarp() between them, and the line / butterfly indices need no divisions; the radix of a stage
// is a template parameter.  `tw` = shared-memory twiddle table.  Caller must __syncthreads() after
// filling `a`; the result is visible to the whole block on return.
template <bool INV>
__device__ float2* block_fft(float2* a, float2* b, int nfft, int stride, const FftPlan& pl,
                             const float2* __restrict__ tw) {
  int Ns = 1;
  for (int s = 0; s < pl.ns; ++s) {
    const int R = pl.radix[s];
    const int shift = pl.shift[s], twstep = pl.twstep[s];
    if (R == 4) fft_stage<INV, 4>(a, b, nfft, stride, pl.n, R, Ns, shift, twstep, tw);
    else if (R == 2) fft_stage<INV, 2>(a, b, nfft, stride, pl.n, R, Ns, shift, twstep, tw);
    else if (R == 5) fft_stage<INV, 5>(a, b, nfft, stride, pl.n, R, Ns, shift, twstep, tw);
    else fft_stage<INV, 0>(a, b, nfft, stride, pl.n, R, Ns, shift, twstep, tw);
    __syncwarp();
    float2* t = a; a = b; b = t;
    Ns *= R;
  }
  __syncthreads();
  return a;
}

template <bool INV, int LOAD>
__global__ void __launch_bounds__(256) fft_rows_kernel(const FftArgs a) {
  extern __shared__ float2 sm[];
  const int W = a.W;
  const long long nrows = (long long)a.B * a.H;
  const long long row0 = (long long)blockIdx.x * a.rpb;
  const int nr = (int)min((long long)a.rpb, nrows - row0);
  float2* bufA = sm;
  float2* bufB = sm + a.rpb * W;
  float2* tws = sm + 2 * a.rpb * W;
  load_twiddles(tws, a.twW, W);
  const long long HW = (long long)a.H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // one warp per row: no per-element divisions
  for (int r = warp; r < nr; r += nw) {
    const long long row = row0 + r;           // = b*H + h
    const long long rowoff = row * W;         // offset in [B,H,W]
    const long long b = row / a.H;
    const long long g = (LOAD == LD_PLANAR_S) ? b / a.C : b;
    const long long hw0 = rowoff - b * HW;
    float2* dst = bufA + r * W;
    for (int w = lane; w < W; w += 32) {
      float2 v;
      if (LOAD == LD_C64) {
        v = a.in_c[rowoff + w];
      } else if (LOAD == LD_C64_COLMASK) {
        v = a.in_c[rowoff + w];
        const float m = a.colmask[w];
        v.x *= m; v.y *= m;
      } else {
        v.x = a.in_p[(g * 2) * HW + hw0 + w];
        v.y = a.in_p[(g * 2 + 1) * HW + hw0 + w];
        if (LOAD == LD_PLANAR_S) v = cmul(v, a.sens[rowoff + w]);
      }
      dst[w] = v;
    }
  }
  __syncthreads();
  float2* res = block_fft<INV>(bufA, bufB, nr, W, a.planW, tws);
  for (int i = threadIdx.x; i < nr * W; i += blockDim.x) a.tmp[row0 * W + i] = res[i];
}

// grid: (ceil(W/ct), G) where G = N for the coil-reducing stores and B otherwise.
template <bool INV, int STORE, int CT>
__global__ void __launch_bounds__(256) fft_cols_kernel(const FftArgs a) {
  extern __shared__ float2 sm[];
  const int H = a.H, W = a.W;
  const int ld = H + 1;
  float2* bufA = sm;
  float2* bufB = sm + CT * ld;
  float2* tws = sm + 2 * CT * ld;
  float2* acc = tws + H;           // only for coil-reducing stores with C > 1
  load_twiddles(tws, a.twH, H);
  const int w0 = blockIdx.x * CT;
  const int ncol = min(CT, W - w0);
  const long long HW = (long long)H * W;
  const bool reducing = (STORE == ST_REDUCE || STORE == ST_RSS);
  const int ncoil = reducing ? a.C : 1;
  const long long g = blockIdx.y;
  const int nel = H * CT;
  for (int c = 0; c < ncoil; ++c) {
    const long long b = reducing ? g * a.C + c : g;
    const float2* src = a.tmp + b * HW;
    if (c) __syncthreads();
    for (int i = threadIdx.x; i < nel; i += blockDim.x) {
      const int h = i / CT, col = i - h * CT;
      float2 v = make_float2(0.f, 0.f);
      if (col < ncol) v = src[(long long)h * W + w0 + col];
      bufA[col * ld + h] = v;
    }
    __syncthreads();
    const float2* res = block_fft<INV>(bufA, bufB, ncol, ld, a.planH, tws);
    for (int i = threadIdx.x; i < nel; i += blockDim.x) {
      const int h = i / CT, col = i - h * CT;
      if (col >= ncol) continue;
      const int w = w0 + col;
      const long long hw = (long long)h * W + w;
      const long long off = b * HW + hw;
      float2 v = res[col * ld + h];
      v.x *= a.scale; v.y *= a.scale;
      if (STORE == ST_C64) {
        a.out_c[off] = v;
      } else if (STORE == ST_C64_COLMASK) {
        const float m = a.colmask[w];
        a.out_c[off] = make_float2(v.x * m, v.y * m);
      } else if (STORE == ST_PLANAR) {
        a.out_p[(b * 2) * HW + hw] = v.x;
        a.out_p[(b * 2 + 1) * HW + hw] = v.y;
      } else if (STORE == ST_DC) {
        const float2 kk = a.k[off];
        float2 o = kk;
        if (a.dcmask[w]) {
          const float2 k0 = a.k0[off];
          const float wgt = __ldg(a.dcw);
          o.x = kk.x - (kk.x - k0.x) * wgt;
          o.y = kk.y - (kk.y - k0.y) * wgt;
        }
        a.out_c[off] = make_float2(o.x - v.x, o.y - v.y);
      } else if (STORE == ST_REDUCE) {
        if (a.out_u) a.out_u[off] = v;
        float2 t = cmulc(v, a.sens[off]);
        if (ncoil > 1) {
          if (c) { float2 p = acc[col * ld + h]; t.x += p.x; t.y += p.y; }
          if (c + 1 < ncoil) acc[col * ld + h] = t;
        }
        if (c + 1 == ncoil) {
          a.out_p[(g * 2) * HW + hw] = t.x;
          a.out_p[(g * 2 + 1) * HW + hw] = t.y;
        }
      } else if (STORE == ST_RSS) {
        if (a.out_u) a.out_u[off] = v;
        float t = v.x * v.x + v.y * v.y;
        if (ncoil > 1) {
          if (c) t += acc[col * ld + h].x;
          if (c + 1 < ncoil) acc[col * ld + h].x = t;
        }
        if (c + 1 == ncoil) a.out_p[g * HW + hw] = sqrtf(t);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// Column pass with TMA-staged epilogue operands (H = 320, 8 columns per CTA, one coil per launch unit).
// The soft-DC store reads k and k0 next to the transform, the coil-reducing store reads the sensitivity map: 20 to 40
// dependent HBM loads per thread AFTER the FFT in fft_cols_v2_kernel (batched there, at 124 registers and 3 CTAs per SM).
// Here ONE thread requests the CTA's [320 x 8] tiles of those operands as 2-D TMA tensor copies (cp.async.bulk.tensor,
// SASS UTMALDG) into shared memory before the column loads are even issued: they arrive while the transform runs, cost
// no registers, and the epilogue reads them from shared memory (a warp reads 4 rows x 64 B = 256 contiguous bytes).
struct alignas(64) FftTmaArgs {
  CUtensorMap m0;      // k (soft DC) or sens (coil reduce): [B][H][W] of 8-byte elements
  CUtensorMap m1;      // k0 (soft DC)
  FftArgs a;
};
constexpr int TMA_LINES = 8;
constexpr int TMA_BOX_H = 160;                                   // box height (<= 256): two boxes per tile
constexpr int TMA_TILE_BYTES = V2_N * TMA_LINES * 8;             // 20480
constexpr int tma_smem_bytes(int tiles) { return tiles * TMA_TILE_BYTES + TMA_LINES * V2Ex::SIZE * 8 + V2_N * 8 + 16; }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

template <bool INV, int STORE>
__global__ void __launch_bounds__(TMA_LINES * V2_N2) fft_cols_tma_kernel(const __grid_constant__ FftTmaArgs ta) {
  static_assert(STORE == ST_DC || STORE == ST_REDUCE, "TMA-staged epilogue: soft DC and coil reduce");
  extern __shared__ __align__(128) uint8_t smraw[];
  const FftArgs& a = ta.a;
  constexpr int LINES = TMA_LINES, H = V2_N;
  constexpr int TILES = (STORE == ST_DC) ? 2 : 1;                 // k + k0 | sens
  float2* e0s = (float2*)smraw;                                   // [H][LINES]
  float2* e1s = (float2*)(smraw + (TILES - 1) * TMA_TILE_BYTES);  // (aliases e0s when there is no second operand)
  float2* z = (float2*)(smraw + TILES * TMA_TILE_BYTES);
  float2* tws = z + LINES * V2Ex::SIZE;
  const uint32_t bar = smem_u32(tws + V2_N);
  const int W = a.W;
  const int w0 = blockIdx.x * LINES;
  const int ncol = (W - w0) < LINES ? (W - w0) : LINES;
  const long long HW = (long long)H * W;
  const long long b = blockIdx.y;
  const float2* src = a.tmp + b * HW;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    bool second = false;
    if (STORE == ST_DC) {                            // k0 is only read in sampled columns
      if (ncol == LINES && ((uintptr_t)(a.dcmask + w0) & 7) == 0) second = *(const unsigned long long*)(a.dcmask + w0) != 0ull;
      else for (int c = 0; c < ncol; ++c) second |= a.dcmask[w0 + c] != 0;
    }
    mbar_expect_tx(bar, (second ? 2u : 1u) * TMA_TILE_BYTES);
    tma_load_3d(smem_u32(e0s), &ta.m0, w0, 0, (int)b, bar);
    tma_load_3d(smem_u32(e0s) + TMA_BOX_H * LINES * 8, &ta.m0, w0, TMA_BOX_H, (int)b, bar);
    if (second) {
      tma_load_3d(smem_u32(e1s), &ta.m1, w0, 0, (int)b, bar);
      tma_load_3d(smem_u32(e1s) + TMA_BOX_H * LINES * 8, &ta.m1, w0, TMA_BOX_H, (int)b, bar);
    }
  }
  {
    const int l = threadIdx.x / LINES, col = threadIdx.x - l * LINES;
    float2 v[V2_N1];
    if (col < ncol) {
#pragma unroll
      for (int j = 0; j < V2_N1; ++j) v[j] = src[(long long)(V2_N2 * j + l) * W + w0 + col];
    }
    load_twiddles(tws, a.twH, V2_N);
    __syncthreads();                                 // twiddles + the initialised barrier visible
    if (col < ncol) {
      fft_small::phase1<INV, V2_N1, V2_N2>(v, l, tws);
#pragma unroll
      for (int k1 = 0; k1 < V2_N1; ++k1) z[col * V2Ex::SIZE + V2Ex::at(k1, l)] = v[k1];
    }
  }
  __syncthreads();
  if (threadIdx.x < LINES * V2_N1) {
    const int k1 = threadIdx.x / LINES, col = threadIdx.x - k1 * LINES;
    float2 v[V2_N2];
    if (col < ncol) {
#pragma unroll
      for (int l = 0; l < V2_N2; ++l) v[l] = z[col * V2Ex::SIZE + V2Ex::at(k1, l)];
      fft_small::phase2<INV, V2_N1, V2_N2>(v);
    }
    mbar_wait(bar, 0);                               // the operand tiles have landed
    if (col < ncol) {
      const int w = w0 + col;
      const bool dc_on = (STORE == ST_DC) ? (a.dcmask[w] != 0) : false;
      const float wgt = (STORE == ST_DC) ? __ldg(a.dcw) : 0.f;
#pragma unroll
      for (int k2 = 0; k2 < V2_N2; ++k2) {
        const int h = k1 + V2_N1 * k2;
        const long long hw = (long long)h * W + w;
        const long long off = b * HW + hw;
        const float2 o = make_float2(v[k2].x * a.scale, v[k2].y * a.scale);
        const float2 e0 = e0s[h * LINES + col];
        if (STORE == ST_DC) {
          float2 d = e0;
          if (dc_on) {
            const float2 k0 = e1s[h * LINES + col];
            d.x = e0.x - (e0.x - k0.x) * wgt;
            d.y = e0.y - (e0.y - k0.y) * wgt;
          }
          a.out_c[off] = make_float2(d.x - o.x, d.y - o.y);
        } else {
          if (a.out_u) a.out_u[off] = o;
          const float2 t = cmulc2(o, e0);
          a.out_p[(b * 2) * HW + hw] = t.x;
          a.out_p[(b * 2 + 1) * HW + hw] = t.y;
        }
      }
    }
  }
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (TmapEncodeFn)p;
  }();
  return fn;
}
// [B][H][W] array of 8-byte elements, box = [1][TMA_BOX_H][TMA_LINES]
bool make_tile_map(CUtensorMap* m, const void* base, int B, int H, int W) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc || ((uintptr_t)base & 15) || (W & 1)) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 8, (cuuint64_t)H * W * 8};
  const cuuint32_t box[3] = {TMA_LINES, TMA_BOX_H, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// SAN_FFT_TMA = 0: the register-batched epilogue loads of fft_cols_v2_kernel (A/B runs)
bool fft_tma_enabled() {
  static const bool on = [] { const char* e = getenv("SAN_FFT_TMA"); return e ? atoi(e) != 0 : true; }();
  return on;
}

template <bool INV, int STORE>
int launch_cols_tma(const FftArgs& a, cudaStream_t st, bool* done) {
  *done = false;
  if (!(STORE == ST_DC || STORE == ST_REDUCE)) return SAN_OK;
  if (!fft_tma_enabled() || a.H != V2_N || a.C != 1 || a.W % TMA_LINES != 0 || 2 * TMA_BOX_H != V2_N) return SAN_OK;
  FftTmaArgs ta;
  ta.a = a;
  const void* p0 = (STORE == ST_DC) ? (const void*)a.k : (const void*)a.sens;
  if (!make_tile_map(&ta.m0, p0, a.B, a.H, a.W)) return SAN_OK;
  if (STORE == ST_DC) {
    if (!make_tile_map(&ta.m1, a.k0, a.B, a.H, a.W)) return SAN_OK;
  } else {
    ta.m1 = ta.m0;
  }
  constexpr int kStore = (STORE == ST_DC || STORE == ST_REDUCE) ? STORE : ST_DC;     // (other stores never get here)
  auto kern = fft_cols_tma_kernel<INV, kStore>;
  constexpr int smem = tma_smem_bytes(kStore == ST_DC ? 2 : 1);
  static bool attr_set = false;                      // (one static per template instantiation)
  if (!attr_set) {
    SAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  kern<<<dim3(a.W / TMA_LINES, a.B), TMA_LINES * V2_N2, smem, st>>>(ta);
  SAN_LAUNCH_CHECK();
  *done = true;
  return SAN_OK;
}

// SAN_FFT_V2: 1 / unset = register FFT for 320-point lines, 8 columns per column-pass CTA (default since round 2:
// fft_expand_dc 0.147 -> 0.111 ms at bs 64, profiles/r2a_fft_v2_ab.txt); 2 = 16 columns per CTA; 0 = Stockham kernels only
int fft_v2_mode() {
  static const int mode = [] { const char* e = getenv("SAN_FFT_V2"); return e ? atoi(e) : 1; }();
  return mode;
}
bool fft_v2_enabled() { return fft_v2_mode() != 0; }

template <bool INV, int LOAD>
int launch_rows(const FftArgs& a, cudaStream_t st) {
  if (fft_v2_enabled() && a.W == V2_N) {
    const long long nrows = (long long)a.B * a.H;
    fft_rows_v2_kernel<INV, LOAD><<<san_cdiv(nrows, V2_LINES), V2_THREADS, 0, st>>>(a);
    SAN_LAUNCH_CHECK();
    return SAN_OK;
  }
  const size_t smem = ((size_t)2 * a.rpb * a.W + a.W) * sizeof(float2);
  auto kern = fft_rows_kernel<INV, LOAD>;
  if (smem > 48 * 1024)
    SAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nrows = (long long)a.B * a.H;
  kern<<<san_cdiv(nrows, a.rpb), a.rpb >= 8 ? 256 : 32 * a.rpb, smem, st>>>(a);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

template <bool INV, int STORE>
int launch_cols(const FftArgs& a, cudaStream_t st) {
  const bool reducing = (STORE == ST_REDUCE || STORE == ST_RSS);
  if (fft_v2_enabled() && a.H == V2_N && fft_v2_mode() != 2) {
    bool done = false;
    const int rc = launch_cols_tma<INV, STORE>(a, st, &done);
    if (rc != SAN_OK || done) return rc;
  }
  if (fft_v2_enabled() && a.H == V2_N) {
    const int lines = fft_v2_mode() == 2 ? 16 : 8;
    dim3 grid2(san_cdiv(a.W, lines), reducing ? a.B / a.C : a.B);
    if (reducing && a.C > 1) {
      if (lines == 16) fft_cols_v2_kernel<INV, STORE, true, 16><<<grid2, 16 * V2_N2, 0, st>>>(a);
      else fft_cols_v2_kernel<INV, STORE, true, 8><<<grid2, 8 * V2_N2, 0, st>>>(a);
    } else {
      if (lines == 16) fft_cols_v2_kernel<INV, STORE, false, 16><<<grid2, 16 * V2_N2, 0, st>>>(a);
      else fft_cols_v2_kernel<INV, STORE, false, 8><<<grid2, 8 * V2_N2, 0, st>>>(a);
    }
    SAN_LAUNCH_CHECK();
    return SAN_OK;
  }
  const int nbuf = (reducing && a.C > 1) ? 3 : 2;
  const size_t smem = ((size_t)nbuf * a.ct * (a.H + 1) + a.H) * sizeof(float2);
  dim3 grid(san_cdiv(a.W, a.ct), reducing ? a.B / a.C : a.B);
  if (a.ct == 8) {
    auto kern = fft_cols_kernel<INV, STORE, 8>;
    if (smem > 48 * 1024)
      SAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(a);
  } else {
    auto kern = fft_cols_kernel<INV, STORE, 4>;
    if (smem > 48 * 1024)
      SAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(a);
  }
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int run_fft(FftArgs& a, int inverse, int load, int store, cudaStream_t st) {
  SAN_CHECK_ARG(a.B > 0 && a.C > 0 && a.H > 0 && a.W > 0 && a.B % a.C == 0, "fft: bad dims B=%d C=%d H=%d W=%d", a.B, a.C, a.H, a.W);
  SAN_CHECK_ARG(a.B / a.C <= 65535, "fft: batch %d too large for grid.y", a.B / a.C);
  SAN_CHECK_ARG(make_plan(a.W, &a.planW) && make_plan(a.H, &a.planH), "fft: unsupported length (prime factor > 32) H=%d W=%d", a.H, a.W);
  int rc;
  if ((rc = get_twiddles(a.W, &a.twW)) != SAN_OK) return rc;
  if ((rc = get_twiddles(a.H, &a.twH)) != SAN_OK) return rc;
  a.scale *= (float)(1.0 / sqrt((double)a.H * (double)a.W));
  a.rpb = 2560 / a.W; if (a.rpb < 1) a.rpb = 1; if (a.rpb > 16) a.rpb = 16;
  a.ct = a.H <= 672 ? 8 : 4;   // 8 columns = 64 B row segments; 41 KB per CTA at H = 320 -> 5 CTAs per SM overlap load / FFT / store
  SAN_CHECK_ARG((size_t)3 * a.ct * (a.H + 1) * sizeof(float2) <= 200 * 1024 && (size_t)2 * a.rpb * a.W * sizeof(float2) <= 200 * 1024,
                "fft: line too long for shared memory H=%d W=%d", a.H, a.W);
#define ROWS(L) (inverse ? launch_rows<true, L>(a, st) : launch_rows<false, L>(a, st))
#define COLS(S) (inverse ? launch_cols<true, S>(a, st) : launch_cols<false, S>(a, st))
  switch (load) {
    case LD_C64: rc = ROWS(LD_C64); break;
    case LD_C64_COLMASK: rc = ROWS(LD_C64_COLMASK); break;
    case LD_PLANAR: rc = ROWS(LD_PLANAR); break;
    default: rc = ROWS(LD_PLANAR_S); break;
  }
  if (rc != SAN_OK) return rc;
  switch (store) {
    case ST_C64: rc = COLS(ST_C64); break;
    case ST_C64_COLMASK: rc = COLS(ST_C64_COLMASK); break;
    case ST_PLANAR: rc = COLS(ST_PLANAR); break;
    case ST_REDUCE: rc = COLS(ST_REDUCE); break;
    case ST_DC: rc = COLS(ST_DC); break;
    default: rc = COLS(ST_RSS); break;
  }
#undef ROWS
#undef COLS
  return rc;
}

}  // namespace

extern "C" {

size_t san_fft_workspace_bytes(int B, int H, int W) { return (size_t)B * H * W * sizeof(float2); }

int san_fft2(const void* in, int in_planar, const float* colmask_in, void* out, int out_planar,
             const float* colmask_out, void* tmp, int B, int H, int W, int inverse, void* stream) {
  SAN_CHECK_ARG(in && out && tmp, "san_fft2: null pointer");
  SAN_CHECK_ARG(!(in_planar && colmask_in) && !(out_planar && colmask_out), "san_fft2: column mask only with complex64 layout");
  FftArgs a{};
  a.in_c = (const float2*)in; a.in_p = (const float*)in; a.colmask = colmask_in ? colmask_in : colmask_out;
  SAN_CHECK_ARG(!(colmask_in && colmask_out), "san_fft2: one column mask at a time");
  a.tmp = (float2*)tmp; a.out_c = (float2*)out; a.out_p = (float*)out;
  a.B = B; a.C = 1; a.H = H; a.W = W; a.scale = 1.f;
  const int load = in_planar ? LD_PLANAR : (colmask_in ? LD_C64_COLMASK : LD_C64);
  const int store = out_planar ? ST_PLANAR : (colmask_out ? ST_C64_COLMASK : ST_C64);
  return run_fft(a, inverse, load, store, (cudaStream_t)stream);
}

int san_fft_reduce(const void* k, const void* sens, float* x_planar, void* u_out, void* tmp, int N, int C,
                   int H, int W, int inverse, float sign, void* stream) {
  SAN_CHECK_ARG(k && sens && x_planar && tmp, "san_fft_reduce: null pointer");
  FftArgs a{};
  a.in_c = (const float2*)k; a.sens = (const float2*)sens; a.out_p = x_planar; a.out_u = (float2*)u_out;
  a.tmp = (float2*)tmp; a.B = N * C; a.C = C; a.H = H; a.W = W; a.scale = sign;
  // note: out_u receives the scaled transform (sign included)
  return run_fft(a, inverse, LD_C64, ST_REDUCE, (cudaStream_t)stream);
}

int san_fft_expand_dc(const float* x_planar, const void* sens, const void* k, const void* k0,
                      const unsigned char* mask, const float* dc_weight, void* out, void* tmp, int N, int C,
                      int H, int W, int inverse, void* stream) {
  SAN_CHECK_ARG(x_planar && sens && out && tmp, "san_fft_expand_dc: null pointer");
  FftArgs a{};
  a.in_p = x_planar; a.sens = (const float2*)sens; a.k = (const float2*)k; a.k0 = (const float2*)k0;
  a.dcmask = mask; a.dcw = dc_weight; a.out_c = (float2*)out; a.tmp = (float2*)tmp;
  a.B = N * C; a.C = C; a.H = H; a.W = W; a.scale = 1.f;
  if (k) SAN_CHECK_ARG(k0 && mask && dc_weight, "san_fft_expand_dc: k given without k0/mask/dc_weight");
  return run_fft(a, inverse, LD_PLANAR_S, k ? ST_DC : ST_C64, (cudaStream_t)stream);
}

int san_fft_rss(const void* k, float* out, void* u_out, void* tmp, int N, int C, int H, int W, int inverse,
                void* stream) {
  SAN_CHECK_ARG(k && out && tmp, "san_fft_rss: null pointer");
  FftArgs a{};
  a.in_c = (const float2*)k; a.out_p = out; a.out_u = (float2*)u_out; a.tmp = (float2*)tmp;
  a.B = N * C; a.C = C; a.H = H; a.W = W; a.scale = 1.f;
  return run_fft(a, inverse, LD_C64, ST_RSS, (cudaStream_t)stream);
}

}  // extern "C"
