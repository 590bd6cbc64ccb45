// Types shared by the FFT kernels of fft.cu and the opt-in v2 line-FFT kernels (length 320 = 16 x 20 in registers).
//
// This header compiles in two worlds:
//   * nvcc (included by fft.cu): real kernels;
//   * a plain host compiler with SAN_FFT_EMULATE defined (tests/host/fft_v2_emul.cpp): the SAME kernel source runs
//     with one OS thread per CUDA thread and a pthread barrier for __syncthreads(), block after block, and is
//     checked against a direct fp64 2-D DFT for every fused load / store variant.  The harness defines
//     threadIdx / blockIdx / blockDim / gridDim and __syncthreads() before including this file.
#pragma once
#include "fft_small.cuh"

#ifdef SAN_FFT_EMULATE
#define SAN_GLOBAL
#define SAN_SHARED static
#define SAN_LAUNCH_BOUNDS(n)
#define SAN_LDG(p) (*(p))
#define SAN_DEVICE inline
#else
#define SAN_GLOBAL __global__
#define SAN_SHARED __shared__
#define SAN_LAUNCH_BOUNDS(n) __launch_bounds__(n)
#define SAN_LDG(p) __ldg(p)
#define SAN_DEVICE __device__ __forceinline__
#endif

namespace san_fft {

struct FftPlan {
  int n;
  int ns;
  int radix[12];
  int shift[12];    // log2(Ns) of the stage when Ns (product of the previous radices) is a power of two, else -1
  int twstep[12];   // n / (Ns * R): twiddle index step of the stage
};

SAN_DEVICE float2 cmul2(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
SAN_DEVICE float2 cmulc2(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

// copy the twiddle table of length n into shared memory (all threads)
SAN_DEVICE void load_twiddles(float2* dst, const float2* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = SAN_LDG(src + i);
}

enum { LD_C64 = 0, LD_C64_COLMASK = 1, LD_PLANAR = 2, LD_PLANAR_S = 3 };
enum { ST_C64 = 0, ST_C64_COLMASK = 1, ST_PLANAR = 2, ST_REDUCE = 3, ST_DC = 4, ST_RSS = 5 };

struct FftArgs {
  const float2* in_c;
  const float* in_p;
  const float2* sens;
  const float* colmask;
  float2* tmp;
  float2* out_c;
  float* out_p;
  float2* out_u;
  const float2* k;
  const float2* k0;
  const unsigned char* dcmask;
  const float* dcw;
  int B, C, H, W;
  float scale;
  const float2* twW;
  const float2* twH;
  FftPlan planW, planH;
  int rpb;  // rows per block (row pass)
  int ct;   // columns per block (column pass)
};


// ------------------------------------------------------------------------------------------------------------
// v2 line FFT for the benchmark length 320 = 16 x 20 (default; SAN_FFT_V2=0 selects the Stockham kernels; measured 1.33-2.3x over them, profiles/r2a_fft_v2_ab.txt).
// The Stockham kernels above make 4 shared-memory round trips per line with ~2000 warp instructions per line
// (ncu: instruction / latency bound at 0.26 of the HBM roofline).  Here every thread owns one sub-transform in
// REGISTERS (csrc/fft_small.cuh, checked on the host by tests/host/fft_small_test.cu): phase 1 = 16-point DFT of the
// stride-20 elements it loads straight from global memory + twiddle, ONE shared-memory exchange, phase 2 = 20-point
// DFT whose outputs go straight to global memory through the same fused epilogues.  ~400 warp instructions per
// line, 16 / 20 independent loads in flight per thread.  Same load / store variants, same `tmp` layout, so a
// v2 row pass combines with a v1 column pass (and vice versa) when only one dimension is 320.
constexpr int V2_N1 = 16, V2_N2 = 20, V2_N = V2_N1 * V2_N2;
constexpr int V2_LINES = 8;                        // rows (row pass) / adjacent columns (column pass) per CTA
constexpr int V2_THREADS = V2_LINES * V2_N2;       // 160: phase 1 uses all, phase 2 the first V2_LINES * V2_N1
using V2Ex = fft_small::Exchange<V2_N1, V2_N2>;

template <bool INV, int LOAD>
SAN_GLOBAL void SAN_LAUNCH_BOUNDS(V2_THREADS) fft_rows_v2_kernel(const FftArgs a) {
  SAN_SHARED float2 z[V2_LINES * V2Ex::SIZE];
  SAN_SHARED float2 tws[V2_N];
  constexpr int W = V2_N;
  const long long nrows = (long long)a.B * a.H;
  const long long row0 = (long long)blockIdx.x * V2_LINES;
  const long long HW = (long long)a.H * W;
  {
    const int r = threadIdx.x / V2_N2, l = threadIdx.x - r * V2_N2;
    const long long row = row0 + r;
    float2 v[V2_N1];
    // the line's elements are requested FIRST, the twiddle table is copied while they are in flight: with the copy (a
    // dependent global -> shared round trip) and its barrier in front, every CTA paid two memory latencies in series
    if (row < nrows) {
      const long long rowoff = row * W;
      const long long b = row / a.H;
      const long long g = (LOAD == LD_PLANAR_S) ? b / a.C : b;
      const long long hw0 = rowoff - b * HW;
      float2 sv[LOAD == LD_PLANAR_S ? V2_N1 : 1];
#pragma unroll
      for (int j = 0; j < V2_N1; ++j) {
        const int w = V2_N2 * j + l;
        if (LOAD == LD_C64) {
          v[j] = SAN_LDG(a.in_c + rowoff + w);
        } else if (LOAD == LD_C64_COLMASK) {
          v[j] = SAN_LDG(a.in_c + rowoff + w);
          const float m = SAN_LDG(a.colmask + w);
          v[j].x *= m; v[j].y *= m;
        } else {
          v[j].x = SAN_LDG(a.in_p + (g * 2) * HW + hw0 + w);
          v[j].y = SAN_LDG(a.in_p + (g * 2 + 1) * HW + hw0 + w);
          if (LOAD == LD_PLANAR_S) sv[j] = SAN_LDG(a.sens + rowoff + w);
        }
      }
      if (LOAD == LD_PLANAR_S) {
#pragma unroll
        for (int j = 0; j < V2_N1; ++j) v[j] = cmul2(v[j], sv[LOAD == LD_PLANAR_S ? j : 0]);
      }
    }
    load_twiddles(tws, a.twW, V2_N);
    __syncthreads();
    if (row < nrows) {
      fft_small::phase1<INV, V2_N1, V2_N2>(v, l, tws);
#pragma unroll
      for (int k1 = 0; k1 < V2_N1; ++k1) z[r * V2Ex::SIZE + V2Ex::at(k1, l)] = v[k1];
    }
  }
  __syncthreads();
  if (threadIdx.x < V2_LINES * V2_N1) {
    const int r = threadIdx.x / V2_N1, k1 = threadIdx.x - r * V2_N1;
    const long long row = row0 + r;
    if (row < nrows) {
      float2 v[V2_N2];
#pragma unroll
      for (int l = 0; l < V2_N2; ++l) v[l] = z[r * V2Ex::SIZE + V2Ex::at(k1, l)];
      fft_small::phase2<INV, V2_N1, V2_N2>(v);
      float2* dst = a.tmp + row * W + k1;
#pragma unroll
      for (int k2 = 0; k2 < V2_N2; ++k2) dst[V2_N1 * k2] = v[k2];
    }
  }
}

// grid: (ceil(W / LINES), G), G = N for the coil-reducing stores and B otherwise; H == 320, any W; LINES * 20 threads.
// LINES = adjacent columns per CTA: 8 (64 B global segments, 24 KB smem) or 16 (128 B segments, 45 KB smem).
// MULTI: coil-reducing store with C > 1 (per-thread register accumulators across the coil loop).
template <bool INV, int STORE, bool MULTI, int LINES>
SAN_GLOBAL void SAN_LAUNCH_BOUNDS(LINES * V2_N2) fft_cols_v2_kernel(const FftArgs a) {
  SAN_SHARED float2 z[LINES * V2Ex::SIZE];
  SAN_SHARED float2 tws[V2_N];
  constexpr int H = V2_N;
  const int W = a.W;
  const int w0 = blockIdx.x * LINES;
  const int ncol = (W - w0) < LINES ? (W - w0) : LINES;
  const long long HW = (long long)H * W;
  constexpr bool reducing = (STORE == ST_REDUCE || STORE == ST_RSS);
  const int ncoil = MULTI ? a.C : 1;
  const long long g = blockIdx.y;
  float2 acc[MULTI ? V2_N2 : 1];                     // coil accumulation of this thread's outputs (registers)
#pragma unroll
  for (int i = 0; i < (MULTI ? V2_N2 : 1); ++i) acc[i] = make_float2(0.f, 0.f);
  for (int c = 0; c < ncoil; ++c) {
    const long long b = (reducing && MULTI) ? g * a.C + c : g;
    const float2* src = a.tmp + b * HW;
    {
      const int l = threadIdx.x / LINES, col = threadIdx.x - l * LINES;
      float2 v[V2_N1];
      // column elements requested before the twiddle copy / the barrier (see the row pass); plain loads: `tmp` may alias
      // the output buffer (adjoint use), the row pass has written it in this stream
      if (col < ncol) {
#pragma unroll
        for (int j = 0; j < V2_N1; ++j) v[j] = src[(long long)(V2_N2 * j + l) * W + w0 + col];
      }
      if (c == 0) load_twiddles(tws, a.twH, V2_N);
      __syncthreads();                               // twiddles visible / previous coil's phase 2 done with z
      if (col < ncol) {
        fft_small::phase1<INV, V2_N1, V2_N2>(v, l, tws);
#pragma unroll
        for (int k1 = 0; k1 < V2_N1; ++k1) z[col * V2Ex::SIZE + V2Ex::at(k1, l)] = v[k1];
      }
    }
    __syncthreads();
    if (threadIdx.x < LINES * V2_N1) {
      const int k1 = threadIdx.x / LINES, col = threadIdx.x - k1 * LINES;
      if (col < ncol) {
        float2 v[V2_N2];
#pragma unroll
        for (int l = 0; l < V2_N2; ++l) v[l] = z[col * V2Ex::SIZE + V2Ex::at(k1, l)];
        fft_small::phase2<INV, V2_N1, V2_N2>(v);
        const int w = w0 + col;
        // The operands the epilogue reads next to the transform (k, k0, sens) are fetched in BATCHES of V2_EPB through the
        // read-only path BEFORE the batch's stores: written as load-use-store per element, the (non-restrict) stores kept
        // every later load behind them - 20 to 40 dependent HBM round trips per thread, which made the soft-DC column pass
        // 3.4x slower than the plain one (71 vs 21 us, profiles/r2a_fft_v2_ab.txt).
        constexpr int V2_EPB = 10;
        const long long col_off = b * HW + w;            // + h * W per element
        const bool dc_on = (STORE == ST_DC) ? (a.dcmask[w] != 0) : false;
        const float wgt = (STORE == ST_DC) ? SAN_LDG(a.dcw) : 0.f;
        const float cm = (STORE == ST_C64_COLMASK) ? a.colmask[w] : 1.f;
#pragma unroll
        for (int kb = 0; kb < V2_N2; kb += V2_EPB) {
          float2 e0[V2_EPB], e1[V2_EPB];
          if (STORE == ST_DC) {
#pragma unroll
            for (int i = 0; i < V2_EPB; ++i) e0[i] = SAN_LDG(a.k + col_off + (long long)(k1 + V2_N1 * (kb + i)) * W);
            if (dc_on) {
#pragma unroll
              for (int i = 0; i < V2_EPB; ++i) e1[i] = SAN_LDG(a.k0 + col_off + (long long)(k1 + V2_N1 * (kb + i)) * W);
            }
          } else if (STORE == ST_REDUCE) {
#pragma unroll
            for (int i = 0; i < V2_EPB; ++i) e0[i] = SAN_LDG(a.sens + col_off + (long long)(k1 + V2_N1 * (kb + i)) * W);
          }
#pragma unroll
          for (int i = 0; i < V2_EPB; ++i) {
            const int k2 = kb + i;
            const int h = k1 + V2_N1 * k2;
            const long long hw = (long long)h * W + w;
            const long long off = b * HW + hw;
            float2 o = make_float2(v[k2].x * a.scale, v[k2].y * a.scale);
            if (STORE == ST_C64) {
              a.out_c[off] = o;
            } else if (STORE == ST_C64_COLMASK) {
              a.out_c[off] = make_float2(o.x * cm, o.y * cm);
            } else if (STORE == ST_PLANAR) {
              a.out_p[(b * 2) * HW + hw] = o.x;
              a.out_p[(b * 2 + 1) * HW + hw] = o.y;
            } else if (STORE == ST_DC) {
              const float2 kk = e0[i];
              float2 d = kk;
              if (dc_on) {
                const float2 k0 = e1[i];
                d.x = kk.x - (kk.x - k0.x) * wgt;
                d.y = kk.y - (kk.y - k0.y) * wgt;
              }
              a.out_c[off] = make_float2(d.x - o.x, d.y - o.y);
            } else if (STORE == ST_REDUCE) {
              if (a.out_u) a.out_u[off] = o;
              float2 t = cmulc2(o, e0[i]);
              if (MULTI) {
                t.x += acc[MULTI ? k2 : 0].x; t.y += acc[MULTI ? k2 : 0].y;
                acc[MULTI ? k2 : 0] = t;
              }
              if (c + 1 == ncoil) {
                a.out_p[(g * 2) * HW + hw] = t.x;
                a.out_p[(g * 2 + 1) * HW + hw] = t.y;
              }
            } else if (STORE == ST_RSS) {
              if (a.out_u) a.out_u[off] = o;
              float t = o.x * o.x + o.y * o.y;
              if (MULTI) {
                t += acc[MULTI ? k2 : 0].x;
                acc[MULTI ? k2 : 0].x = t;
              }
              if (c + 1 == ncoil) a.out_p[g * HW + hw] = sqrtf(t);
            }
          }
        }
      }
    }
  }
}

}  // namespace san_fft
