// Host emulation of the opt-in v2 FFT kernels (csrc/fft_v2.cuh): the SAME kernel source compiled by a plain
// host compiler with SAN_FFT_EMULATE, one OS thread per CUDA thread and a pthread barrier for __syncthreads(),
// one block after the other.  Every fused load / store variant the library launches is run on small batches
// of 320x320 slices and compared with a direct fp64 2-D DFT of the same arithmetic
// (reference signal_utils.py:4-12, varnet.py:395-402,486,508-530).  Built and run by tests/test_cpu_host.py:
//     g++ -std=c++17 -O2 -pthread -I/usr/local/cuda/include -DSAN_FFT_EMULATE tests/host/fft_v2_emul.cpp
#include <pthread.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

struct EmuDim { int x = 0, y = 0, z = 0; };
static thread_local EmuDim threadIdx;
static EmuDim blockIdx, blockDim, gridDim;
static pthread_barrier_t g_barrier;
static inline void __syncthreads() { pthread_barrier_wait(&g_barrier); }

#include "../../spatialalignmentnetwork_b200/csrc/fft_v2.cuh"

using namespace san_fft;
using cd = std::complex<double>;

// <<<grid(gx, gy), block(nt)>>> : blocks sequentially, threads concurrently
template <typename K>
static void launch(K kernel, int gx, int gy, int nt, const FftArgs& a) {
  blockDim.x = nt; gridDim.x = gx; gridDim.y = gy;
  pthread_barrier_init(&g_barrier, nullptr, nt);
  for (int by = 0; by < gy; ++by)
    for (int bx = 0; bx < gx; ++bx) {
      blockIdx.x = bx; blockIdx.y = by;
      std::vector<std::thread> th;
      th.reserve(nt);
      for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] { threadIdx.x = t; kernel(a); });
      for (auto& t : th) t.join();
    }
  pthread_barrier_destroy(&g_barrier);
}

static std::vector<float2> twiddles(int n) {
  std::vector<float2> t(n);
  for (int j = 0; j < n; ++j) { const double a = -2.0 * M_PI * j / n; t[j] = make_float2((float)cos(a), (float)sin(a)); }
  return t;
}

// direct separable fp64 2-D DFT (ortho scaling applied by the caller)
static std::vector<cd> dft2(const std::vector<cd>& x, int H, int W, bool inv) {
  std::vector<cd> t(H * W), o(H * W);
  const double sg = inv ? 2.0 : -2.0;
  std::vector<cd> wW(W), wH(H);
  for (int j = 0; j < W; ++j) wW[j] = std::polar(1.0, sg * M_PI * j / W);
  for (int j = 0; j < H; ++j) wH[j] = std::polar(1.0, sg * M_PI * j / H);
  for (int h = 0; h < H; ++h)
    for (int k = 0; k < W; ++k) {
      cd s = 0;
      for (int w = 0; w < W; ++w) s += x[h * W + w] * wW[(long long)w * k % W];
      t[h * W + k] = s;
    }
  for (int k = 0; k < H; ++k)
    for (int w = 0; w < W; ++w) {
      cd s = 0;
      for (int h = 0; h < H; ++h) s += t[h * W + w] * wH[(long long)h * k % H];
      o[k * W + w] = s;
    }
  return o;
}

static double frand() { return drand48() - 0.5; }
static double g_worst = 0.0;
static void track(const char* what, double err, double scale) {
  const double rel = err / scale;
  printf("%-34s max abs err %.3e (scale %.2f)\n", what, err, scale);
  if (rel > g_worst) g_worst = rel;
}

struct Case {
  int N, C, H, W;
  bool inv;
  std::vector<float2> in_c, sens, k, k0, tmp, out_c, out_u;
  std::vector<float> in_p, colmask, out_p;
  std::vector<unsigned char> dcmask;
  float dcw = 0.7f;
  std::vector<float2> twW, twH;
  FftArgs a{};
  Case(int N_, int C_, int H_, int W_, bool inv_) : N(N_), C(C_), H(H_), W(W_), inv(inv_) {
    const size_t B = (size_t)N * C, P = (size_t)H * W;
    in_c.resize(B * P); sens.resize(B * P); k.resize(B * P); k0.resize(B * P); tmp.resize(B * P); out_c.resize(B * P);
    out_u.resize(B * P); in_p.resize(B * 2 * P); out_p.resize(B * 2 * P); colmask.resize(W); dcmask.resize(W);
    for (auto& v : in_c) v = make_float2((float)frand(), (float)frand());
    for (auto& v : sens) v = make_float2((float)frand(), (float)frand());
    for (auto& v : k) v = make_float2((float)frand(), (float)frand());
    for (auto& v : k0) v = make_float2((float)frand(), (float)frand());
    for (auto& v : in_p) v = (float)frand();
    for (int w = 0; w < W; ++w) { colmask[w] = (w % 3 == 0) ? 0.f : 1.f; dcmask[w] = (w % 4 != 1); }
    twW = twiddles(W); twH = twiddles(H);
    a.in_c = in_c.data(); a.in_p = in_p.data(); a.sens = sens.data(); a.colmask = colmask.data(); a.tmp = tmp.data();
    a.out_c = out_c.data(); a.out_p = out_p.data(); a.out_u = out_u.data(); a.k = k.data(); a.k0 = k0.data();
    a.dcmask = dcmask.data(); a.dcw = &dcw; a.B = (int)B; a.C = C; a.H = H; a.W = W;
    a.scale = (float)(1.0 / sqrt((double)H * W)); a.twW = twW.data(); a.twH = twH.data();
  }
  // LINES: columns per column-pass CTA (8 or 16), alternated between the cases below
  template <int LOAD, int STORE, bool MULTI, int LINES = 8>
  void run() {
    const bool reducing = (STORE == ST_REDUCE || STORE == ST_RSS);
    const int nrows = a.B * H;
    if (inv) {
      launch(fft_rows_v2_kernel<true, LOAD>, (nrows + V2_LINES - 1) / V2_LINES, 1, V2_THREADS, a);
      launch(fft_cols_v2_kernel<true, STORE, MULTI, LINES>, (W + LINES - 1) / LINES, reducing ? a.B / C : a.B, LINES * V2_N2, a);
    } else {
      launch(fft_rows_v2_kernel<false, LOAD>, (nrows + V2_LINES - 1) / V2_LINES, 1, V2_THREADS, a);
      launch(fft_cols_v2_kernel<false, STORE, MULTI, LINES>, (W + LINES - 1) / LINES, reducing ? a.B / C : a.B, LINES * V2_N2, a);
    }
  }
  // reference transform of slice b of the loaded input (LOAD semantics), ortho-scaled times `extra`
  std::vector<cd> ref(int b, int LOAD, double extra) const {
    const size_t P = (size_t)H * W;
    std::vector<cd> x(P);
    for (size_t i = 0; i < P; ++i) {
      const int w = (int)(i % W);
      if (LOAD == LD_C64) x[i] = cd(in_c[b * P + i].x, in_c[b * P + i].y);
      else if (LOAD == LD_C64_COLMASK) x[i] = cd(in_c[b * P + i].x, in_c[b * P + i].y) * (double)colmask[w];
      else {
        const int g = (LOAD == LD_PLANAR_S) ? b / C : b;
        x[i] = cd(in_p[(size_t)(g * 2) * P + i], in_p[(size_t)(g * 2 + 1) * P + i]);
        if (LOAD == LD_PLANAR_S) x[i] *= cd(sens[b * P + i].x, sens[b * P + i].y);
      }
    }
    auto X = dft2(x, H, W, inv);
    const double s = extra / sqrt((double)H * W);
    for (auto& v : X) v *= s;
    return X;
  }
};

int main() {
  srand48(11);
  const int H = 320, W = 320;
  {  // plain fft2 / ifft2 (signal_utils.py:4-12)
    for (int inv = 0; inv < 2; ++inv) {
      Case c(2, 1, H, W, inv != 0);
      c.run<LD_C64, ST_C64, false>();
      double e = 0;
      for (int b = 0; b < 2; ++b) {
        auto X = c.ref(b, LD_C64, 1.0);
        for (size_t i = 0; i < X.size(); ++i) e = fmax(e, std::abs(X[i] - cd(c.out_c[b * X.size() + i].x, c.out_c[b * X.size() + i].y)));
      }
      track(inv ? "ifft2 c64->c64" : "fft2 c64->c64", e, 0.5);
    }
  }
  {  // planar(ifft2(colmask * k)) and its adjoint (varnet.py:395-407)
    Case c(2, 1, H, W, true);
    c.run<LD_C64_COLMASK, ST_PLANAR, false, 16>();
    double e = 0;
    const size_t P = (size_t)H * W;
    for (int b = 0; b < 2; ++b) {
      auto X = c.ref(b, LD_C64_COLMASK, 1.0);
      for (size_t i = 0; i < P; ++i)
        e = fmax(e, std::abs(X[i] - cd(c.out_p[(size_t)(b * 2) * P + i], c.out_p[(size_t)(b * 2 + 1) * P + i])));
    }
    track("ifft2 colmask -> planar", e, 0.5);
    Case d(2, 1, H, W, false);
    d.run<LD_PLANAR, ST_C64_COLMASK, false>();
    e = 0;
    for (int b = 0; b < 2; ++b) {
      auto X = d.ref(b, LD_PLANAR, 1.0);
      for (size_t i = 0; i < P; ++i)
        e = fmax(e, std::abs(X[i] * (double)d.colmask[i % W] - cd(d.out_c[b * P + i].x, d.out_c[b * P + i].y)));
    }
    track("fft2 planar -> colmask", e, 0.5);
  }
  {  // sens_expand + soft DC (varnet.py:508-509,525-530), 2 coils
    Case c(1, 2, H, W, false);
    c.run<LD_PLANAR_S, ST_DC, false, 16>();
    double e = 0;
    const size_t P = (size_t)H * W;
    for (int b = 0; b < 2; ++b) {
      auto X = c.ref(b, LD_PLANAR_S, 1.0);
      for (size_t i = 0; i < P; ++i) {
        const cd kk(c.k[b * P + i].x, c.k[b * P + i].y), k0(c.k0[b * P + i].x, c.k0[b * P + i].y);
        const cd dc = c.dcmask[i % W] ? (kk - k0) * (double)c.dcw : cd(0, 0);
        e = fmax(e, std::abs((kk - dc - X[i]) - cd(c.out_c[b * P + i].x, c.out_c[b * P + i].y)));
      }
    }
    track("expand + soft DC (2 coils)", e, 0.5);
  }
  for (int C = 1; C <= 2; ++C) {  // sens_reduce (varnet.py:511-512) and rss(ifft2(k)) (varnet.py:486)
    Case c(1, C, H, W, true);
    if (C == 1) c.run<LD_C64, ST_REDUCE, false, 16>(); else c.run<LD_C64, ST_REDUCE, true, 16>();
    const size_t P = (size_t)H * W;
    std::vector<cd> acc(P, cd(0, 0));
    double e = 0, eu = 0;
    for (int b = 0; b < C; ++b) {
      auto X = c.ref(b, LD_C64, 1.0);
      for (size_t i = 0; i < P; ++i) {
        acc[i] += X[i] * std::conj(cd(c.sens[b * P + i].x, c.sens[b * P + i].y));
        eu = fmax(eu, std::abs(X[i] - cd(c.out_u[b * P + i].x, c.out_u[b * P + i].y)));
      }
    }
    for (size_t i = 0; i < P; ++i) e = fmax(e, std::abs(acc[i] - cd(c.out_p[i], c.out_p[P + i])));
    track(C == 1 ? "reduce (1 coil)" : "reduce (2 coils, MULTI)", fmax(e, eu), 0.5);
    Case r(1, C, H, W, true);
    if (C == 1) r.run<LD_C64, ST_RSS, false>(); else r.run<LD_C64, ST_RSS, true>();
    std::vector<double> ss(P, 0.0);
    for (int b = 0; b < C; ++b) {
      auto X = r.ref(b, LD_C64, 1.0);
      for (size_t i = 0; i < P; ++i) ss[i] += std::norm(X[i]);
    }
    e = 0;
    for (size_t i = 0; i < P; ++i) e = fmax(e, fabs(sqrt(ss[i]) - r.out_p[i]));
    track(C == 1 ? "rss(ifft2) (1 coil)" : "rss(ifft2) (2 coils, MULTI)", e, 0.5);
  }
  {  // column pass alone on a narrow image (W = 20: last CTA has 4 of 8 columns): tmp = row-transformed input
    const int Wn = 20;
    Case c(2, 1, H, Wn, false);
    const size_t P = (size_t)H * Wn;
    std::vector<cd> x(P);
    std::vector<std::vector<cd>> refs;
    for (int b = 0; b < 2; ++b) {
      for (size_t i = 0; i < P; ++i) x[i] = cd(c.in_c[b * P + i].x, c.in_c[b * P + i].y);
      // row pass on the host (fp64), column pass by the emulated kernel
      std::vector<cd> t(P);
      for (int h = 0; h < H; ++h)
        for (int k = 0; k < Wn; ++k) {
          cd s = 0;
          for (int w = 0; w < Wn; ++w) s += x[h * Wn + w] * std::polar(1.0, -2.0 * M_PI * ((long long)w * k % Wn) / Wn);
          t[h * Wn + k] = s;
          c.tmp[b * P + h * Wn + k] = make_float2((float)s.real(), (float)s.imag());
        }
      refs.push_back(dft2(x, H, Wn, false));
    }
    launch(fft_cols_v2_kernel<false, ST_C64, false, 8>, (Wn + 7) / 8, 2, 8 * V2_N2, c.a);
    {
      double e8 = 0;
      for (int b = 0; b < 2; ++b)
        for (size_t i = 0; i < P; ++i)
          e8 = fmax(e8, std::abs(refs[b][i] * (double)c.a.scale - cd(c.out_c[b * P + i].x, c.out_c[b * P + i].y)));
      track("column pass, W = 20, 8 per CTA", e8, 0.5);
      for (auto& v : c.out_c) v = make_float2(0.f, 0.f);
    }
    launch(fft_cols_v2_kernel<false, ST_C64, false, 16>, (Wn + 15) / 16, 2, 16 * V2_N2, c.a);
    double e = 0;
    for (int b = 0; b < 2; ++b)
      for (size_t i = 0; i < P; ++i)
        e = fmax(e, std::abs(refs[b][i] * (double)c.a.scale - cd(c.out_c[b * P + i].x, c.out_c[b * P + i].y)));
    track("column pass, W = 20, 16 per CTA", e, 0.5);
  }
  printf("worst relative error %.3e\n", g_worst);
  return g_worst < 2e-5 ? 0 : 1;
}
