// Register-resident small DFTs (4, 5, 16 = 4x4, 20 = 4x5) and the two-phase N = N1 * N2 decomposition used by
// the v2 FFT kernels of fft.cu.  Host + device: tests/host/fft_small_test.cu runs exactly this code on the CPU
// against a direct fp64 DFT (tests/test_cpu_host.py::test_fft_small_host).
//
//   n = N2*j + l (j < N1, l < N2),  k = k1 + N1*k2 (k1 < N1, k2 < N2):
//   X[k1 + N1*k2] = sum_l ( w_N^{l*k1} * sum_j x[N2*j + l] w_N1^{j*k1} ) * w_N2^{l*k2}
// phase 1: one thread per (line, l): N1-point DFT over j in registers, times the twiddle w_N^{l*k1};
// phase 2: one thread per (line, k1): N2-point DFT over l in registers; outputs land in natural order.
// All twiddles are forward, exp(-2*pi*i*m/N); INV conjugates them.
#pragma once
#include <vector_types.h>
#include <vector_functions.h>

#if defined(__CUDACC__)
#define FS_HD __host__ __device__ __forceinline__
#else
#define FS_HD inline
#endif

namespace fft_small {

FS_HD constexpr float cos_16(int m) {
  constexpr float t[16] = {1.000000000e+00f, 9.238795325e-01f, 7.071067812e-01f, 3.826834324e-01f, 6.123233996e-17f, -3.826834324e-01f, -7.071067812e-01f, -9.238795325e-01f, -1.000000000e+00f, -9.238795325e-01f, -7.071067812e-01f, -3.826834324e-01f, -1.836970199e-16f, 3.826834324e-01f, 7.071067812e-01f, 9.238795325e-01f};
  return t[m];
}
FS_HD constexpr float sin_16(int m) {
  constexpr float t[16] = {0.000000000e+00f, 3.826834324e-01f, 7.071067812e-01f, 9.238795325e-01f, 1.000000000e+00f, 9.238795325e-01f, 7.071067812e-01f, 3.826834324e-01f, 1.224646799e-16f, -3.826834324e-01f, -7.071067812e-01f, -9.238795325e-01f, -1.000000000e+00f, -9.238795325e-01f, -7.071067812e-01f, -3.826834324e-01f};
  return t[m];
}
FS_HD constexpr float cos_20(int m) {
  constexpr float t[20] = {1.000000000e+00f, 9.510565163e-01f, 8.090169944e-01f, 5.877852523e-01f, 3.090169944e-01f, 6.123233996e-17f, -3.090169944e-01f, -5.877852523e-01f, -8.090169944e-01f, -9.510565163e-01f, -1.000000000e+00f, -9.510565163e-01f, -8.090169944e-01f, -5.877852523e-01f, -3.090169944e-01f, -1.836970199e-16f, 3.090169944e-01f, 5.877852523e-01f, 8.090169944e-01f, 9.510565163e-01f};
  return t[m];
}
FS_HD constexpr float sin_20(int m) {
  constexpr float t[20] = {0.000000000e+00f, 3.090169944e-01f, 5.877852523e-01f, 8.090169944e-01f, 9.510565163e-01f, 1.000000000e+00f, 9.510565163e-01f, 8.090169944e-01f, 5.877852523e-01f, 3.090169944e-01f, 1.224646799e-16f, -3.090169944e-01f, -5.877852523e-01f, -8.090169944e-01f, -9.510565163e-01f, -1.000000000e+00f, -9.510565163e-01f, -8.090169944e-01f, -5.877852523e-01f, -3.090169944e-01f};
  return t[m];
}
FS_HD constexpr float cos_5(int m) {
  constexpr float t[5] = {1.000000000e+00f, 3.090169944e-01f, -8.090169944e-01f, -8.090169944e-01f, 3.090169944e-01f};
  return t[m];
}
FS_HD constexpr float sin_5(int m) {
  constexpr float t[5] = {0.000000000e+00f, 9.510565163e-01f, 5.877852523e-01f, -5.877852523e-01f, -9.510565163e-01f};
  return t[m];
}

// a * exp(-+ 2*pi*i*m/N) given c = cos, s = sin of 2*pi*m/N (forward: a * (c - i s); inverse: a * (c + i s))
template <bool INV>
FS_HD float2 rot(float2 a, float c, float s) {
  return INV ? make_float2(a.x * c - a.y * s, a.x * s + a.y * c) : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}
// multiply by -i (forward) / +i (inverse)
template <bool INV>
FS_HD float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// 4-point DFT, natural order in and out
template <bool INV>
FS_HD void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
  const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), a3 = mul_mi<INV>(make_float2(v1.x - v3.x, v1.y - v3.y));
  v0 = make_float2(a0.x + a2.x, a0.y + a2.y);
  v1 = make_float2(a1.x + a3.x, a1.y + a3.y);
  v2 = make_float2(a0.x - a2.x, a0.y - a2.y);
  v3 = make_float2(a1.x - a3.x, a1.y - a3.y);
}

// 5-point DFT, natural order in and out
template <bool INV>
FS_HD void dft5(float2& v0, float2& v1, float2& v2, float2& v3, float2& v4) {
  constexpr float c1 = cos_5(1), c2 = cos_5(2);
  const float s1 = INV ? sin_5(1) : -sin_5(1), s2 = INV ? sin_5(2) : -sin_5(2);   // w = c + i s, sign carries INV
  const float2 p1 = make_float2(v1.x + v4.x, v1.y + v4.y), d1 = make_float2(v1.x - v4.x, v1.y - v4.y);
  const float2 p2 = make_float2(v2.x + v3.x, v2.y + v3.y), d2 = make_float2(v2.x - v3.x, v2.y - v3.y);
  const float2 r1 = make_float2(v0.x + c1 * p1.x + c2 * p2.x, v0.y + c1 * p1.y + c2 * p2.y);
  const float2 r2 = make_float2(v0.x + c2 * p1.x + c1 * p2.x, v0.y + c2 * p1.y + c1 * p2.y);
  const float2 q1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
  const float2 q2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
  v0 = make_float2(v0.x + p1.x + p2.x, v0.y + p1.y + p2.y);
  // X1 = r1 + i q1, X4 = r1 - i q1, X2 = r2 + i q2, X3 = r2 - i q2   (i q = (-q.y, q.x))
  v1 = make_float2(r1.x - q1.y, r1.y + q1.x);
  v4 = make_float2(r1.x + q1.y, r1.y - q1.x);
  v2 = make_float2(r2.x - q2.y, r2.y + q2.x);
  v3 = make_float2(r2.x + q2.y, r2.y - q2.x);
}

// 16-point DFT in place, natural order: n = 4*n1 + n2, k = k1 + 4*k2
template <bool INV>
FS_HD void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);      // over n1 -> k1 at v[4*k1 + n2]
#pragma unroll
  for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
    for (int n2 = 1; n2 < 4; ++n2) v[4 * k1 + n2] = rot<INV>(v[4 * k1 + n2], cos_16(n2 * k1), sin_16(n2 * k1));
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // over n2 -> k2
  // now v[4*k1 + k2] = X[k1 + 4*k2]: transpose to natural order
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) { const float2 t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
}

// 20-point DFT in place, natural order: n = 5*n1 + n2 (n1 < 4, n2 < 5), k = k1 + 4*k2 (k1 < 4, k2 < 5)
template <bool INV>
FS_HD void dft20(float2 (&v)[20]) {
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) dft4<INV>(v[n2], v[5 + n2], v[10 + n2], v[15 + n2]);     // over n1 -> k1 at v[5*k1 + n2]
#pragma unroll
  for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
    for (int n2 = 1; n2 < 5; ++n2) v[5 * k1 + n2] = rot<INV>(v[5 * k1 + n2], cos_20(n2 * k1), sin_20(n2 * k1));
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft5<INV>(v[5 * k1], v[5 * k1 + 1], v[5 * k1 + 2], v[5 * k1 + 3], v[5 * k1 + 4]);
  // v[5*k1 + k2] = X[k1 + 4*k2] -> natural order
  float2 t[20];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) t[k1 + 4 * k2] = v[5 * k1 + k2];
#pragma unroll
  for (int i = 0; i < 20; ++i) v[i] = t[i];
}

template <bool INV, int R>
FS_HD void dft(float2 (&v)[R]);
template <> FS_HD void dft<false, 16>(float2 (&v)[16]) { dft16<false>(v); }
template <> FS_HD void dft<true, 16>(float2 (&v)[16]) { dft16<true>(v); }
template <> FS_HD void dft<false, 20>(float2 (&v)[20]) { dft20<false>(v); }
template <> FS_HD void dft<true, 20>(float2 (&v)[20]) { dft20<true>(v); }

// shared-memory exchange layout of one line between the phases: element (k1, l) at k1 * (N2 + 1) + l.
// Bank behaviour (8-byte elements, 16 per wavefront half): the row pass walks l (phase 1) and k1 (phase 2, stride
// LD = 21 = 5 mod 16, coprime) within a line; the column pass puts 8 adjacent columns = 8 LINES in consecutive
// threads, so the line pitch SIZE must not be a multiple of 16 elements (336 would be an 8-way conflict): it is
// padded to 2 mod 16, which makes the 16 threads of a half-warp (2 values of l or k1 x 8 lines) hit 16 distinct
// bank pairs in both phases.
template <int N1, int N2>
struct Exchange {
  static constexpr int LD = N2 + 1;
  static constexpr int SIZE = N1 * LD + (18 - (N1 * LD) % 16) % 16;
  FS_HD static int at(int k1, int l) { return k1 * LD + l; }
};

// phase 1 of line element l: v[j] = x[N2*j + l] on entry; on exit v[k1] = w_N^{l*k1} * DFT_N1(v)[k1].
// tw = forward twiddle table of length N = N1*N2 (exp(-2*pi*i*m/N)).
template <bool INV, int N1, int N2>
FS_HD void phase1(float2 (&v)[N1], int l, const float2* tw) {
  dft<INV, N1>(v);
#pragma unroll
  for (int k1 = 1; k1 < N1; ++k1) {
    const float2 w = tw[l * k1];               // l*k1 < N2*N1
    v[k1] = INV ? make_float2(v[k1].x * w.x + v[k1].y * w.y, v[k1].y * w.x - v[k1].x * w.y)
                : make_float2(v[k1].x * w.x - v[k1].y * w.y, v[k1].x * w.y + v[k1].y * w.x);
  }
}

// phase 2 for k1: v[l] = exchanged element (k1, l) on entry; on exit v[k2] = X[k1 + N1*k2].
template <bool INV, int N1, int N2>
FS_HD void phase2(float2 (&v)[N2]) { dft<INV, N2>(v); }

}  // namespace fft_small
