// Host check of csrc/fft_small.cuh: the small DFTs and the two-phase N1 x N2 decomposition (exactly the code the
// v2 FFT kernels run per thread) against a direct fp64 DFT.  Built and run by tests/test_cpu_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../spatialalignmentnetwork_b200/csrc/fft_small.cuh"

using namespace fft_small;

static double maxerr = 0.0;

template <bool INV, int R>
void check_small() {
  float2 v[R];
  double xr[R], xi[R];
  for (int i = 0; i < R; ++i) { xr[i] = drand48() - 0.5; xi[i] = drand48() - 0.5; v[i] = make_float2((float)xr[i], (float)xi[i]); }
  dft<INV, R>(v);
  for (int k = 0; k < R; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < R; ++n) {
      const double a = (INV ? 2.0 : -2.0) * M_PI * (double)(n * k % R) / R;
      sr += xr[n] * cos(a) - xi[n] * sin(a);
      si += xr[n] * sin(a) + xi[n] * cos(a);
    }
    maxerr = fmax(maxerr, fmax(fabs(sr - v[k].x), fabs(si - v[k].y)));
  }
}

template <bool INV, int N1, int N2>
void check_line() {
  constexpr int N = N1 * N2;
  std::vector<float2> tw(N), x(N), z(Exchange<N1, N2>::SIZE), X(N);
  for (int m = 0; m < N; ++m) { const double a = -2.0 * M_PI * m / N; tw[m] = make_float2((float)cos(a), (float)sin(a)); }
  for (int n = 0; n < N; ++n) x[n] = make_float2((float)(drand48() - 0.5), (float)(drand48() - 0.5));
  for (int l = 0; l < N2; ++l) {            // phase 1 "threads"
    float2 v[N1];
    for (int j = 0; j < N1; ++j) v[j] = x[N2 * j + l];
    phase1<INV, N1, N2>(v, l, tw.data());
    for (int k1 = 0; k1 < N1; ++k1) z[Exchange<N1, N2>::at(k1, l)] = v[k1];
  }
  for (int k1 = 0; k1 < N1; ++k1) {         // phase 2 "threads"
    float2 v[N2];
    for (int l = 0; l < N2; ++l) v[l] = z[Exchange<N1, N2>::at(k1, l)];
    phase2<INV, N1, N2>(v);
    for (int k2 = 0; k2 < N2; ++k2) X[k1 + N1 * k2] = v[k2];
  }
  for (int k = 0; k < N; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      const double a = (INV ? 2.0 : -2.0) * M_PI * (double)((long long)n * k % N) / N;
      sr += x[n].x * cos(a) - x[n].y * sin(a);
      si += x[n].x * sin(a) + x[n].y * cos(a);
    }
    maxerr = fmax(maxerr, fmax(fabs(sr - X[k].x), fabs(si - X[k].y)));
  }
}

int main() {
  srand48(7);
  check_small<false, 16>(); check_small<true, 16>(); check_small<false, 20>(); check_small<true, 20>();
  const double e_small = maxerr;
  maxerr = 0.0;
  for (int rep = 0; rep < 3; ++rep) { check_line<false, 16, 20>(); check_line<true, 16, 20>(); check_line<false, 20, 16>(); check_line<true, 20, 16>(); }
  printf("small %.3e line %.3e\n", e_small, maxerr);
  return (e_small < 5e-6 && maxerr < 5e-5) ? 0 : 1;
}
