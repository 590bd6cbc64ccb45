// Multi-tensor AdamW step: the reference steps up to five torch.optim.AdamW instances
// (model.py:72-81, lr 1e-4, betas (0.9, 0.999), eps 1e-8, weight_decay 0) once per iteration.  Here all
// parameter tensors of an optimiser are updated by ONE launch per <= 48 tensors: a table of (p, g, m, v, n)
// travels in the kernel parameters, every CTA owns one 16 K-element chunk of one tensor.
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)            (torch/optim/adamw.py, no amsgrad)
// HBM-bound: 28 B per parameter (read p, g, m, v; write p, m, v).
#include <cmath>

#include "san_common.cuh"
#include "../../include/san_b200.h"

namespace {

constexpr int AD_MAXT = 48;
constexpr int AD_CHUNK = 16384;

struct AdamTable {
  float* p[AD_MAXT];
  const float* g[AD_MAXT];
  float* m[AD_MAXT];
  float* v[AD_MAXT];
  long long n[AD_MAXT];
  int chunk0[AD_MAXT + 1];   // first CTA of each tensor
  int nt;
};

// decay = 1 - lr*wd, omb1 = 1 - beta1, omb2 = 1 - beta2, step = lr / (1 - beta1^t): formed in fp64 on the host and
// rounded once, as torch does with its python-float hyper-parameters (1.f - 0.999f would be off by 1.3e-5)
// sched (optional, device): {lr / (1 - beta1^t), sqrt(1 - beta2^t)} written by adamw_sched_kernel from the DEVICE-side
// step counter - the CUDA-graph-capturable form (a host-side step number would be frozen into the graph)
__global__ void __launch_bounds__(256) adamw_kernel(const AdamTable t, float decay, float omb1, float b2, float omb2,
                                                    float eps, float step, float bc2_sqrt, const float* __restrict__ sched) {
  if (sched) { step = sched[0]; bc2_sqrt = sched[1]; }
  int ti = 0;
  while (ti + 1 < t.nt && (int)blockIdx.x >= t.chunk0[ti + 1]) ++ti;
  const long long base = (long long)((int)blockIdx.x - t.chunk0[ti]) * AD_CHUNK;
  long long end = base + AD_CHUNK;
  if (end > t.n[ti]) end = t.n[ti];
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * decay;
    const float mi = m[i] + (gi - m[i]) * omb1;                // lerp, as torch
    const float vi = v[i] * b2 + omb2 * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// t += 1 on the device; the two bias-correction scalars in fp64 exactly like the host path
__global__ void adamw_sched_kernel(int* step, float* sched, double lr, double beta1, double beta2) {
  const int t = *step + 1;
  *step = t;
  sched[0] = (float)(lr / (1.0 - pow(beta1, (double)t)));
  sched[1] = (float)sqrt(1.0 - pow(beta2, (double)t));
}

}  // namespace

extern "C" {

int san_adamw_step(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                   const long long* numel, int ntensors, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int step, int* step_dev, float* sched_dev, void* stream) {
  SAN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel && ntensors >= 0 && (step >= 1 || (step_dev && sched_dev)),
                "san_adamw_step: bad args");
  SAN_CHECK_ARG(!step_dev == !sched_dev, "san_adamw_step: step_dev and sched_dev go together");
  const double bc1 = step_dev ? 1.0 : 1.0 - pow(beta1, (double)step);
  const float bc2_sqrt = step_dev ? 1.f : (float)sqrt(1.0 - pow(beta2, (double)step));
  if (step_dev) {      // device-side step counter (CUDA-graph capture): one increment per call
    adamw_sched_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, sched_dev, lr, beta1, beta2);
    SAN_LAUNCH_CHECK();
  }
  for (int t0 = 0; t0 < ntensors; t0 += AD_MAXT) {
    AdamTable tab;
    tab.nt = ntensors - t0 < AD_MAXT ? ntensors - t0 : AD_MAXT;
    int blocks = 0;
    for (int i = 0; i < tab.nt; ++i) {
      SAN_CHECK_ARG(params[t0 + i] && grads[t0 + i] && exp_avg[t0 + i] && exp_avg_sq[t0 + i] && numel[t0 + i] > 0,
                    "san_adamw_step: null tensor %d", t0 + i);
      tab.p[i] = (float*)params[t0 + i];
      tab.g[i] = (const float*)grads[t0 + i];
      tab.m[i] = (float*)exp_avg[t0 + i];
      tab.v[i] = (float*)exp_avg_sq[t0 + i];
      tab.n[i] = numel[t0 + i];
      tab.chunk0[i] = blocks;
      blocks += (int)((numel[t0 + i] + AD_CHUNK - 1) / AD_CHUNK);
    }
    tab.chunk0[tab.nt] = blocks;
    adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tab, (float)(1.0 - lr * weight_decay), (float)(1.0 - beta1),
                                                           (float)beta2, (float)(1.0 - beta2), (float)eps,
                                                           (float)(lr / bc1), bc2_sqrt, sched_dev);
    SAN_LAUNCH_CHECK();
  }
  return SAN_OK;
}

}  // extern "C"
