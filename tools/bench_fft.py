"""Micro-benchmark (GPU box): fused FFT + data-consistency kernels at the benchmark size."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spatialalignmentnetwork_b200 import _lib  # noqa: E402

L = _lib


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    C, H, W = 1, 320, 320
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    k = torch.randn(N, C, H, W, dtype=torch.complex64, device="cuda")
    k0 = torch.randn_like(k); S = torch.randn_like(k); out = torch.empty_like(k); tmp = torch.empty_like(k)
    x = torch.randn(N, 2, H, W, device="cuda")
    mask = (torch.rand(W, device="cuda") > 0.75)
    dcw = torch.ones(1, device="cuda")
    P = N * H * W
    t = timeit(lambda: L.call("fft_expand_dc", x, S, k, k0, mask, dcw, out, tmp, N, C, H, W, 0), reps)
    print(f"fft_expand_dc  {t:.4f} ms  {(32 * C + 8) * P / t / 1e6:8.1f} GB/s algorithmic")
    xo = torch.empty(N, 2, H, W, device="cuda")
    t = timeit(lambda: L.call("fft_reduce", k, S, xo, None, tmp, N, C, H, W, 1, 1.0), reps)
    print(f"fft_reduce     {t:.4f} ms  {(16 * C + 8) * P / t / 1e6:8.1f} GB/s algorithmic")
    t = timeit(lambda: L.call("fft2", k.view(N * C, H, W), 0, None, out.view(N * C, H, W), 0, None, out, N * C, H, W, 0), reps)
    print(f"fft2           {t:.4f} ms  {16 * C * P / t / 1e6:8.1f} GB/s algorithmic")


if __name__ == "__main__":
    main()
