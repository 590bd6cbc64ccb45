"""Micro-benchmark (GPU box): InstanceNorm + LeakyReLU backward, one fused kernel (san_in_bwd_fused_map) vs the three-kernel
path (act_bwd_reduce_map + in_finalize_bwd + act_bwd_apply_map), per cascade U-Net tensor shape at bs 64.
Algorithmic bytes: y and dx read once, dy written once = 12 B per element."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spatialalignmentnetwork_b200 import _lib as L  # noqa: E402


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
    print(f"N={N}: C HW | fused ms (GB/s algorithmic) | 3-kernel ms (GB/s)")
    for C, HW in [(18, 320), (36, 160), (72, 80), (144, 40), (288, 20)]:
        y = torch.randn(N, C, HW, HW, device="cuda")
        dx = torch.randn(N, C, HW, HW, device="cuda") * 1e-5
        mu = y.mean((2, 3)).reshape(-1).contiguous()
        a = (1 / torch.sqrt(y.var((2, 3), unbiased=False) + 1e-5)).reshape(-1).contiguous()
        dy, am = torch.empty_like(y), torch.empty(1, device="cuda")
        wk = torch.empty(5, N * C, device="cuda")
        fused = lambda: L.call("in_bwd_fused_map", dx, C, 0, 0, y, mu, a, 0.2, dy, N, C, HW, HW, am)

        def three():
            L.call("act_bwd_reduce_map", dx, C, 0, 0, y, mu, a, None, a, 0.2, wk[0], wk[1], N, C, HW, HW)
            L.call("in_finalize_bwd", wk[0], wk[1], a, wk[2], wk[3], wk[4], N * C, HW * HW)
            L.call("act_bwd_apply_map", dx, C, 0, 0, y, mu, a, None, 0.2, wk[2], wk[3], wk[4], dy, N, C, HW, HW, am)
        tf, t3 = timeit(fused), timeit(three)
        by = 12.0 * y.numel()
        print(f"{C:4d} {HW:4d} | {tf:7.3f} ({by / tf / 1e6:7.1f}) | {t3:7.3f} ({by / t3 / 1e6:7.1f})")


if __name__ == "__main__":
    main()
