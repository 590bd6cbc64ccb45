"""Micro-benchmark (GPU box): tcgen05 conv path vs the fp32 direct-conv kernel, per U-Net layer shape at bs=64."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spatialalignmentnetwork_b200 import _lib, ops  # noqa: E402

L = _lib


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    shapes = [(3, 18, 320, 3), (18, 18, 320, 3), (36, 18, 320, 3), (18, 36, 160, 3), (36, 36, 160, 3), (72, 36, 160, 3),
              (36, 72, 80, 3), (72, 72, 80, 3), (144, 72, 80, 3), (72, 144, 40, 3), (144, 144, 40, 3), (288, 144, 40, 3),
              (144, 288, 20, 3), (288, 288, 20, 3), (288, 576, 20, 1), (144, 288, 40, 1), (72, 144, 80, 1), (36, 72, 160, 1),
              (18, 2, 320, 1), (2, 32, 320, 3), (32, 32, 320, 3), (64, 64, 160, 3), (128, 64, 160, 3)]
    if len(sys.argv) > 2:      # e.g. "18,18,320,3;36,72,160,1"
        shapes = [tuple(int(v) for v in t.split(",")) for t in sys.argv[2].split(";")]
    print(f"N={N}: Cin Cout HW K | stage ms | tc conv ms (TF/s algorithmic) | fp32 conv ms (TF/s) | speedup(conv only)")
    for Cin, Cout, HW, K in shapes:
        x = torch.randn(N, Cin, HW, HW, device="cuda")
        w = torch.randn(Cout, Cin, K, K, device="cuda") / math.sqrt(Cin * K * K)
        xs = torch.empty(L.lib().san_tc_staged_act_elems(N, HW, HW, Cin), dtype=torch.bfloat16, device="cuda")
        ws = torch.empty(L.lib().san_tc_staged_weight_elems(HW, HW, Cout, Cin, K), dtype=torch.bfloat16, device="cuda")
        y = torch.empty(N, Cout, HW, HW, device="cuda")
        Cpad = (Cin + 7) // 8 * 8
        st = lambda: L.call("tc_stage_act", xs, N, HW, HW, Cpad, x, None, None, None, 1.0, Cin, 0,
                            None, None, None, None, 1.0, 0, 0, None, None, None, None, 1.0, 0, 0, 1)
        L.call("tc_stage_weights", w, ws, HW, HW, Cout, Cin, K, 0, 1)
        cv = lambda: L.call("tc_conv", xs, ws, None, y, N, HW, HW, Cin, Cout, K, 0, 3, None)  # fp16 pairs, as the forward runs
        t_cs = float("nan")
        if L.lib().san_tc_conv_stats_supported(HW, HW, Cin, Cout, K):      # the same conv with the statistics epilogue
            sums = torch.empty(2 * N * Cout, dtype=torch.float64, device="cuda")
            t_cs = timeit(lambda: L.call("tc_conv_stats", xs, ws, None, y, N, HW, HW, Cin, Cout, K, 0, 3, None, sums))
        t_rr = t_rr0 = float("nan")
        if K == 3 and L.lib().san_tc_conv_rows_supported(HW, HW, Cin, Cout, 3):     # row-ring conv: stages its raw input itself
            wsr = torch.empty(L.lib().san_tc_rows_weight_elems(HW, HW, Cout, Cin), dtype=torch.bfloat16, device="cuda")
            L.call("tc_stage_weights_rows", w, wsr, HW, HW, Cout, Cin, 0, 1)
            mu_ = torch.zeros(N * Cin, device="cuda"); a_ = torch.ones(N * Cin, device="cuda")
            sums2 = torch.empty(2 * N * Cout, dtype=torch.float64, device="cuda")
            xs2 = torch.empty_like(xs)
            t_rr = timeit(lambda: L.call("tc_conv_rows", x, mu_, a_, None, 0.2, None, xs2, wsr, None, y, sums2, N, HW, HW, Cin, Cout))
            t_rr0 = timeit(lambda: L.call("tc_conv_rows", x, mu_, a_, None, 0.2, None, None, wsr, None, y, None, N, HW, HW, Cin, Cout))
        wp = ops._pack(w, False)
        y2 = torch.empty_like(y)
        fp = lambda: L.call("conv2d_fwd", x, wp, None, y2, N, Cin, HW, HW, Cout, K, 0, 0)
        gy = torch.randn(N, Cout, HW, HW, device="cuda")
        gys = torch.empty(L.lib().san_tc_staged_act_elems(N, HW, HW, Cout), dtype=torch.bfloat16, device="cuda")
        L.call("tc_stage_act", gys, N, HW, HW, (Cout + 7) // 8 * 8, gy, None, None, None, 1.0, Cout, 0,
               None, None, None, None, 1.0, 0, 0, None, None, None, None, 1.0, 0, 0, 1)
        dw = torch.empty_like(w)
        wg = lambda: L.call("tc_wgrad", gys, xs, dw, None, None, N, HW, HW, Cin, Cout, K, 3, None)   # fp16 pairs (timing only)
        t_wg = timeit(wg)
        t_st, t_cv, t_fp = timeit(st), timeit(cv), timeit(fp)
        fl = 2.0 * N * Cout * HW * HW * Cin * K * K
        err = ((y - y2).norm() / y2.norm()).item()
        print(f"{Cin:4d} {Cout:4d} {HW:4d} {K} | {t_st:7.3f} | {t_cv:7.3f} ({fl / t_cv / 1e9:6.1f}) | {t_fp:7.3f} ({fl / t_fp / 1e9:6.1f}) | "
              f"{t_fp / t_cv:5.2f}x  err {err:.1e} | wgrad {t_wg:7.3f} ms ({fl / t_wg / 1e9:6.1f} TF/s) | conv+stats {t_cs:7.3f} ms | rows(stage+conv+stats+store) {t_rr:7.3f} ms, no store/stats {t_rr0:7.3f} ms")
        del x, xs, y, y2, gy, gys


if __name__ == "__main__":
    main()
