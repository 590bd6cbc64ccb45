"""Analytic cycle model of conv_tc_kernel (no GPU needed): per U-Net layer shape, the strip geometry the kernel
picks (san_tc_describe) -> tcgen05.mma count, shared-memory operand-read cycles, tensor-pipe cycles, HBM bytes ->
a lower bound on the launch time, next to the measured time of profiles/r1h_conv_tc_microbench.txt.

    python tools/conv_model.py            # table for the benchmark layers at bs 64

Per MMA (M = 128, N = Npad, K = 16, bf16): the A tile is 4 KB and the B tile Npad*32 B of shared memory, read at
128 B/clk/SM; the tensor pipe needs Npad/2 clk (4096 MAC/clk/SM dense).  For Npad <= 48 the operand read
(32 + Npad/4 clk) exceeds the tensor time: those layers are bound by shared-memory bandwidth, which is what ncu
shows (profiles/r1d_tc_ncu_summary.txt: shared pipe 88 % busy, tensor pipe 36 % active on 18->18 @320)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spatialalignmentnetwork_b200 import _lib  # noqa: E402

SMS, CLK_GHZ, HBM_GBS = 148, 1.93, 6552.6
MEASURED_MS = {  # profiles/r1h_conv_tc_microbench.txt (bs 64, CUDA events, kernel alone)
    (3, 18, 320, 3): 0.292, (18, 18, 320, 3): 0.553, (36, 18, 320, 3): 0.813, (18, 36, 160, 3): 0.152,
    (36, 36, 160, 3): 0.219, (72, 36, 160, 3): 0.351, (36, 72, 80, 3): 0.105, (72, 72, 80, 3): 0.144,
    (144, 72, 80, 3): 0.221, (72, 144, 40, 3): 0.066, (144, 144, 40, 3): 0.098, (288, 144, 40, 3): 0.181,
    (144, 288, 20, 3): 0.060, (288, 288, 20, 3): 0.106, (288, 576, 20, 1): 0.069, (144, 288, 40, 1): 0.107,
    (72, 144, 80, 1): 0.128, (36, 72, 160, 1): 0.181, (18, 2, 320, 1): 0.138, (2, 32, 320, 3): 0.385,
    (32, 32, 320, 3): 0.556, (64, 64, 160, 3): 0.329, (128, 64, 160, 3): 0.619,
}


MEASURED_FORM = {}      # shape -> 1 where MEASURED_MS was taken with the DXN form (round-2 numbers), else the 9-tap form


def describe(L, H, W, Cin, Cout, K):
    out = (ctypes.c_int * 16)()
    assert L.san_tc_describe(H, W, Cin, Cout, K, ctypes.addressof(out)) == 0
    keys = "Cin_pad KG KS nsplit Npad Wp Hp R T S_alloc strips stages acc_stages a_bytes b_bytes smem_bytes".split()
    g = dict(zip(keys, list(out)))
    form = (ctypes.c_int * 6)()
    assert L.san_tc_describe_form(H, W, Cin, Cout, K, ctypes.addressof(form)) == 0
    g.update(dict(zip("dxn Np wtaps xchg_bytes hls Ncol".split(), list(form))))
    return g


def model(L, N, Cin, Cout, HW, K):
    g = describe(L, HW, HW, Cin, Cout, K)
    taps = g["wtaps"]                               # 9, or 3 in the DXN form (horizontal taps in the MMA N dimension)
    mma_per_unit = g["T"] * taps * 3 * g["KS"]
    smem_clk = 32 + g["Npad"] / 4.0                 # A 4 KB + B Npad*32 B at 128 B/clk
    tens_clk = g["Npad"] / 2.0
    if g["hls"]:     # 2 MMAs per tap (N = 2 Npad and N = Npad) instead of 3 of N = Npad: express as 3 MMA-equivalents
        smem_clk = ((32 + g["Npad"] / 2.0) + (32 + g["Npad"] / 4.0)) / 3.0
        tens_clk = (g["Npad"] + g["Npad"] / 2.0) / 3.0
    units = N * g["strips"] * g["nsplit"]
    waves = -(-units // SMS)
    mma_ms = waves * mma_per_unit * max(smem_clk, tens_clk) / (CLK_GHZ * 1e6)
    tens_ms = waves * mma_per_unit * tens_clk / (CLK_GHZ * 1e6)
    # HBM: staged operand (hi + lo, padded channels, halo rows re-read per strip) + fp32 result
    rows_in = g["R"] + 2 * (K // 2) if K == 3 else g["R"] + 2      # the kernel always loads R + 2 padded rows
    in_bytes = N * g["strips"] * g["nsplit"] * rows_in * g["Wp"] * g["Cin_pad"] * 4
    out_bytes = N * Cout * HW * HW * 4
    hbm_ms = (in_bytes + out_bytes) / (HBM_GBS * 1e6)
    # epilogue: every accumulator column of every tile goes TMEM -> registers -> global; overlapped only when the
    # accumulators are double-buffered
    epi_clk = g["T"] * g["Npad"] / 8.0 * 40.0 / 4.0      # ~40 clk per 8-column tcgen05.ld + stores, 4 warps per quarter
    if g["dxn"]:
        epi_clk = max(epi_clk, g["T"] * g["Npad"] * 8.0)  # TMEM read of the 3 column blocks at 64 B/clk
    if g["hls"]:
        epi_clk *= 2.0                                    # two column blocks per pixel
    epi_ms = waves * epi_clk / (CLK_GHZ * 1e6)
    bound = max(mma_ms, hbm_ms) + (0.0 if g["acc_stages"] == 2 else epi_ms)
    return g, dict(mma=mma_ms, tensor=tens_ms, hbm=hbm_ms, epi=epi_ms, bound=bound, waves=waves,
                   limiter="smem" if smem_clk > tens_clk else "tensor")


def main():
    L = _lib.lib()
    N = 64
    print(f"bs {N}; clk {CLK_GHZ} GHz; 'mma' = issue-bound time of the MMA phase, 'tensor' = tensor-pipe time alone")
    print(" Cin Cout  HW K | Npad nsplit  R  T acc | waves |  mma ms  tensor ms  hbm ms | model ms  measured ms  ratio | limiter")
    tot_m = tot_b = 0.0
    for (Cin, Cout, HW, K), meas in MEASURED_MS.items():
        g, m = model(L, N, Cin, Cout, HW, K)
        tot_m += meas; tot_b += m["bound"]
        print(f"{Cin:4d} {Cout:4d} {HW:4d} {K} | {g['Npad']:4d} {g['nsplit']:6d} {g['R']:2d} {g['T']:2d} {g['acc_stages']:3d} | {m['waves']:5d} |"
              f" {m['mma']:7.3f} {m['tensor']:9.3f} {m['hbm']:7.3f} | {m['bound']:8.3f} {meas:11.3f} {meas / m['bound']:6.2f} | {m['limiter']}")
    print(f"sum: model {tot_b:.3f} ms, measured {tot_m:.3f} ms")


if __name__ == "__main__":
    main()
