// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for the U-Net conv stacks, sm_100a only.
//
// Replaces the cuDNN convs of the reference's VarNet / alignment U-Nets (reference
// varnet.py:139-146 3x3 no-bias, :75-80 1x1 + bias, :176-179 ConvTranspose2d 2x2 s2 run as a 1x1
// conv + pixel shuffle; unet.py:119-140 3x3 / 1x1 + bias).
//
// Precision: fp32 parity needs more than one bf16 pass (single-pass BF16 is 2.9e-2
// off after 12 cascades, BF16x3 is 7e-5).  Operands are split x = hi + lo (two bf16) and the
// product is accumulated in fp32 TMEM as  hi*hi + lo*hi + hi*lo  (3 tcgen05.mma per K-step).
//
// Formulation ("flattened padded pixels"): activations are STAGED by san_tc_stage_act as
//     Xs[n][hl][kg][slot][8]  bf16,   slot = (h+1)*Wp + (w+1),  Wp = W+2, Hp = H+2, zero border,
//     hl = 0 (hi) / 1 (lo), kg = channel group of 8, ceil(Cin / 8) groups (the last one zero-padded; there is NO all-zero
//     group: with an odd group count the last K = 16 step has one real group, see tap pairing below).
// For one (hl, kg) a span of image rows is ONE contiguous byte range, so the producer warp moves it
// with a single TMA bulk copy (cp.async.bulk, UBLKCP), and in shared memory it is exactly the
// canonical no-swizzle K-major UMMA operand: 8 channels = 16 B per pixel row, uniform 16 B row
// pitch (SBO = 128 B), channel groups LBO apart.  Output pixel q = r*Wp + x of a strip of R rows
// needs, for tap (dy, dx), the operand row  q + dy*Wp + dx: a 3x3 tap is just a different START
// ADDRESS of the same staged tile -- no im2col copies, each input element is loaded once per
// strip.  Columns x >= W (2 per row) are computed and dropped (0.6 % waste at W = 320).
//
// Narrow layers ("DXN" form, 3x3 with <= 40 output channels: the full- and half-resolution layers that carry the
// step): an MMA with N = 32 reads its 4 KB A tile from shared memory in 32 clk but keeps the tensor pipe busy for
// 16, and the 9 taps x 3 hi/lo products re-read A 27 times per K-step.  There the three horizontal taps are moved
// into the MMA's N dimension: B = [W(dy,dx=0) | W(dy,dx=1) | W(dy,dx=2)] (N = 3*pad8(Cout), padded to 16), one A
// window per filter ROW, so A is read 9 times per K-step (and N = 80 balances operand read and tensor time).
// The accumulator then holds E_dx[q] = sum_dy A[q + dy*Wp] W[dy][dx], and y[q] = E_0[q] + E_1[q+1] + E_2[q+2]:
// the epilogue adds the three column blocks across neighbouring TMEM lanes (warp shuffles; the two lanes at a
// 32-lane quarter boundary are completed through a small shared-memory exchange).
//
// Work unit = (image n, strip of R rows, N-split); persistent CTAs walk units.  Warp roles:
//   warp 0      TMA producer  (bulk copies of the A row span + the B weight block per K-step)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..9  epilogue (two warps per TMEM lane quarter, alternating tiles): tcgen05.ld the fp32 accumulators, + bias, store NCHW fp32 (coalesced:
//               TMEM lane = pixel, so a warp writes 32 consecutive pixels of one channel)
// Pipelines: smem full/empty per K-step stage, TMEM accumulator full/empty (double-buffered when
// 2*T*Npad <= 512 columns) so the epilogue of unit i overlaps the MMAs of unit i+1.
#include "tc_common.cuh"
#include "../../include/san_b200.h"

namespace {

// ------------------------------------------------------------------------------------ geometry
constexpr int TC_EPI_WARPS = 16;      // epilogue warps: 2 per TMEM lane quarter, alternating tiles (a single warp per
                                      // quarter was instruction-latency bound: ~1450 dependent instructions per unit)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_SMEM_HEADER = 256;   // barriers + tmem pointer
constexpr int TC_SMEM_MAX = 225 * 1024;
constexpr int TC_NMAX = 160;          // largest UMMA N used per unit

struct TcGeom {
  int Cin_pad, KG, KS;     // input channels padded to 16; REAL groups of 8 = staged planes per (image, half); K-steps of 16
  int nsplit, Npad;        // output channels split into nsplit units of Npad (multiple of 16)
  int Wp, Hp, PS;          // padded width/height, pixel slots per image
  int R, T, S_alloc;       // rows per strip, 128-pixel M-tiles per strip, smem slots per (hl, kk)
  int strips, stages, acc_stages;
  int a_bytes, b_bytes, stage_bytes, smem_bytes;
  int dxn, Np;             // DXN form: horizontal taps in the MMA N dimension, Np = Cout padded to 8 (Npad = pad16(3*Np))
  int wtaps;               // weight blocks per K-step: 9 (3x3), 3 (DXN: one per filter row) or 1 (1x1)
  int hls, Ncol;           // HLS form: [W_hi | W_lo] stacked along the MMA N dimension; Ncol = TMEM columns per tile
  int pair;                // the last K-step has ONE real channel group: filter taps are paired inside its K = 16 MMAs
  int xchg_bytes;          // DXN: shared-memory exchange area of the epilogue (quarter-boundary lanes)
};

inline int pad8(int c) { return (c + 7) / 8 * 8; }
int tc_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// SAN_TC_DXN = 1 enables the DXN form (off by default: first measurement 1.34 vs 0.55 ms on 18->18 @320, the per-tile
// epilogue chain of one warp is the critical path - profiles/r2d_*); SAN_TC_DXN_R = r forces its strip height (tuning runs)
int tc_dxn_enabled() { static const int v = tc_env_int("SAN_TC_DXN", 0); return v; }
int tc_dxn_force_r() { static const int v = tc_env_int("SAN_TC_DXN_R", 0); return v; }
// SAN_TC_HLS = 0 disables the hi/lo-stacked form of narrow 3x3 layers (A/B runs); SAN_TC_HLS_R = r forces its strip height
int tc_hls_enabled() { static const int v = tc_env_int("SAN_TC_HLS", 1); return v; }
int tc_hls_force_r() { static const int v = tc_env_int("SAN_TC_HLS_R", 0); return v; }
// SAN_TC_PAIR = 0 disables tap pairing in the half-empty last K-step (A/B runs)
int tc_pair_enabled() { static const int v = tc_env_int("SAN_TC_PAIR", 1); return v; }

// HLS geometry for narrow 3x3 layers (<= 32 padded output channels): B = [W_hi | W_lo] stacked along N, so that
// A_hi is read from shared memory ONCE for the two products hi*hi and hi*lo (N = 2*Npad), plus one MMA A_lo x W_hi
// (N = Npad): 2 reads of the 4 KB A tile per tap instead of 3 (the operand read, not the tensor pipe, bounds these
// layers); the epilogue adds the two column blocks of a pixel.  Costs 2x the TMEM columns, so strips are one image
// row at W = 320 to keep the accumulators double-buffered; chosen by the same cycle estimate as above.
bool tc_geometry_hls(int H, int W, int Cin, int Cout, TcGeom* g) {
  const int Npad = pad16(Cout);
  if (Npad > 32) return false;       // measured: Npad = 48 (N = 96) loses (18->36 @160: 0.156 -> 0.19-0.22 ms), 16 / 32 win
  const int Ncol = 2 * Npad, Wp = W + 2, KS = pad16(Cin) / 16;
  const int b_bytes = 9 * 4 * Npad * 16;
  const int force = tc_hls_force_r();
  int bestR = 0;
  double best = 1e30;
  for (int R = 1; R <= H; ++R) {
    const int T = (R * Wp + 127) / 128;
    if (T * Ncol > 512) break;
    const int S = ((128 * T + 2 * Wp + 2) + 7) / 8 * 8;
    if (2 * (4 * S * 16 + b_bytes) + TC_SMEM_HEADER > TC_SMEM_MAX) break;
    if (force && R != force) continue;
    const double rd1 = (4096.0 + Ncol * 32.0) / 128.0, rd2 = (4096.0 + Npad * 32.0) / 128.0;
    const double mma = (double)T * KS * 9.0 * (fmax(rd1, Ncol / 2.0) + fmax(rd2, Npad / 2.0));
    const double epi = (double)T * Ncol * 8.0 * 2.0;          // TMEM read at 64 B/clk + the per-warp latency chain
    const bool dbl = 2 * T * Ncol <= 512;
    const int strips = (H + R - 1) / R;
    const double per_row = (dbl ? fmax(mma, epi) : mma + epi) * strips / (double)H * (1.0 + 0.1 / R);
    if (per_row < best - 1e-9) { best = per_row; bestR = R; }
  }
  if (!bestR) return false;
  g->hls = 1; g->nsplit = 1; g->Npad = Npad; g->Ncol = Ncol; g->b_bytes = b_bytes;
  g->R = bestR;
  g->T = (g->R * Wp + 127) / 128;
  return true;
}

// DXN geometry; false if the layer does not qualify.  The strip height minimises a cycle estimate per output row:
// MMA phase (9 MMAs per tile and K-step, each max(operand read at 128 B/clk, tensor N/2 clk)) against the epilogue's
// TMEM read (64 B/clk per SM), overlapped only when the accumulators are double-buffered.
bool tc_geometry_dxn(int H, int W, int Cin, int Cout, TcGeom* g) {
  const int Np = pad8(Cout);
  if (Np > 40) return false;
  const int Ndx = pad16(3 * Np), Wp = W + 2, KS = pad16(Cin) / 16;
  const int b_bytes = 3 * 4 * Ndx * 16;
  const int force = tc_dxn_force_r();
  int bestR = 0;
  double best = 1e30;
  for (int R = 1; R <= H; ++R) {
    const int T = (R * Wp + 127) / 128;
    if (T * Ndx > 512) break;
    const int S = ((128 * T + 2 * Wp + 2) + 7) / 8 * 8;
    const int xchg = (80 * T * Np + 127) / 128 * 128;
    if (2 * (4 * S * 16 + b_bytes) + TC_SMEM_HEADER + xchg > TC_SMEM_MAX) break;
    if (force && R != force) continue;
    const double mma = (double)T * KS * 9.0 * fmax((4096.0 + Ndx * 32.0) / 128.0, Ndx / 2.0);
    const double epi = (double)T * Ndx * 8.0;
    const bool dbl = 2 * T * Ndx <= 512;
    const int strips = (H + R - 1) / R;
    const double per_row = (dbl ? fmax(mma, epi) : mma + epi) * strips / (double)H * (1.0 + 0.1 / R);   // mild halo penalty
    if (per_row < best - 1e-9) { best = per_row; bestR = R; }
  }
  if (!bestR) return false;
  g->dxn = 1; g->Np = Np; g->wtaps = 3;
  g->nsplit = 1; g->Npad = Ndx; g->b_bytes = b_bytes;
  g->R = bestR;
  g->T = (g->R * Wp + 127) / 128;
  g->xchg_bytes = (80 * g->T * Np + 127) / 128 * 128;
  return true;
}


// Rows per strip for a given output-channel split; 0 if nothing fits (shared memory / TMEM columns).
static int tc_pick_rows(int H, int W, int K, int Npad, int b_bytes) {
  const int Wp = W + 2;
  int bestR = 0;
  double best = -1.0;
  for (int R = 1; R <= H; ++R) {
    const int T = (R * Wp + 127) / 128;
    if (T * Npad > 512) break;
    const int S = ((128 * T + 2 * Wp + 2) + 7) / 8 * 8;
    const int stage = 4 * S * 16 + b_bytes;
    if (2 * stage + TC_SMEM_HEADER > TC_SMEM_MAX) break;
    // useful MMA rows x re-read factor of the input rows (halo) x tail waste of the last strip
    const int strips = (H + R - 1) / R;
    // ... x a penalty when the accumulators cannot be double-buffered (epilogue then serialises with the MMAs)
    const double eff = ((double)R * W / (128.0 * T)) * ((double)R / (R + 2 * (K / 2))) * ((double)H / (strips * R)) *
                       (2 * T * Npad <= 512 ? 1.0 : 0.75);
    if (eff > best + 1e-9) { best = eff; bestR = R; }
  }
  return bestR;
}

bool tc_geometry(int H, int W, int Cin, int Cout, int K, TcGeom* g) {
  if (K != 1 && K != 3) return false;
  g->Cin_pad = pad16(Cin); g->KG = (Cin + 7) / 8; g->KS = g->Cin_pad / 16;
  g->Wp = W + 2; g->Hp = H + 2; g->PS = g->Hp * g->Wp;
  const int ntaps = K * K;
  g->dxn = 0; g->Np = 0; g->wtaps = ntaps; g->xchg_bytes = 0; g->hls = 0;
  if (K == 3 && tc_dxn_enabled() && tc_geometry_dxn(H, W, Cin, Cout, g)) {
  } else if (K == 3 && tc_hls_enabled() && tc_geometry_hls(H, W, Cin, Cout, g)) {
  } else {
    // split the output channels until a strip (A rows + the weight block of one K-step, two stages) fits
    int bestR = 0;
    for (g->nsplit = (pad16(Cout) + TC_NMAX - 1) / TC_NMAX; g->nsplit <= 16; ++g->nsplit) {
      g->Npad = pad16((Cout + g->nsplit - 1) / g->nsplit);
      g->b_bytes = ntaps * 4 * g->Npad * 16;
      bestR = tc_pick_rows(H, W, K, g->Npad, g->b_bytes);
      if (bestR || g->Npad == 16) break;
    }
    if (bestR == 0) return false;
    g->R = bestR;
    g->T = (g->R * g->Wp + 127) / 128;
  }
  if (!g->hls) g->Ncol = g->Npad;
  // Tap pairing: with an odd number of 8-channel groups (Cin = 18 -> 3 groups, 3 -> 1, 36 -> 5, 72 -> 9) the last K = 16
  // step would multiply one real group and one all-zero group for each of the 9 taps.  Instead the zero group's slot
  // of the MMA is pointed (descriptor leading-byte-offset) at the SAME real group shifted by the next tap's offset:
  // 5 steps (taps 0+1, 2+3, 4+5, 6+7, 8+none) instead of 9, and the zero group is never loaded.
  g->pair = (K == 3 && !g->dxn && (((Cin + 7) / 8) & 1) && tc_pair_enabled()) ? 1 : 0;
  g->S_alloc = ((128 * g->T + 2 * g->Wp + 2) + 7) / 8 * 8;
  g->a_bytes = 4 * g->S_alloc * 16;
  g->stage_bytes = g->a_bytes + g->b_bytes;
  g->stages = (TC_SMEM_MAX - TC_SMEM_HEADER - g->xchg_bytes) / g->stage_bytes;
  if (g->stages > 4) g->stages = 4;
  if (g->stages < 2) return false;
  g->strips = (H + g->R - 1) / g->R;
  g->acc_stages = (2 * g->T * g->Ncol <= 512) ? 2 : 1;
  g->smem_bytes = TC_SMEM_HEADER + g->xchg_bytes + g->stages * g->stage_bytes;
  if (g->smem_bytes < 120 * 1024) g->smem_bytes = 120 * 1024;  // one CTA per SM (each allocates all 512 TMEM columns)
  return true;
}

struct ConvTcParams {
  const __nv_bfloat16* xs;  // staged activations [N][2][KG][PS][8]
  const __nv_bfloat16* ws;  // staged weights [nsplit][KS][ntaps][2][2][Npad][8]
  const float* bias;        // [Cout] or null
  float* y;                 // [N][Cout][H][W] (batch stride y_bs)
  long long y_bs;
  int N, H, W, Cout, ntaps;
  int nunits;
  int fmt;                  // TC_FMT_* bits: which operands are fp16 pairs (else bf16 pairs)
  float out_scale;          // exact inverse of the operands' static scales
  const float* a_absmax;    // null, or device scalar max|A| when A was staged with the dynamic scale (dY in the data gradient)
  double* sums;             // null, or [N][Cout][2] fp64, pre-zeroed: per-plane sum / sum of squares of y, accumulated by the epilogue
  int contig;               // 1: a CTA walks a CONTIGUOUS range of units (consecutive strips of one image) instead of a strided one
  int stat_off;             // byte offset of the epilogue's per-warp statistics accumulators in shared memory
  TcGeom g;
};

// Units of this CTA: [u0, u1) with step us.  Strided (u = blockIdx.x + k * gridDim.x): neighbouring CTAs work on neighbouring
// strips at the same time; contiguous: one CTA walks consecutive strips of the same image (the statistics epilogue then
// flushes its accumulators only a few times per launch).
__device__ __forceinline__ void tc_unit_range(const ConvTcParams& p, int& u0, int& u1, int& us) {
  if (p.contig) {
    u0 = (int)((long long)blockIdx.x * p.nunits / gridDim.x);
    u1 = (int)((long long)(blockIdx.x + 1) * p.nunits / gridDim.x);
    us = 1;
  } else {
    u0 = blockIdx.x; u1 = p.nunits; us = gridDim.x;
  }
}

// Sum over the 32 lanes of 16 per-lane values w[0..15] (a transposing butterfly: 16 shuffles instead of 80): on return
// every lane holds the total of value index (lane >> 1) & 15 (both lanes of a pair hold the same one).
__device__ __forceinline__ float warp_sum16(const float (&w)[16], int lane) {
  float r[8], q[4], t[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? w[i] : w[8 + i], keep = b4 ? w[8 + i] : w[i];
    r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? r[i] : r[4 + i], keep = b3 ? r[4 + i] : r[i];
    q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? q[i] : q[2 + i], keep = b2 ? q[2 + i] : q[i];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? t[0] : t[1], keep = b1 ? t[1] : t[0];
  float u = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  u += __shfl_xor_sync(0xffffffffu, u, 1);
  return u;
}

// ---------------------------------------------------------------------------------- MMA issue
// One K-step of one unit: NSTEP (A window, B block) descriptor pairs per 128-pixel tile.  a_word[i] = low descriptor word
// of step i without the stage address: window offset (16 B units) | leading-byte-offset << 16.
template <int NSTEP, bool HLS>
__device__ __forceinline__ void issue_kstep(const uint32_t (&a_word)[NSTEP], uint32_t a_s, uint32_t a_hi, uint32_t a_losplit,
                                            uint32_t b_base, uint32_t b_hi, uint32_t b_losplit, uint32_t b_tapstride,
                                            uint32_t d0, int T, int Ncol, uint32_t idesc, uint32_t idesc2, uint32_t first) {
  uint32_t d = d0, a_t = a_s;
  for (int t = 0; t < T; ++t, d += Ncol, a_t += 128) {
#pragma unroll
    for (int i = 0; i < NSTEP; ++i) {
      const uint32_t al = a_t + a_word[i];
      const uint32_t bl = b_base + i * b_tapstride;
      const uint64_t A_hi = ((uint64_t)a_hi << 32) | al, A_lo = ((uint64_t)a_hi << 32) | (al + a_losplit);
      if (HLS) {
        // B block = [kk][2*Npad rows: W_hi then W_lo][8]: A_hi x [W_hi | W_lo] (N = 2*Npad) + A_lo x W_hi (N = Npad)
        const uint64_t B_all = ((uint64_t)b_hi << 32) | bl;
        tc_mma_bf16(d, A_hi, B_all, idesc2, i == 0 ? first : 1u);   // hi*hi -> columns [0, Npad), hi*lo -> [Npad, 2 Npad)
        tc_mma_bf16(d, A_lo, B_all, idesc, 1u);                     // lo*hi -> columns [0, Npad)
      } else {
        const uint64_t B_hi = ((uint64_t)b_hi << 32) | bl, B_lo = ((uint64_t)b_hi << 32) | (bl + b_losplit);
        tc_mma_bf16(d, A_hi, B_hi, idesc, i == 0 ? first : 1u);     // hi*hi
        tc_mma_bf16(d, A_lo, B_hi, idesc, 1u);                      // lo*hi
        tc_mma_bf16(d, A_hi, B_lo, idesc, 1u);                      // hi*lo
      }
    }
  }
}

template <int NTAPS>
__device__ __forceinline__ void mma_issue_loop(const ConvTcParams& p, const TcGeom& g, uint32_t tmem_base,
                                               uint32_t stage0, uint32_t bar_full, uint32_t bar_empty,
                                               uint32_t bar_accf, uint32_t bar_acce) {
  const uint32_t idesc = umma_idesc_16(128, g.Npad, p.fmt);
  const uint32_t idesc2 = umma_idesc_16(128, 2 * g.Npad, p.fmt);      // HLS form: N = [W_hi | W_lo]
  const bool hls = g.hls != 0;
  // descriptor templates (address field = 0) and per-tap offsets in 16 B units
  const uint64_t a_tmpl = umma_desc(0, (uint32_t)g.S_alloc * 16, 128);
  const uint64_t b_tmpl = umma_desc(0, (uint32_t)(hls ? 2 * g.Npad : g.Npad) * 16, 128);   // LBO = one kk plane of B rows
  const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), b_hi = (uint32_t)(b_tmpl >> 32);
  const uint32_t a_lo0 = (uint32_t)a_tmpl, b_lo0 = (uint32_t)b_tmpl;
  uint32_t tap_off[NTAPS], a_word[NTAPS];
#pragma unroll
  for (int tap = 0; tap < NTAPS; ++tap) {   // 9: (dy, dx) windows; 3: DXN form, one window per filter row; 1: 1x1 centre
    tap_off[tap] = (NTAPS == 9) ? (uint32_t)((tap / 3) * g.Wp + (tap % 3)) : (NTAPS == 3) ? (uint32_t)(tap * g.Wp) : (uint32_t)(g.Wp + 1);
    a_word[tap] = a_lo0 + tap_off[tap];     // K group 1 = the next channel-group plane (LBO = S_alloc slots)
  }
  // paired last K-step (g.pair): K group 0 = the real channel group at tap 2i, K group 1 = the SAME plane at tap 2i+1
  // (LBO = the two taps' window distance); the 9th tap pairs with itself (LBO 0) against zero weights
  uint32_t p_word[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const uint32_t o0 = (NTAPS == 9) ? tap_off[(2 * i) % NTAPS] : 0u;
    const uint32_t o1 = (NTAPS == 9 && i < 4) ? tap_off[(2 * i + 1) % NTAPS] : o0;
    p_word[i] = ((o1 - o0) << 16) + o0;
  }
  const uint32_t a_losplit = 2u * g.S_alloc;     // hi -> lo half of A, 16 B units
  const uint32_t b_losplit = 2u * g.Npad;        // hi -> lo half of B
  const uint32_t b_tapstride = 4u * g.Npad;
  int s = 0, as = 0;
  uint32_t ph = 0, aph = 0;
  int u0, u1, us;
  tc_unit_range(p, u0, u1, us);
  for (int u = u0; u < u1; u += us) {
    mbar_wait(bar_acce + 8 * as, aph ^ 1);
    tc_fence_after();
    const uint32_t acc0 = tmem_base + (uint32_t)(as * g.T * g.Ncol);
    for (int ks = 0; ks < g.KS; ++ks) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      const uint32_t a_s = (stage0 + (uint32_t)s * g.stage_bytes) >> 4;   // 16 B units
      const uint32_t b_s = a_s + ((uint32_t)g.a_bytes >> 4);
      const uint32_t first = (ks != 0);
      const bool paired = NTAPS == 9 && g.pair && ks == g.KS - 1;
      if (elect_one_sync()) {
        if (paired) {
          if (hls) issue_kstep<5, true>(p_word, a_s, a_hi, a_losplit, b_lo0 + b_s, b_hi, b_losplit, b_tapstride, acc0, g.T, g.Ncol, idesc, idesc2, first);
          else issue_kstep<5, false>(p_word, a_s, a_hi, a_losplit, b_lo0 + b_s, b_hi, b_losplit, b_tapstride, acc0, g.T, g.Ncol, idesc, idesc2, first);
        } else {
          if (hls) issue_kstep<NTAPS, true>(a_word, a_s, a_hi, a_losplit, b_lo0 + b_s, b_hi, b_losplit, b_tapstride, acc0, g.T, g.Ncol, idesc, idesc2, first);
          else issue_kstep<NTAPS, false>(a_word, a_s, a_hi, a_losplit, b_lo0 + b_s, b_hi, b_losplit, b_tapstride, acc0, g.T, g.Ncol, idesc, idesc2, first);
        }
        tc_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
        if (ks == g.KS - 1) tc_commit(bar_accf + 8 * as);   // accumulators of this unit complete
      }
      __syncwarp();
      if (++s == g.stages) { s = 0; ph ^= 1; }
    }
    if (++as == g.acc_stages) { as = 0; aph ^= 1; }
  }
}

// add this warp's accumulators of segment seg = n * nsplit + ns to sums[n][c_base + c][q] and clear them
__device__ __forceinline__ void tc_flush_stats(const ConvTcParams& p, const TcGeom& g, float* sacc, int seg, int lane) {
  __syncwarp();
  const int n = seg / g.nsplit, ns = seg - n * g.nsplit;
  const int c_base = ns * g.Npad;
  const int c_cnt = min(g.Npad, p.Cout - c_base);
  // fp64 atomics: a warp's partial is a deterministic fp32 sum (fixed unit order); the few partials per plane are added in
  // fp64, where the order of the atomics no longer shows after the final rounding - the statistics are reproducible
  double* dst = p.sums + ((long long)n * p.Cout + c_base) * 2;
  for (int i = lane; i < 2 * c_cnt; i += 32) {
    const float v = sacc[i];
    if (v != 0.f) atomicAdd(dst + i, (double)v);
    sacc[i] = 0.f;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------- main kernel
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvTcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const TcGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // header: full[4] empty[4] acc_full[2] acc_empty[2] (8 B each) | tmem base (4 B)
  const uint32_t hdr = smem_u32(smem);
  const uint32_t bar_full = hdr, bar_empty = hdr + 32, bar_accf = hdr + 64, bar_acce = hdr + 80;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + 96);
  const uint32_t stage0 = hdr + TC_SMEM_HEADER + (uint32_t)g.xchg_bytes;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((void*)tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if ((g.KG & 1) && !g.pair) {
    // A planes that no bulk copy ever fills (see the producer) must hold finite values: zero the stage area once
    uint4* z = (uint4*)(smem + TC_SMEM_HEADER + g.xchg_bytes);
    const int n16 = g.stages * g.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores before async-proxy (TMA / UMMA) accesses
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform by construction

  const int units_per_image = g.strips * g.nsplit;
  const long long plane_elems = (long long)g.PS * 8;  // elements of one (n, hl, kg) plane

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int u0, u1, us;
      tc_unit_range(p, u0, u1, us);
      for (int u = u0; u < u1; u += us) {
        const int n = u / units_per_image;
        const int rem = u - n * units_per_image;
        const int ns = rem / g.strips, st = rem - ns * g.strips;
        const int y0 = st * g.R;
        const int rows_in = min(g.R + 2, g.Hp - y0);
        const uint32_t bytesA = (uint32_t)rows_in * g.Wp * 16;
        for (int ks = 0; ks < g.KS; ++ks) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t sbase = stage0 + (uint32_t)s * g.stage_bytes;
          // an odd number of 8-channel groups: the staged tensor has no plane for the padding group of the last
          // K-step.  With tap pairing it is never read; otherwise the MMA reads that (zero-initialised, later stale
          // but finite) shared-memory plane against all-zero weight rows.
          const int nkk = (2 * ks + 1 < g.KG) ? 2 : 1;
          mbar_expect_tx(bar_full + 8 * s, 2 * nkk * bytesA + (uint32_t)g.b_bytes);
#pragma unroll
          for (int hl = 0; hl < 2; ++hl)
            for (int kk = 0; kk < nkk; ++kk) {
              const __nv_bfloat16* src =
                  p.xs + ((long long)(n * 2 + hl) * g.KG + 2 * ks + kk) * plane_elems + (long long)y0 * g.Wp * 8;
              bulk_g2s(sbase + (uint32_t)(hl * 2 + kk) * g.S_alloc * 16, src, bytesA, bar_full + 8 * s);
            }
          const __nv_bfloat16* wsrc = p.ws + ((long long)ns * g.KS + ks) * (g.b_bytes / 2);
          bulk_g2s(sbase + g.a_bytes, wsrc, (uint32_t)g.b_bytes, bar_full + 8 * s);
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // One thread issues every tcgen05.mma of the CTA, so the loop is written for issue rate: the
    // shared-memory descriptors are built once and advanced by adding 16-byte-unit offsets to
    // their low word; taps and the three hi/lo products are fully unrolled.
    // (whole warp: the loops are warp-uniform, one elected lane issues)
    if (g.wtaps == 9) mma_issue_loop<9>(p, g, tmem_base, stage0, bar_full, bar_empty, bar_accf, bar_acce);
    else if (g.wtaps == 3) mma_issue_loop<3>(p, g, tmem_base, stage0, bar_full, bar_empty, bar_accf, bar_acce);
    else mma_issue_loop<1>(p, g, tmem_base, stage0, bar_full, bar_empty, bar_accf, bar_acce);
  } else {
    // ================================ epilogue ====================================
    const int wq = warp & 3;            // TMEM lane quarter this warp may access
    const int egrp = (warp - 2) >> 2;   // which of the warps sharing that quarter: takes tiles t = egrp, egrp + G, ...
    const float oscale = p.a_absmax ? p.out_scale / tc_dyn_scale(__ldg(p.a_absmax)) : p.out_scale;
    constexpr int EG = TC_EPI_WARPS / 4;
    int as = 0;
    uint32_t aph = 0;
    const long long HW = (long long)p.H * p.W;
    // statistics epilogue (p.sums): per-warp accumulators [Npad][2] in shared memory (sum, sum of squares of the values this
    // warp stored for the current (image, channel split) segment), flushed with atomicAdd when the segment changes
    float* sacc = (float*)(smem + p.stat_off) + (size_t)(warp - 2) * (2 * g.Npad);
    int seg = -1;
    if (p.sums)
      for (int i = lane; i < 2 * g.Npad; i += 32) sacc[i] = 0.f;
    int u0, u1, us;
    tc_unit_range(p, u0, u1, us);
    for (int u = u0; u < u1; u += us) {
      const int n = u / units_per_image;
      const int rem = u - n * units_per_image;
      const int ns = rem / g.strips, st = rem - ns * g.strips;
      const int y0 = st * g.R;
      const int c_base = ns * g.Npad;
      const int c_cnt = min(g.Npad, p.Cout - c_base);
      if (p.sums && seg != u / g.strips) {
        if (seg >= 0) tc_flush_stats(p, g, sacc, seg, lane);
        seg = u / g.strips;
      }
      mbar_wait(bar_accf + 8 * as, aph);
      tc_fence_after();
      float* yn = p.y + (long long)n * p.y_bs;
      if (g.dxn) {
        // ---- DXN form: columns = (dx, co); y[q] = E_0[q] + E_1[q+1] + E_2[q+2] across TMEM lanes
        const int Np = g.Np;
        float* xch = (float*)(smem + TC_SMEM_HEADER);      // [T][4 quarters][5][Np]
        for (int t = egrp; t < g.T; t += EG) {
          const int q = t * 128 + wq * 32 + lane;
          const int r = q / g.Wp, x = q - r * g.Wp;
          const int yy = y0 + r;
          const bool valid = (r < g.R) && (x < p.W) && (yy < p.H) && lane < 30;
          float* dst = yn + (long long)yy * p.W + x;
          float* xc = xch + (size_t)((t * 4 + wq) * 5) * Np;
          const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * g.T * g.Ncol + t * g.Ncol);
          for (int c0 = 0; c0 < Np; c0 += 8) {
            float e0[8], e1[8], e2[8];
            tc_ld8x3(trow + c0, trow + Np + c0, trow + 2 * Np + c0, e0, e1, e2);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = c0 + j;
              const float s1 = __shfl_down_sync(0xffffffffu, e1[j], 1);
              const float s2 = __shfl_down_sync(0xffffffffu, e2[j], 2);
              const float pe = e0[j] + s1;          // complete up to E_2 for lanes <= 30
              if (lane == 0) { xc[c] = e1[j]; xc[Np + c] = e2[j]; }
              else if (lane == 1) xc[2 * Np + c] = e2[j];
              else if (lane == 30) xc[3 * Np + c] = pe;
              else if (lane == 31) xc[4 * Np + c] = e0[j];
              if (valid && c < p.Cout) {
                const float bv = p.bias ? __ldg(p.bias + c) : 0.f;
                dst[(long long)c * HW] = fmaf(pe + s2, oscale, bv);
              }
            }
          }
        }
        // the accumulators have been read: hand them back to the MMA warp before the lane fix-up
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce + 8 * as);
        if (++as == g.acc_stages) { as = 0; aph ^= 1; }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
        // lanes 30 / 31 of every quarter: E_2 (and E_1) of the next quarter's lanes 0 / 1
        for (int t = egrp; t < g.T; t += EG) {
          const int tn = (wq == 3) ? t + 1 : t, wn = (wq == 3) ? 0 : wq + 1;
          if (tn >= g.T) continue;
          const float* xc = xch + (size_t)((t * 4 + wq) * 5) * Np;
          const float* xn = xch + (size_t)((tn * 4 + wn) * 5) * Np;
          for (int idx = lane; idx < 2 * Np; idx += 32) {
            const int l = idx >= Np ? 1 : 0, c = idx - l * Np;
            const int q = t * 128 + wq * 32 + 30 + l;
            const int r = q / g.Wp, x = q - r * g.Wp;
            const int yy = y0 + r;
            if (r < g.R && x < p.W && yy < p.H && c < p.Cout) {
              const float v = l ? (xc[4 * Np + c] + xn[c]) + xn[2 * Np + c] : xc[3 * Np + c] + xn[Np + c];
              const float bv = p.bias ? __ldg(p.bias + c) : 0.f;
              yn[(long long)c * HW + (long long)yy * p.W + x] = fmaf(v, oscale, bv);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");   // exchange area free for the next unit
        continue;
      }
      for (int t = egrp; t < g.T; t += EG) {
        const int q = t * 128 + wq * 32 + lane;
        const int r = q / g.Wp, x = q - r * g.Wp;
        const int yy = y0 + r;
        const bool valid = (r < g.R) && (x < p.W) && (yy < p.H);
        // channel c_base of this pixel; the chunk loop advances it by 8 planes, a store by one: the per-store 64-bit
        // (c_base + c) * HW products, the per-channel bias predicates and a divergent branch around every store made the
        // stores 12 instructions per element (ncu: 87 M of the kernel's 309 M warp instructions on 18->18 @320)
        float* dstc = yn + (long long)yy * p.W + x + (long long)c_base * HW;
        const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * g.T * g.Ncol + t * g.Ncol);
        for (int c0 = 0; c0 < c_cnt; c0 += 8, dstc += 8 * HW) {
          float v[8];
          if (g.hls) {                 // hi*hi + lo*hi in columns [0, Npad), hi*lo in [Npad, 2 Npad): same pixel, same thread
            float v2[8];
            tc_ld8x2(trow + c0, trow + g.Npad + c0, v, v2);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += v2[j];
          } else {
            tc_ld8(trow + c0, v);
          }
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], oscale, (c0 + j < c_cnt) ? __ldg(p.bias + c_base + c0 + j) : 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= oscale;
          }
          const int nst = valid ? min(8, c_cnt - c0) : 0;      // channels of this chunk this thread stores
          float* pc = dstc;
#pragma unroll
          for (int j = 0; j < 8; ++j, pc += HW)
            if (j < nst) *pc = v[j];
          if (p.sums) {
            float w16[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = valid ? v[j] : 0.f;
              w16[j] = x; w16[8 + j] = x * x;
            }
            const float tot = warp_sum16(w16, lane);
            const int vi = (lane >> 1) & 15, c = c0 + (vi & 7);
            if (!(lane & 1) && c < c_cnt) sacc[2 * c + (vi >> 3)] += tot;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8 * as);
      if (++as == g.acc_stages) { as = 0; aph ^= 1; }
    }
    if (p.sums && seg >= 0) tc_flush_stats(p, g, sacc, seg, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------- operand staging
constexpr int STAGE_MAX_TERMS = 6;
struct StageTerm {
  const float* y;    // source tensor (fp32 NCHW at the source resolution)
  const float* mu;   // per-plane centre, scale, shift (a null = identity); plane = n*C + c
  const float* a;
  const float* b;
  float slope;       // leaky slope (1 = none)
  int C;             // channels of this term
  int mode;          // 0 direct, 1 avg-pool 2x2 of the activated source (source is 2H x 2W),
                     // 2 depth-to-space (source [N,4C,H/2,W/2]), 3 nearest x2 (source is H/2 x W/2)
  int c0;            // first output channel this term writes / adds to
};
struct StageArgs {
  StageTerm s[STAGE_MAX_TERMS];
  int nterms;
  __nv_bfloat16* xs;
  int N, H, W, Wp, PS, KG;
  int fmt;           // 0: bf16 pairs, 1: fp16 pairs scaled by TC_SX (forward operands) or by the dynamic scale below
  const float* absmax;   // null, or device scalar max|y| of the (single, identity) source: dynamic scale (gradients)
};

__device__ __forceinline__ float act1(float v, float mu, float a, float b, float slope) {
  const float z = fmaf(a, v - mu, b);
  return z > 0.f ? z : z * slope;
}

// ================================================================================== fused-staging conv ("row ring")
// The narrow full-resolution 3x3 layers (<= 24 input, <= 32 output channels) WITHOUT a staged operand in HBM on the way in:
// the CTA converts raw fp32 rows itself.  A CTA walks a contiguous range of output rows of one image after the other; the
// padded input rows it needs live in a RING of shared-memory row slots, each holding the row as the canonical K-major
// operand planes [hl][kg][slot][8 x 16 bit] - with one-row units every filter-tap window of a 128-pixel tile lies inside ONE
// padded row, so the three rows of a unit need not be contiguous and every row is converted ONCE per CTA (a strip-wise
// kernel would convert each row three times).  Converter warps read the raw rows (coalesced per channel plane), apply the
// producing layer's normalisation + LeakyReLU (or nothing: gradients), split into the fp16 pair and write the planes; one
// thread optionally copies each finished row to the staged tensor in HBM with TMA bulk stores (the weight-gradient GEMM
// still reads that form); the MMA warp addresses taps as shifted descriptors into the ring; all K groups and the whole
// weight block stay resident, so there is no per-unit operand reload at all.  Form: HLS ([W_hi | W_lo] along N), taps of the
// half-empty last K-step paired INSIDE a filter row (a pair across rows would need the ring distance as a descriptor offset).
//   warp 0       weight load (once), TMA row stores          warp 1        TMEM + MMA issue
//   warps 2..17  epilogue (+ statistics) as in conv_tc_kernel   warps 18..25  converters
constexpr int RR_EPI_WARPS = 12;      // three per TMEM lane quarter: one per 128-pixel tile of a 320-wide row
constexpr int RR_CONV_WARPS = 16;
constexpr int RR_THREADS = 64 + 32 * RR_EPI_WARPS + 32 * RR_CONV_WARPS;      // 960
constexpr int RR_NR_MAX = 8;      // ring slots: 3 rows in use + the rows being converted ahead (as many as shared memory allows)
constexpr int RR_CMAX = 24;       // input channels (3 groups of 8)
constexpr int RR_GROUPS = 2;      // converter groups: group g converts rows k = g (mod 2), so two rows' loads are in flight
constexpr int RR_GT = RR_CONV_WARPS * 32 / RR_GROUPS;      // threads per converter group
constexpr int RR_MI = 4;          // (pixel slot, channel group) items per converter thread and row

struct RrGeom {
  int KG, KS, Npad, Ncol, Wp, Hp, PS, T, RS, pair;
  int plane_bytes, row_bytes, wblk_bytes, w_bytes, NR;
  int w_off, ring_off, coef_off, stat_off, smem_bytes;
};

bool rr_geometry(int H, int W, int Cin, int Cout, int K, bool stats, RrGeom* g) {
  if (K != 3 || Cin > RR_CMAX || pad16(Cout) > 32 || W < 30 || H < 2) return false;
  g->KG = (Cin + 7) / 8; g->KS = (g->KG + 1) / 2; g->pair = (g->KG & 1) ? 2 : 0;
  g->Npad = pad16(Cout); g->Ncol = 2 * g->Npad;
  g->Wp = W + 2; g->Hp = H + 2; g->PS = g->Hp * g->Wp;
  g->T = (g->Wp + 127) / 128;
  if (2 * g->T * g->Ncol > 512 || g->Wp * g->KG > RR_MI * RR_GT) return false;
  g->RS = 128 * g->T + 8;
  g->plane_bytes = g->RS * 16;
  g->row_bytes = 2 * g->KG * g->plane_bytes;
  g->wblk_bytes = 4 * g->Npad * 16;
  g->w_bytes = g->KS * 9 * g->wblk_bytes;
  g->w_off = TC_SMEM_HEADER;
  g->ring_off = g->w_off + g->w_bytes;
  // The row -> MMA -> "slot free" -> next row chain is a latency loop: with NR slots NR - 2 rows are in flight around it
  const int fixed = g->ring_off + RR_GROUPS * RR_CMAX * 16 + (stats ? RR_EPI_WARPS * 2 * g->Npad * (int)sizeof(float) : 0);
  g->NR = (TC_SMEM_MAX - fixed) / g->row_bytes;
  if (g->NR > RR_NR_MAX) g->NR = RR_NR_MAX;
  if (g->NR < 4) return false;
  g->coef_off = g->ring_off + g->NR * g->row_bytes;
  g->stat_off = g->coef_off + RR_GROUPS * RR_CMAX * 16;
  g->smem_bytes = g->stat_off + (stats ? RR_EPI_WARPS * 2 * g->Npad * (int)sizeof(float) : 0);
  if (g->smem_bytes > TC_SMEM_MAX) return false;
  if (g->smem_bytes < 120 * 1024) g->smem_bytes = 120 * 1024;      // one CTA per SM (all 512 TMEM columns)
  return true;
}

struct RrParams {
  const float* x;            // raw fp32 input [N][Cin][H][W]
  const float* mu;           // per-plane centre / scale / shift of the producing layer's normalisation (a null: identity)
  const float* a;
  const float* b;
  float slope;
  const float* absmax;       // null: static activation scale; else device scalar max|x| -> dynamic scale (gradient operand)
  __nv_bfloat16* xs_out;     // null, or the staged tensor (past the lead-in) that receives every converted row
  const __nv_bfloat16* ws;   // staged weights, HLS layout, pair = 2 in the last K-step when KG is odd
  const float* bias;
  float* y;
  double* sums;
  int N, H, W, Cin, Cout, nunits;
  float out_scale;
  RrGeom g;
};

__device__ __forceinline__ void rr_range(const RrParams& p, int& u0, int& u1) {
  u0 = (int)((long long)blockIdx.x * p.nunits / gridDim.x);
  u1 = (int)((long long)(blockIdx.x + 1) * p.nunits / gridDim.x);
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(RR_THREADS, 1) conv_rows_kernel(const RrParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const RrGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t hdr = smem_u32(smem);
  const uint32_t bar_rfull = hdr, bar_rempty = hdr + 64, bar_accf = hdr + 128, bar_acce = hdr + 144, bar_w = hdr + 160;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + 168);
  const uint32_t w0 = hdr + g.w_off, ring0 = hdr + g.ring_off;
  const int NRr = g.NR;                                                  // ring slots (<= RR_NR_MAX)

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.NR; ++i) { mbar_init(bar_rfull + 8 * i, RR_CONV_WARPS / RR_GROUPS); mbar_init(bar_rempty + 8 * i, p.xs_out ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, RR_EPI_WARPS); }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((void*)tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {   // the slots past a row's Wp pixels are read by the last tile's windows (results dropped): finite values once
    uint4* z = (uint4*)(smem + g.ring_off);
    const int n16 = g.NR * g.row_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  int u0, u1;
  rr_range(p, u0, u1);

  if (warp == 0) {
    // ============================ weights (once) + TMA stores of the converted rows ============================
    if (lane == 0) {
      mbar_expect_tx(bar_w, (uint32_t)g.w_bytes);
      bulk_g2s(w0, p.ws, (uint32_t)g.w_bytes, bar_w);
      if (p.xs_out) {
        const uint32_t bytes = (uint32_t)g.Wp * 16;
        int k = 0;
        for (int u = u0; u < u1; ++u) {
          const int n = u / p.H, yy = u - n * p.H;
          const bool fresh = (u == u0) || (yy == 0);
          for (int r = fresh ? 0 : 2; r < 3; ++r, ++k) {
            const int pr = yy + r, s = k % NRr;
            mbar_wait_sleep(bar_rfull + 8 * s, (uint32_t)(k / NRr) & 1u);
            const uint32_t src0 = ring0 + (uint32_t)s * g.row_bytes;
            for (int hl = 0; hl < 2; ++hl)
              for (int kg = 0; kg < g.KG; ++kg) {
                __nv_bfloat16* dst = p.xs_out + (((long long)(n * 2 + hl) * g.KG + kg) * g.PS + (long long)pr * g.Wp) * 8;
                bulk_s2g(dst, src0 + (uint32_t)(hl * g.KG + kg) * g.plane_bytes, bytes);
              }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (k > 0) {      // the previous row's stores have read their shared-memory source: its slot may be refilled
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              mbar_arrive(bar_rempty + 8 * ((k - 1) % NRr));
            }
          }
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (k > 0) mbar_arrive(bar_rempty + 8 * ((k - 1) % NRr));
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
  } else if (warp == 1) {
    // ========================================= MMA issue =========================================
    const uint32_t idesc = umma_idesc_16(128, g.Npad, 3), idesc2 = umma_idesc_16(128, 2 * g.Npad, 3);
    const uint32_t hiw = (uint32_t)(umma_desc(0, 0, 128) >> 32);        // SBO 128 B, version bit
    const uint32_t PPu = (uint32_t)g.RS;                                 // plane pitch in 16 B units
    const uint32_t a_lbo_pp = PPu << 16;                                 // K group 1 = the next channel-group plane
    const uint32_t b_lbo = (uint32_t)(2 * g.Npad) << 16;                 // K group 1 of B = the next kk plane of the block
    const uint32_t wblk16 = (uint32_t)g.wblk_bytes >> 4;
    mbar_wait_warp(bar_w, 0, lane);
    int k_next = 0, as = 0;
    uint32_t aph = 0;
    for (int u = u0; u < u1; ++u) {
      const int yy = u % p.H;
      const bool fresh = (u == u0) || (yy == 0);
      const int nnew = fresh ? 3 : 1;
      const int base = fresh ? k_next : k_next - 2;
      for (int c = k_next; c < k_next + nnew; ++c) mbar_wait_warp(bar_rfull + 8 * (c % NRr), (uint32_t)(c / NRr) & 1u, lane);
      k_next += nnew;
      mbar_wait_warp(bar_acce + 8 * as, aph ^ 1, lane);
      tc_fence_after();
      const bool next_fresh = (u + 1 == u1) || ((u + 1) % p.H == 0);
      if (elect_one_sync()) {
        uint32_t row16[3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) row16[dy] = (ring0 + (uint32_t)((base + dy) % NRr) * g.row_bytes) >> 4;
        const uint32_t losplit = (uint32_t)g.KG * PPu;                   // hi -> lo planes of a row slot
        const uint32_t acc0 = tmem_base + (uint32_t)(as * g.T * g.Ncol);
        for (int t = 0; t < g.T; ++t) {
          const uint32_t d = acc0 + (uint32_t)(t * g.Ncol);
          uint32_t started = 0;
          for (int ks = 0; ks < g.KS; ++ks) {
            const uint32_t bks = (w0 >> 4) + (uint32_t)ks * 9u * wblk16;
            if (g.pair && ks == g.KS - 1) {
              const uint32_t goff = (uint32_t)(g.KG - 1) * PPu + (uint32_t)t * 128u;
#pragma unroll
              for (int blk = 0; blk < 6; ++blk) {
                const int dy = blk >> 1;
                const uint32_t al = row16[dy] + goff + ((blk & 1) ? 2u : 0u) + ((blk & 1) ? 0u : (1u << 16));   // pair: LBO = 1 slot
                const uint32_t bl = bks + (uint32_t)blk * wblk16 + b_lbo;
                const uint64_t A_hi = ((uint64_t)hiw << 32) | al, A_lo = ((uint64_t)hiw << 32) | (al + losplit);
                const uint64_t B_all = ((uint64_t)hiw << 32) | bl;
                tc_mma_bf16(d, A_hi, B_all, idesc2, started);
                tc_mma_bf16(d, A_lo, B_all, idesc, 1u);
                started = 1;
              }
            } else {
              const uint32_t goff = (uint32_t)(2 * ks) * PPu + (uint32_t)t * 128u + a_lbo_pp;
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t al = row16[tap / 3] + goff + (uint32_t)(tap % 3);
                const uint32_t bl = bks + (uint32_t)tap * wblk16 + b_lbo;
                const uint64_t A_hi = ((uint64_t)hiw << 32) | al, A_lo = ((uint64_t)hiw << 32) | (al + losplit);
                const uint64_t B_all = ((uint64_t)hiw << 32) | bl;
                tc_mma_bf16(d, A_hi, B_all, idesc2, started);
                tc_mma_bf16(d, A_lo, B_all, idesc, 1u);
                started = 1;
              }
            }
          }
        }
        tc_commit(bar_accf + 8 * as);
        // rows no later unit reads: the oldest one, or all three at the end of an image / of the range
        tc_commit(bar_rempty + 8 * (base % NRr));
        if (next_fresh) { tc_commit(bar_rempty + 8 * ((base + 1) % NRr)); tc_commit(bar_rempty + 8 * ((base + 2) % NRr)); }
      }
      __syncwarp();
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else if (warp < 2 + RR_EPI_WARPS) {
    // ========================================= epilogue =========================================
    const int wq = warp & 3, egrp = (warp - 2) >> 2;
    constexpr int EG = RR_EPI_WARPS / 4;
    const float oscale = p.absmax ? p.out_scale / tc_dyn_scale(__ldg(p.absmax)) : p.out_scale;
    const long long HW = (long long)p.H * p.W;
    float* sacc = (float*)(smem + g.stat_off) + (size_t)(warp - 2) * (2 * g.Npad);
    int seg = -1;
    if (p.sums)
      for (int i = lane; i < 2 * g.Npad; i += 32) sacc[i] = 0.f;
    int as = 0;
    uint32_t aph = 0;
    for (int u = u0; u < u1; ++u) {
      const int n = u / p.H, yy = u - n * p.H;
      if (p.sums && seg != n) {
        if (seg >= 0) {
          __syncwarp();
          for (int i = lane; i < 2 * p.Cout; i += 32) {
            const float v = sacc[i];
            if (v != 0.f) atomicAdd(p.sums + (long long)seg * p.Cout * 2 + i, (double)v);
            sacc[i] = 0.f;
          }
          __syncwarp();
        }
        seg = n;
      }
      mbar_wait_warp(bar_accf + 8 * as, aph, lane);
      tc_fence_after();
      float* yn = p.y + (long long)n * p.Cout * HW + (long long)yy * p.W;
      for (int t = egrp; t < g.T; t += EG) {
        const int x = t * 128 + wq * 32 + lane;
        const bool valid = x < p.W;
        float* dst = yn + x;
        const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * g.T * g.Ncol + t * g.Ncol);
        for (int c0 = 0; c0 < p.Cout; c0 += 8) {
          float v[8], v2[8];
          tc_ld8x2(trow + c0, trow + g.Npad + c0, v, v2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const float bv = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
            v[j] = fmaf(v[j] + v2[j], oscale, bv);
            if (valid && c < p.Cout) dst[(long long)c * HW] = v[j];
          }
          if (p.sums) {
            float w16[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float xv = valid ? v[j] : 0.f;
              w16[j] = xv; w16[8 + j] = xv * xv;
            }
            const float tot = warp_sum16(w16, lane);
            const int vi = (lane >> 1) & 15, c = c0 + (vi & 7);
            if (!(lane & 1) && c < p.Cout) sacc[2 * c + (vi >> 3)] += tot;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8 * as);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (p.sums && seg >= 0) {
      __syncwarp();
      for (int i = lane; i < 2 * p.Cout; i += 32) {
        const float v = sacc[i];
        if (v != 0.f) atomicAdd(p.sums + (long long)seg * p.Cout * 2 + i, (double)v);
      }
    }
  } else {
    // ========================================= converters =========================================
    // A thread owns up to RR_MI (pixel slot, channel group) items of every row - the same ones for every row - and keeps the
    // raw values of the NEXT row in registers: the global loads are issued right after a row is handed over and complete
    // while the thread waits for the ring slot, so only arithmetic + shared-memory stores sit between "slot free" and "row
    // full" (with the loads after the wait every row cost a full memory latency: 4 us per row, measured).
    const int cw = warp - 2 - RR_EPI_WARPS;
    const int grp = cw / (RR_CONV_WARPS / RR_GROUPS);                  // this warp's group converts rows k = grp (mod RR_GROUPS)
    const int ct = (cw % (RR_CONV_WARPS / RR_GROUPS)) * 32 + lane;     // 0 .. RR_GT - 1
    float4* coef = (float4*)(smem + g.coef_off) + grp * RR_CMAX;       // per channel of the group's current image: mu, a, b, -
    const float scale = p.absmax ? tc_dyn_scale(__ldg(p.absmax)) : TC_SX;
    const long long HW = (long long)p.H * p.W;
    const int items = g.Wp * g.KG;                                     // <= RR_MI * 512 (rr_geometry)
    // per item (fixed for the whole kernel): channel group, pixel slot, element offset of its first channel at column x - 1
    // inside an image, and how many of its 8 channels exist (the loads of the missing ones repeat the last real channel and
    // are zeroed afterwards: unconditional loads, no per-load predicates / descriptor moves - the first version spent 700
    // warp instructions per row and warp, the kernel was issue-bound)
    int it_kg[RR_MI], it_x[RR_MI], it_off[RR_MI], it_nreal[RR_MI];
    bool it_in[RR_MI];
#pragma unroll
    for (int m = 0; m < RR_MI; ++m) {
      const int i = ct + m * RR_GT;
      it_kg[m] = i < items ? i / g.Wp : -1;
      it_x[m] = i < items ? i - it_kg[m] * g.Wp : 0;
      it_in[m] = it_kg[m] >= 0 && it_x[m] >= 1 && it_x[m] <= p.W;
      const int kg = it_kg[m] < 0 ? 0 : it_kg[m];
      it_nreal[m] = min(8, p.Cin - kg * 8);
      it_off[m] = kg * 8 * (int)HW + (it_in[m] ? it_x[m] - 1 : 0);
    }
    const int HWi = (int)HW;
    float va[RR_MI][8];
    auto load_row = [&](int u, int r, float (&v)[RR_MI][8]) {
      const int n = u / p.H, pr = (u - n * p.H) + r;                   // padded row pr = image row pr - 1
      const int yrow = min(max(pr - 1, 0), p.H - 1);                   // (border rows load row 0 / H - 1 and are zeroed)
      const float* xrow = p.x + (long long)n * p.Cin * HW + (long long)yrow * p.W;
#pragma unroll
      for (int m = 0; m < RR_MI; ++m) {
        if (it_kg[m] >= 0) {
          const float* src = xrow + it_off[m];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[m][j] = __ldg(src + min(j, it_nreal[m] - 1) * HWi);
        }
      }
    };
    int n_cur = -1;
    auto convert_row = [&](int u, int r, int k, float (&v)[RR_MI][8]) {
      const int n = u / p.H, pr = (u - n * p.H) + r, s = k % NRr;
      if (n != n_cur) {
        asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(RR_GT) : "memory");      // the group is done with its old table
        if (ct < RR_CMAX) {
          float4 cf = make_float4(0.f, 1.f, 0.f, 0.f);
          if (p.a && ct < p.Cin) {
            const long long pl = (long long)n * p.Cin + ct;
            cf = make_float4(p.mu ? __ldg(p.mu + pl) : 0.f, __ldg(p.a + pl), p.b ? __ldg(p.b + pl) : 0.f, 0.f);
          }
          coef[ct] = cf;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(RR_GT) : "memory");
        n_cur = n;
      }
      mbar_wait_warp(bar_rempty + 8 * s, ((uint32_t)(k / NRr) & 1u) ^ 1u, lane);
      uint8_t* rowp = smem + g.ring_off + (size_t)s * g.row_bytes;
      const bool inb_row = pr >= 1 && pr <= p.H;
#pragma unroll
      for (int m = 0; m < RR_MI; ++m) {
        if (it_kg[m] >= 0) {
          const bool inb = inb_row && it_in[m];
          uint32_t hw[4], lw[4];
          if (p.a) {
            const float4* cf8 = coef + it_kg[m] * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 cf = cf8[j];
              const float z = act1(v[m][j], cf.x, cf.y, cf.z, p.slope);
              v[m][j] = (inb && j < it_nreal[m]) ? z : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[m][j] = (inb && j < it_nreal[m]) ? v[m][j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) split16x2(v[m][2 * j], v[m][2 * j + 1], true, scale, hw[j], lw[j]);
          uint4* dst = (uint4*)(rowp + (size_t)(it_kg[m] * g.RS + it_x[m]) * 16);
          dst[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          dst[(size_t)g.KG * g.RS] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores before UMMA / TMA reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_rfull + 8 * s);
    };
    // row sequence: (u, r) -> next; the first unit of an image / of the range brings three new rows, every other unit one
    auto advance = [&](int& u, int& r) {
      if (r < 2) {
        ++r;
      } else {
        ++u;
        r = ((u % p.H) == 0) ? 0 : 2;
      }
    };
    // one register set per thread; the two groups alternate rows, so while one group waits for its loads the other converts
    int u = u0, r = 0, k = 0;
    auto skip_to_mine = [&]() {
      while (u < u1 && (k % RR_GROUPS) != grp) { advance(u, r); ++k; }
    };
    skip_to_mine();
    if (u < u1) load_row(u, r, va);
    while (u < u1) {
      convert_row(u, r, k, va);
      advance(u, r); ++k;
      skip_to_mine();
      if (u < u1) load_row(u, r, va);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Per-block work list: for each of the 8 channels of the block's channel group, the (at most
// STAGE_MAX_SUM) terms that contribute to it, with everything that does not depend on the pixel slot
// resolved once (plane base pointer, coefficients).  Missing entries are neutral dummies (a = 0, a
// valid address), so the slot loop has no data-dependent control flow: the 8 (x2 slots) loads of one
// term level are issued back to back and their latencies overlap.  (The first version walked the terms
// per channel with dynamic loops: 8 serialised global loads per pixel, latency-bound at 2.7 TB/s.)
constexpr int STAGE_MAX_SUM = 3;
struct StageEntry {
  const float* base;   // plane base (mode 2: base of the 4 sub-planes of the channel)
  float mu, a, b, slope;
  int mode;            // 0 direct, 1 pool, 2 pixel shuffle, 3 nearest up, 4 dummy
  int pad;
};
__device__ float g_stage_zero[4] = {0.f, 0.f, 0.f, 0.f};

struct SlotOffs { int off0, offu, offs2, offp; bool inb; };
__device__ __forceinline__ SlotOffs slot_offsets(int slot, const StageArgs& A, int Wh, int HWq, int W2) {
  const int hp = slot / A.Wp, wp = slot - hp * A.Wp;
  const int h = hp - 1, w = wp - 1;
  SlotOffs o;
  o.inb = (slot < A.PS) && h >= 0 && h < A.H && w >= 0 && w < A.W;
  const int hh = o.inb ? h : 0, ww = o.inb ? w : 0;
  o.off0 = hh * A.W + ww;
  o.offu = (hh >> 1) * Wh + (ww >> 1);
  o.offs2 = o.offu + ((hh & 1) * 2 + (ww & 1)) * HWq;
  o.offp = (2 * hh) * W2 + 2 * ww;
  return o;
}
__device__ __forceinline__ int pick_off(int mode, const SlotOffs& o) {
  return mode == 0 ? o.off0 : (mode == 2 ? o.offs2 : (mode == 3 ? o.offu : 0));
}
__device__ __forceinline__ void store_split(__nv_bfloat16* xs, long long o_hi, long long o_lo, int slot, const float* v,
                                            bool f16, float scale) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split16x2(v[2 * j], v[2 * j + 1], f16, scale, hw[j], lw[j]);
  *(uint4*)(xs + (o_hi + slot) * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *(uint4*)(xs + (o_lo + slot) * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// one thread = two pixel slots x one channel group of 8 per iteration: four 16 B stores (hi, lo).
// grid = (slot blocks, N*KG).  A channel is the SUM of every term whose channel range [c0, c0 + C)
// contains it (residual adds of unet.py:15-24 are sums of two activated tensors; concatenation =
// disjoint ranges).
__global__ void __launch_bounds__(256, 3) stage_act_kernel(const StageArgs A) {
  __shared__ StageEntry ent[STAGE_MAX_SUM][8];
  __shared__ int s_kmax, s_pool;
  const int n = blockIdx.y / A.KG, kg = blockIdx.y - n * A.KG;
  if (threadIdx.x == 0) { s_kmax = 0; s_pool = 0; }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int j = threadIdx.x;
    int k = 0;
    for (int t = 0; t < A.nterms && k < STAGE_MAX_SUM; ++t) {
      const StageTerm& S = A.s[t];
      const int c = kg * 8 + j - S.c0;
      if (c < 0 || c >= S.C) continue;
      const long long plane = (long long)n * S.C + c;
      StageEntry E;
      E.mode = S.mode; E.slope = S.slope; E.pad = 0;
      E.mu = 0.f; E.a = 1.f; E.b = 0.f;
      if (S.a) { E.mu = S.mu ? S.mu[plane] : 0.f; E.a = S.a[plane]; E.b = S.b ? S.b[plane] : 0.f; }
      const long long src_hw = (S.mode == 0) ? (long long)A.H * A.W : (S.mode == 1) ? 4LL * A.H * A.W : (long long)(A.H / 2) * (A.W / 2);
      E.base = S.y + (S.mode == 2 ? plane * 4 : plane) * src_hw;
      if (S.mode == 1) atomicOr(&s_pool, 1);
      ent[k++][j] = E;
    }
    atomicMax(&s_kmax, k);
    for (; k < STAGE_MAX_SUM; ++k) ent[k][j] = StageEntry{g_stage_zero, 0.f, 0.f, 0.f, 1.f, 4, 0};
  }
  __syncthreads();
  const int kmax = s_kmax;
  const bool pooled = s_pool != 0;
  const bool f16 = A.fmt != 0;
  const float scale = A.absmax ? tc_dyn_scale(__ldg(A.absmax)) : TC_SX;
  const long long o_hi = ((long long)(n * 2 + 0) * A.KG + kg) * A.PS;
  const long long o_lo = ((long long)(n * 2 + 1) * A.KG + kg) * A.PS;
  const int Wh = A.W / 2, HWq = (A.H / 2) * Wh, W2 = 2 * A.W;
  const int stride = gridDim.x * blockDim.x;
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < A.PS; slot += 2 * stride) {
    const int slotB = slot + stride;
    const SlotOffs oA = slot_offsets(slot, A, Wh, HWq, W2), oB = slot_offsets(slotB, A, Wh, HWq, W2);
    float vA[8], vB[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { vA[j] = 0.f; vB[j] = 0.f; }
    if (!pooled) {
      for (int k = 0; k < kmax; ++k) {
        float rA[8], rB[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const StageEntry& E = ent[k][j];
          rA[j] = __ldg(E.base + pick_off(E.mode, oA));
          rB[j] = __ldg(E.base + pick_off(E.mode, oB));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const StageEntry& E = ent[k][j];
          vA[j] += act1(rA[j], E.mu, E.a, E.b, E.slope);
          vB[j] += act1(rB[j], E.mu, E.a, E.b, E.slope);
        }
      }
    } else {
      for (int k = 0; k < kmax; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const StageEntry& E = ent[k][j];
          if (E.mode == 1) {
            const float* qa = E.base + oA.offp;
            const float* qb = E.base + oB.offp;
            const float2 a0 = __ldg((const float2*)qa), a1 = __ldg((const float2*)(qa + W2));
            const float2 b0 = __ldg((const float2*)qb), b1 = __ldg((const float2*)(qb + W2));
            vA[j] += 0.25f * ((act1(a0.x, E.mu, E.a, E.b, E.slope) + act1(a0.y, E.mu, E.a, E.b, E.slope)) +
                              (act1(a1.x, E.mu, E.a, E.b, E.slope) + act1(a1.y, E.mu, E.a, E.b, E.slope)));
            vB[j] += 0.25f * ((act1(b0.x, E.mu, E.a, E.b, E.slope) + act1(b0.y, E.mu, E.a, E.b, E.slope)) +
                              (act1(b1.x, E.mu, E.a, E.b, E.slope) + act1(b1.y, E.mu, E.a, E.b, E.slope)));
          } else {
            vA[j] += act1(__ldg(E.base + pick_off(E.mode, oA)), E.mu, E.a, E.b, E.slope);
            vB[j] += act1(__ldg(E.base + pick_off(E.mode, oB)), E.mu, E.a, E.b, E.slope);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { vA[j] = oA.inb ? vA[j] : 0.f; vB[j] = oB.inb ? vB[j] : 0.f; }
    store_split(A.xs, o_hi, o_lo, slot, vA, f16, scale);
    if (slotB < A.PS) store_split(A.xs, o_hi, o_lo, slotB, vB, f16, scale);
  }
  // zero lead-in / trailing slack of the buffer (see TC_LEAD / TC_TRAIL)
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) *(uint4*)(A.xs - TC_LEAD) = z;
    if (threadIdx.x < TC_TRAIL / 8) *(uint4*)(A.xs + (long long)A.N * 2 * A.KG * A.PS * 8 + threadIdx.x * 8) = z;
  }
}

// Lean variant for the common case - every channel comes from exactly ONE same-resolution source (conv -> conv inside a
// block, the skip / residual-free concatenations, every dY staging): no resampling offsets, no term loop, no pooled
// branch -> 40 instead of 78 registers, twice the resident warps (the general kernel sits at 54 % of the DRAM peak with
// 35 % occupancy, profiles/r2i_ncu_summary_stage_wgrad_norm.txt).  Same work decomposition and the same arithmetic.
__global__ void __launch_bounds__(256, 4) stage_simple_kernel(const StageArgs A) {
  __shared__ StageEntry ent[8];
  const int n = blockIdx.y / A.KG, kg = blockIdx.y - n * A.KG;
  if (threadIdx.x < 8) {
    const int j = threadIdx.x;
    StageEntry E{g_stage_zero, 0.f, 0.f, 0.f, 1.f, 4, 0};
    for (int t = 0; t < A.nterms; ++t) {
      const StageTerm& S = A.s[t];
      const int c = kg * 8 + j - S.c0;
      if (c < 0 || c >= S.C) continue;
      const long long plane = (long long)n * S.C + c;
      E.mode = 0; E.slope = S.slope; E.mu = 0.f; E.a = 1.f; E.b = 0.f;
      if (S.a) { E.mu = S.mu ? S.mu[plane] : 0.f; E.a = S.a[plane]; E.b = S.b ? S.b[plane] : 0.f; }
      E.base = S.y + plane * (long long)A.H * A.W;
      break;
    }
    ent[j] = E;
  }
  __syncthreads();
  const bool f16 = A.fmt != 0;
  const float scale = A.absmax ? tc_dyn_scale(__ldg(A.absmax)) : TC_SX;
  const long long o_hi = ((long long)(n * 2 + 0) * A.KG + kg) * A.PS;
  const long long o_lo = ((long long)(n * 2 + 1) * A.KG + kg) * A.PS;
  const int stride = gridDim.x * blockDim.x;
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < A.PS; slot += 2 * stride) {
    const int slotB = slot + stride;
    int offA, offB;
    bool inA, inB;
    {
      const int hp = slot / A.Wp, wp = slot - hp * A.Wp;
      inA = hp >= 1 && hp <= A.H && wp >= 1 && wp <= A.W;
      offA = inA ? (hp - 1) * A.W + wp - 1 : 0;
      const int hq = slotB / A.Wp, wq = slotB - hq * A.Wp;
      inB = slotB < A.PS && hq >= 1 && hq <= A.H && wq >= 1 && wq <= A.W;
      offB = inB ? (hq - 1) * A.W + wq - 1 : 0;
    }
    float vA[8], vB[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* base = ent[j].mode == 4 ? g_stage_zero : ent[j].base;
      vA[j] = __ldg(base + (ent[j].mode == 4 ? 0 : offA));
      vB[j] = __ldg(base + (ent[j].mode == 4 ? 0 : offB));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const StageEntry& E = ent[j];
      vA[j] = inA ? act1(vA[j], E.mu, E.a, E.b, E.slope) : 0.f;
      vB[j] = inB ? act1(vB[j], E.mu, E.a, E.b, E.slope) : 0.f;
    }
    store_split(A.xs, o_hi, o_lo, slot, vA, f16, scale);
    if (slotB < A.PS) store_split(A.xs, o_hi, o_lo, slotB, vB, f16, scale);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) *(uint4*)(A.xs - TC_LEAD) = z;
    if (threadIdx.x < TC_TRAIL / 8) *(uint4*)(A.xs + (long long)A.N * 2 * A.KG * A.PS * 8 + threadIdx.x * 8) = z;
  }
}

// staged activations back to fp32 NCHW (x = hi + lo): feeds the fp32 weight-gradient kernel
__global__ void __launch_bounds__(256) unstage_act_kernel(const __nv_bfloat16* __restrict__ xs, float* __restrict__ x,
                                                          int N, int C, int H, int W, int KG, int fmt) {
  const int Wp = W + 2, PS = (H + 2) * Wp;
  const long long total = (long long)N * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long t = i / W;
    const int h = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    const int slot = (h + 1) * Wp + w + 1;
    const long long o_hi = (((long long)(n * 2 + 0) * KG + (c >> 3)) * PS + slot) * 8 + (c & 7);
    const long long o_lo = (((long long)(n * 2 + 1) * KG + (c >> 3)) * PS + slot) * 8 + (c & 7);
    x[i] = unsplit16(__bfloat16_as_ushort(xs[o_hi]), __bfloat16_as_ushort(xs[o_lo]), fmt != 0, 1.f / TC_SX);
  }
}

// weights OIHW fp32 -> Ws[nsplit][KS][ntaps][hl][kk][Npad][8] bf16 hi/lo.
// dgrad = 1: the transposed, spatially flipped filter (data gradient = the same conv run on dY):
// "output" channel = original ci, "input" channel = original co.
__global__ void stage_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ ws, int Cout, int Cin,
                                     int KK, int dgrad, int nsplit, int KS, int Npad, int fmt, int dxn_np, int hls, int pair) {
  const int Co_k = dgrad ? Cin : Cout;   // kernel-view output channels
  const int Ci_k = dgrad ? Cout : Cin;   // kernel-view input channels
  if (dxn_np) {
    // DXN form: Ws[KS][dy][hl][kk][Npad][8], B row nn = dx * Np + co  (3x3 only, nsplit = 1)
    const long long total = (long long)KS * 3 * 2 * 2 * Npad * 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      long long t = i;
      const int j = (int)(t % 8); t /= 8;
      const int nn = (int)(t % Npad); t /= Npad;
      const int kk = (int)(t % 2); t /= 2;
      const int hl = (int)(t % 2); t /= 2;
      const int dy = (int)(t % 3);
      const int ks = (int)(t / 3);
      const int dx = nn / dxn_np, co = nn - dx * dxn_np;
      const int ci = ks * 16 + kk * 8 + j;
      const int tap = dy * 3 + dx;
      float v = 0.f;
      if (dx < 3 && co < Co_k && ci < Ci_k) {
        if (!dgrad) v = w[((long long)co * Cin + ci) * 9 + tap];
        else v = w[((long long)ci * Cin + co) * 9 + (8 - tap)];
      }
      unsigned short hi, lo;
      split16(v, fmt != 0, TC_SW, hi, lo);
      ws[i] = __ushort_as_bfloat16(hl ? lo : hi);
    }
    return;
  }
  const long long total = (long long)nsplit * KS * KK * 2 * 2 * Npad * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int j = (int)(t % 8); t /= 8;
    int nn, kk, hl;
    if (hls) {        // [tap][kk][hl * Npad + nn][8]: W_hi and W_lo stacked along the MMA N dimension
      nn = (int)(t % Npad); t /= Npad;
      hl = (int)(t % 2); t /= 2;
      kk = (int)(t % 2); t /= 2;
    } else {          // [tap][hl][kk][nn][8]
      nn = (int)(t % Npad); t /= Npad;
      kk = (int)(t % 2); t /= 2;
      hl = (int)(t % 2); t /= 2;
    }
    int tap = (int)(t % KK); t /= KK;
    const int ks = (int)(t % KS);
    const int ns = (int)(t / KS);
    const int co = ns * Npad + nn;
    int ci = ks * 16 + kk * 8 + j;
    bool live = true;
    if (pair == 2 && ks == KS - 1) {  // row-ring kernel: taps paired inside a filter row: block 2*dy = (dx 0, dx 1), block 2*dy+1 = (dx 2, -)
      ci = ks * 16 + j;
      const int dy = tap >> 1, odd = tap & 1;
      live = tap < 6 && !(odd && kk);
      tap = dy * 3 + (odd ? 2 : kk);
    } else if (pair && ks == KS - 1) {       // block i < 5 holds taps 2i (K group 0) and 2i+1 (K group 1) of the ONE real channel group
      ci = ks * 16 + j;
      live = tap < 5 && 2 * tap + kk < KK;
      tap = 2 * tap + kk;
    }
    float v = 0.f;
    if (live && co < Co_k && nn < Npad && ci < Ci_k) {
      if (!dgrad) v = w[((long long)co * Cin + ci) * KK + tap];
      else v = w[((long long)ci * Cin + co) * KK + (KK - 1 - tap)];
    }
    unsigned short hi, lo;
    split16(v, fmt != 0, TC_SW, hi, lo);
    ws[i] = __ushort_as_bfloat16(hl ? lo : hi);
  }
}

inline int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)san_num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" {

long long san_tc_staged_act_elems(int N, int H, int W, int C) {
  return (long long)N * 2 * ((C + 7) / 8) * (long long)(H + 2) * (W + 2) * 8 + TC_LEAD + TC_TRAIL;
}

long long san_tc_staged_weight_elems(int H, int W, int Cout, int Cin, int K) {
  TcGeom g;
  if (!tc_geometry(H, W, Cin, Cout, K, &g)) return -1;
  return (long long)g.nsplit * g.KS * g.wtaps * 4 * g.Npad * 8;
}

// Host-only: the strip geometry the kernel would use, for tests / tooling.  out[0..15] = Cin_pad, KG, KS, nsplit,
// Npad, Wp, Hp, R, T, S_alloc, strips, stages, acc_stages, a_bytes, b_bytes, smem_bytes.
int san_tc_describe(int H, int W, int Cin, int Cout, int K, int* out) {
  TcGeom g;
  if (!out || !tc_geometry(H, W, Cin, Cout, K, &g)) return SAN_ERR_UNSUPPORTED;
  const int v[16] = {g.Cin_pad, g.KG, g.KS, g.nsplit, g.Npad, g.Wp, g.Hp, g.R, g.T, g.S_alloc, g.strips, g.stages,
                     g.acc_stages, g.a_bytes, g.b_bytes, g.smem_bytes};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return SAN_OK;
}

// Host-only: which formulation the kernel uses.  out[0..6] = dxn (1: horizontal taps in the MMA N dimension), Np (Cout padded
// to 8 in that form), wtaps (weight blocks per K-step: 9, 3 or 1), xchg_bytes (epilogue exchange area), hls (1: W_hi / W_lo
// stacked along N), Ncol (TMEM columns per 128-pixel tile), pair (1: taps paired in the half-empty last K-step).
int san_tc_describe_form(int H, int W, int Cin, int Cout, int K, int* out) {
  TcGeom g;
  if (!out || !tc_geometry(H, W, Cin, Cout, K, &g)) return SAN_ERR_UNSUPPORTED;
  out[0] = g.dxn; out[1] = g.Np; out[2] = g.wtaps; out[3] = g.xchg_bytes; out[4] = g.hls; out[5] = g.Ncol; out[6] = g.pair;
  return SAN_OK;
}

int san_tc_supported(int H, int W, int Cin, int Cout, int K) {
  TcGeom g;
  return tc_geometry(H, W, Cin, Cout, K, &g) ? 1 : 0;
}

static int launch_stage(StageArgs& A, void* xs, int N, int H, int W, int Cpad, int fmt, const float* absmax,
                        cudaStream_t st) {
  SAN_CHECK_ARG(fmt == 0 || fmt == 1, "san_tc_stage: fmt must be 0 (bf16 pairs) or 1 (fp16 pairs)");
  SAN_CHECK_ARG(!absmax || fmt == 1, "san_tc_stage: the dynamic scale applies to fp16 pairs only");
  A.fmt = fmt; A.absmax = absmax;
  int ctot = 0;
  for (int i = 0; i < A.nterms; ++i) {
    SAN_CHECK_ARG(A.s[i].y && A.s[i].C > 0, "san_tc_stage: term %d has no tensor / channels", i);
    SAN_CHECK_ARG(A.s[i].mode >= 0 && A.s[i].mode <= 3, "san_tc_stage: bad mode");
    if (A.s[i].mode >= 2) SAN_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "san_tc_stage: odd size with up-sampling source");
    if (A.s[i].c0 + A.s[i].C > ctot) ctot = A.s[i].c0 + A.s[i].C;
  }
  SAN_CHECK_ARG(ctot <= Cpad && Cpad - ctot < 8, "san_tc_stage: Cpad %d must be the %d channels padded to 8 (the staged tensor holds"
                " ceil(C/8) channel groups, no all-zero groups)", Cpad, ctot);
  A.xs = (__nv_bfloat16*)xs + TC_LEAD;
  A.N = N; A.H = H; A.W = W; A.Wp = W + 2; A.PS = (H + 2) * (W + 2); A.KG = Cpad / 8;
  SAN_CHECK_ARG((long long)N * A.KG <= 65535, "san_tc_stage: N*KG too large for grid.y");
  // ~4 pixel slots per thread: the per-block work-list setup is amortised over 1024 slots and there are
  // enough blocks for dozens of waves (no tail effect)
  int bx = (A.PS + 1023) / 1024;
  if (bx < 1) bx = 1;
  bool simple = tc_env_int("SAN_STAGE_SIMPLE", 1) != 0;        // 0: always the general kernel (A/B runs)
  for (int i = 0; i < A.nterms && simple; ++i) {
    if (A.s[i].mode != 0) simple = false;
    for (int k = 0; k < i; ++k)                                  // a channel range shared by two terms = a residual sum
      if (A.s[k].c0 < A.s[i].c0 + A.s[i].C && A.s[i].c0 < A.s[k].c0 + A.s[k].C) simple = false;
  }
  if (simple) stage_simple_kernel<<<dim3(bx, N * A.KG), 256, 0, st>>>(A);
  else stage_act_kernel<<<dim3(bx, N * A.KG), 256, 0, st>>>(A);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_tc_stage_terms(void* xs, int N, int H, int W, int Cpad, const san_stage_term* terms, int nterms, int fmt,
                       const float* absmax, void* stream) {
  SAN_CHECK_ARG(xs && terms && nterms >= 1 && nterms <= STAGE_MAX_TERMS && N > 0 && H > 0 && W > 0 && Cpad % 8 == 0,
                "san_tc_stage_terms: bad args");
  StageArgs A{};
  A.nterms = nterms;
  int c_next = 0, c_prev = 0;
  for (int i = 0; i < nterms; ++i) {
    const san_stage_term& t = terms[i];
    SAN_CHECK_ARG(!(t.accumulate && i == 0), "san_tc_stage_terms: first term cannot accumulate");
    const int c0 = t.accumulate ? c_prev : c_next;
    A.s[i] = StageTerm{t.y, t.mu, t.a, t.b, t.slope, t.C, t.mode, c0};
    if (t.accumulate) SAN_CHECK_ARG(t.C == A.s[i - 1].C, "san_tc_stage_terms: accumulated term must match the channel count");
    c_prev = c0;
    if (!t.accumulate) c_next = c0 + t.C;
  }
  return launch_stage(A, xs, N, H, W, Cpad, fmt, absmax, (cudaStream_t)stream);
}

int san_tc_stage_act(void* xs, int N, int H, int W, int Cpad,
                     const float* y0, const float* mu0, const float* a0, const float* b0, float slope0, int C0, int mode0,
                     const float* y1, const float* mu1, const float* a1, const float* b1, float slope1, int C1, int mode1,
                     const float* y2, const float* mu2, const float* a2, const float* b2, float slope2, int C2, int mode2,
                     int fmt, void* stream) {
  SAN_CHECK_ARG(xs && y0 && N > 0 && H > 0 && W > 0 && Cpad % 8 == 0 && C0 > 0, "san_tc_stage_act: bad args");
  StageArgs A{};
  A.s[0] = StageTerm{y0, mu0, a0, b0, slope0, C0, mode0, 0};
  A.nterms = 1;
  if (y1) { A.s[1] = StageTerm{y1, mu1, a1, b1, slope1, C1, mode1, C0}; A.nterms = 2; }
  if (y2) { SAN_CHECK_ARG(y1, "san_tc_stage_act: source 2 without source 1"); A.s[2] = StageTerm{y2, mu2, a2, b2, slope2, C2, mode2, C0 + C1}; A.nterms = 3; }
  return launch_stage(A, xs, N, H, W, Cpad, fmt, nullptr, (cudaStream_t)stream);
}

int san_tc_unstage_act(const void* xs, float* x, int N, int C, int H, int W, int fmt, void* stream) {
  SAN_CHECK_ARG(xs && x && N > 0 && C > 0 && H > 0 && W > 0, "san_tc_unstage_act: bad args");
  const long long total = (long long)N * C * H * W;
  unstage_act_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)xs + TC_LEAD, x, N, C, H, W,
                                                                         (C + 7) / 8, fmt);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_tc_stage_weights(const float* w, void* ws, int H, int W, int Cout, int Cin, int K, int dgrad, int fmt,
                         void* stream) {
  SAN_CHECK_ARG(w && ws && Cout > 0 && Cin > 0 && (K == 1 || K == 3), "san_tc_stage_weights: bad args");
  TcGeom g;
  const int Co_k = dgrad ? Cin : Cout, Ci_k = dgrad ? Cout : Cin;
  SAN_CHECK_ARG(tc_geometry(H, W, Ci_k, Co_k, K, &g), "san_tc_stage_weights: unsupported shape");
  const long long total = (long long)g.nsplit * g.KS * g.wtaps * 4 * g.Npad * 8;
  stage_weights_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)ws, Cout, Cin, K * K, dgrad,
                                                                           g.nsplit, g.KS, g.Npad, fmt, g.dxn ? g.Np : 0, g.hls, g.pair);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

// shared-memory bytes of the statistics epilogue's accumulators: one [Npad][2] fp32 table per epilogue warp
static int tc_stat_bytes(const TcGeom& g) { return TC_EPI_WARPS * 2 * g.Npad * (int)sizeof(float); }

// The accumulators go behind the pipeline stages; where the stages fill the shared memory, one of them (of more than two)
// is given up.  Returns the number of stages to run with, 0 = no room.
static int tc_stat_stages(const TcGeom& g) {
  for (int stages = g.stages; stages >= 2 && stages >= g.stages - 1; --stages)
    if (TC_SMEM_HEADER + g.xchg_bytes + stages * g.stage_bytes + tc_stat_bytes(g) <= TC_SMEM_MAX) return stages;
  return 0;
}

int san_tc_conv_stats_supported(int H, int W, int Cin, int Cout, int K) {
  TcGeom g;
  if (!tc_geometry(H, W, Cin, Cout, K, &g) || g.dxn) return 0;
  return tc_stat_stages(g) ? 1 : 0;
}

int san_tc_conv_stats(const void* xs, const void* ws, const float* bias, float* y, int N, int H, int W, int Cin, int Cout,
                      int K, long long y_bs, int fmt, const float* a_absmax, double* sums, void* stream) {
  SAN_CHECK_ARG(xs && ws && y && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (fmt == 0 || fmt == 3),
                "san_tc_conv: bad args (fmt: 0 = bf16 pairs, 3 = fp16 pairs; the operands of an MMA share the format)");
  SAN_CHECK_ARG(!a_absmax || fmt == 3, "san_tc_conv: the dynamic scale applies to fp16 pairs only");
  ConvTcParams p{};
  p.fmt = fmt; p.a_absmax = a_absmax;
  if (a_absmax) fmt &= ~TC_FMT_A_F16;   // the static activation scale is replaced by the dynamic one (undone in the epilogue)
  p.out_scale = ((fmt & TC_FMT_A_F16) ? 1.f / TC_SX : 1.f) * ((fmt & TC_FMT_B_F16) ? 1.f / TC_SW : 1.f);
  SAN_CHECK_ARG(tc_geometry(H, W, Cin, Cout, K, &p.g), "san_tc_conv: unsupported shape H=%d W=%d Cin=%d Cout=%d K=%d", H, W,
                Cin, Cout, K);
  p.xs = (const __nv_bfloat16*)xs + TC_LEAD; p.ws = (const __nv_bfloat16*)ws; p.bias = bias; p.y = y;
  p.y_bs = y_bs > 0 ? y_bs : (long long)Cout * H * W;
  p.N = N; p.H = H; p.W = W; p.Cout = Cout; p.ntaps = K * K;
  p.nunits = N * p.g.strips * p.g.nsplit;
  cudaStream_t st = (cudaStream_t)stream;
  int smem_bytes = p.g.smem_bytes;
  p.sums = sums;
  static const int contig_env = tc_env_int("SAN_TC_CONTIG", -1);     // -1: contiguous unit ranges with the statistics epilogue only
  p.contig = contig_env < 0 ? (sums ? 1 : 0) : (contig_env ? 1 : 0);
  if (sums) {
    const int stages = p.g.dxn ? 0 : tc_stat_stages(p.g);
    SAN_CHECK_ARG(stages, "san_tc_conv_stats: no statistics epilogue for this shape");
    p.g.stages = stages;
    p.stat_off = TC_SMEM_HEADER + p.g.xchg_bytes + p.g.stages * p.g.stage_bytes;
    if (p.stat_off + tc_stat_bytes(p.g) > smem_bytes) smem_bytes = p.stat_off + tc_stat_bytes(p.g);
    SAN_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)N * Cout, st));
  }
  static bool attr_set = false;
  if (!attr_set) {
    SAN_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
    attr_set = true;
  }
  const int grid = p.nunits < san_num_sms() ? p.nunits : san_num_sms();
  conv_tc_kernel<<<grid, TC_THREADS, smem_bytes, st>>>(p);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_tc_conv(const void* xs, const void* ws, const float* bias, float* y, int N, int H, int W, int Cin, int Cout,
                int K, long long y_bs, int fmt, const float* a_absmax, void* stream) {
  return san_tc_conv_stats(xs, ws, bias, y, N, H, W, Cin, Cout, K, y_bs, fmt, a_absmax, nullptr, stream);
}

// ---- row-ring conv (conv_rows_kernel): the conv reads RAW fp32 rows and stages them itself -------------------------------
int san_tc_conv_rows_supported(int H, int W, int Cin, int Cout, int K) {
  RrGeom g;
  return rr_geometry(H, W, Cin, Cout, K, true, &g) ? 1 : 0;
}

// Host-only: the row-ring geometry.  out[0..11] = KG, KS, Npad, Ncol, Wp, T, RS, pair, NR, row_bytes, w_bytes, smem_bytes.
int san_tc_conv_rows_describe(int H, int W, int Cin, int Cout, int stats, int* out) {
  RrGeom g;
  if (!out || !rr_geometry(H, W, Cin, Cout, 3, stats != 0, &g)) return SAN_ERR_UNSUPPORTED;
  const int v[12] = {g.KG, g.KS, g.Npad, g.Ncol, g.Wp, g.T, g.RS, g.pair, g.NR, g.row_bytes, g.w_bytes, g.smem_bytes};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
  return SAN_OK;
}

long long san_tc_rows_weight_elems(int H, int W, int Cout, int Cin) {
  RrGeom g;
  if (!rr_geometry(H, W, Cin, Cout, 3, false, &g)) return -1;
  return (long long)g.w_bytes / 2;
}

int san_tc_stage_weights_rows(const float* w, void* ws, int H, int W, int Cout, int Cin, int dgrad, int fmt, void* stream) {
  SAN_CHECK_ARG(w && ws && Cout > 0 && Cin > 0 && fmt == 1, "san_tc_stage_weights_rows: bad args (fp16 pairs only)");
  RrGeom g;
  const int Co_k = dgrad ? Cin : Cout, Ci_k = dgrad ? Cout : Cin;
  SAN_CHECK_ARG(rr_geometry(H, W, Ci_k, Co_k, 3, false, &g), "san_tc_stage_weights_rows: unsupported shape");
  const long long total = (long long)g.KS * 9 * 4 * g.Npad * 8;
  stage_weights_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)ws, Cout, Cin, 9, dgrad, 1, g.KS,
                                                                           g.Npad, fmt, 0, 1, g.pair);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

int san_tc_conv_rows(const float* x, const float* mu, const float* a, const float* b, float slope, const float* absmax,
                     void* xs_out, const void* ws, const float* bias, float* y, double* sums, int N, int H, int W, int Cin,
                     int Cout, void* stream) {
  SAN_CHECK_ARG(x && ws && y && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "san_tc_conv_rows: bad args");
  SAN_CHECK_ARG(!(absmax && a), "san_tc_conv_rows: the dynamic scale is for un-normalised (gradient) operands");
  RrParams p{};
  SAN_CHECK_ARG(rr_geometry(H, W, Cin, Cout, 3, sums != nullptr, &p.g), "san_tc_conv_rows: unsupported shape H=%d W=%d Cin=%d Cout=%d",
                H, W, Cin, Cout);
  cudaStream_t st = (cudaStream_t)stream;
  p.x = x; p.mu = mu; p.a = a; p.b = b; p.slope = slope; p.absmax = absmax;
  p.xs_out = xs_out ? (__nv_bfloat16*)xs_out + TC_LEAD : nullptr;
  p.ws = (const __nv_bfloat16*)ws; p.bias = bias; p.y = y; p.sums = sums;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.nunits = N * H;
  p.out_scale = (absmax ? 1.f : 1.f / TC_SX) / TC_SW;
  if (sums) SAN_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)N * Cout, st));
  if (xs_out) {     // lead-in / trailing slack of the staged buffer (TC_LEAD / TC_TRAIL): zero, like the staging kernels leave them
    SAN_CUDA(cudaMemsetAsync(xs_out, 0, TC_LEAD * sizeof(__nv_bfloat16), st));
    SAN_CUDA(cudaMemsetAsync((__nv_bfloat16*)xs_out + TC_LEAD + (long long)N * 2 * p.g.KG * p.g.PS * 8, 0,
                             TC_TRAIL * sizeof(__nv_bfloat16), st));
  }
  static bool attr_set = false;
  if (!attr_set) {
    SAN_CUDA(cudaFuncSetAttribute(conv_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
    attr_set = true;
  }
  const int grid = p.nunits < san_num_sms() ? p.nunits : san_num_sms();
  conv_rows_kernel<<<grid, RR_THREADS, p.g.smem_bytes, st>>>(p);
  SAN_LAUNCH_CHECK();
  return SAN_OK;
}

}  // extern "C"
