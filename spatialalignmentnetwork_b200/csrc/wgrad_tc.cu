// tcgen05 weight-gradient kernel for the U-Net convolutions (sm_100a only):
//     dW[co][ci][dy][dx] = sum_{n, pixel p} dY[n][co][p] * X[n][ci][p + (dy-1)*Wp + (dx-1)]
// (the backward-filter of the reference's cuDNN convs, varnet.py:139-146,176-179, unet.py:119-140).
//
// Both operands are the SAME staged BF16 hi/lo tensors the forward / data-gradient kernels use
// (conv_tc.cu: Xs[n][hl][kg][slot][8], zero border), read here as MN-major UMMA operands: the GEMM's
// K dimension is the pixel slot (16 B row pitch, 8-slot groups 128 B apart = LBO), its M (dY) / N (X)
// dimension is the channel (8 contiguous, channel groups one staged plane apart = SBO).  A filter tap
// is again only a shifted start address of the X tile.  Border pixels of dY are zero, so the padding
// columns and the zero rows between images contribute nothing and need no masking.
//
// Decomposition: group = (block of <= 128 output channels, one of <= 16 chunks of <= 160 input channels, filter
// row dy); the CTAs of a group split the pixel range of all images into chunks of KC slots and keep
// three [128 x Nn] fp32 accumulators (dx = 0, 1, 2) in TMEM for the whole kernel; one atomicAdd pass
// per CTA at the end.  BF16x3 as in the forward: hi*hi + lo*hi + hi*lo.
//
// Narrow layers (<= 64 padded output channels) would leave most of the 128 UMMA rows empty while the
// MMA still reads a full A tile from shared memory (the A read, not the tensor pipe, bounds N <= 64
// MMAs).  There the hi and lo halves of dY - adjacent in the smem tile - are used as ONE stacked A
// operand [hi; lo]: D = [hi; lo] * X_hi + [hi; lo] * X_lo gives all four partial products in 2 MMAs
// (rows co and co + Cpad are both added into dW[co]); with <= 32 padded channels the stack is 64 rows
// and the M = 64 UMMA shape halves the A read again.
// "Filter rows in N" form (3x3, <= 48 padded input channels: the full- and half-resolution layers): an MMA with
// N = 32 keeps the tensor pipe busy for 16 clk but needs 24 clk of operand reads, and the three filter rows used to be
// three CTA groups that each re-loaded the dY chunk.  Here the X span is loaded THREE times, shifted by one image
// row each ((dy-1)*Wp slots), as separate shared-memory planes [hl][dy][kg]: the channel-group stride of the B
// operand stays uniform, so ONE MMA covers the three filter rows (N = 3*Nn = 96: tensor-bound), dY is loaded once,
// and a CTA accumulates all nine taps ([3 dx] x [M x 3 Nn] fp32 in TMEM).
//   warp 0: TMA producer (bulk copies of the dY chunk and the X span per stage)
//   warp 1: TMEM allocator + single-thread MMA issuer
//   warps 2..5: final epilogue (tcgen05.ld -> atomicAdd into dW[Cout][Cin][K][K])
#include "tc_common.cuh"
#include "../../include/san_b200.h"

namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_HEADER = 256;
constexpr int WG_SMEM_LIMIT = 225 * 1024;
constexpr int WG_KC_MAX = 512;   // pixel slots per stage (measured: fewer, larger stages win; per-stage TMA issue + barrier cost)

struct WgGeom {
  int KGo, KGi;            // staged channel groups of dY / X
  int nmb, nnc, ndy;       // M blocks (128 co), N chunks, filter rows
  int Nn, KGn;             // input channels per chunk (multiple of 16), its groups
  int KC, XS;              // pixel slots per chunk (multiple of 16), X slots per stage (KC + 16)
  int a_bytes, b_bytes, stage_bytes, stages, smem_bytes;
  int nchunks;             // pixel chunks per image
  int Wp, PS, range0, range_len;
  int rown;                // 1: the three filter rows sit in the MMA N dimension (three row-shifted X copies per stage)
  int ncp;                 // X copies per stage (3 in that form, else 1)
};

int wg_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// Shared memory the kernel may take per CTA (SAN_WG_SMEM_KB: tuning runs).  The weight gradient runs next to the
// element-wise backward kernels of the same layer (tc.py: side stream), whose CTAs need ~2 KB of shared memory each on the
// same SM; measured (profiles/r2o_*): leaving 16 or 48 KB free does not help (385.1 ms/step at 225 KB, 387.2 at 209,
// 393.2 at 177) - the geometries that carry the step do not fill 225 KB anyway.
int wg_smem_max() {
  static const int v = wg_env_int("SAN_WG_SMEM_KB", 225);
  const int b = v * 1024;
  return b > WG_SMEM_LIMIT ? WG_SMEM_LIMIT : (b < 64 * 1024 ? 64 * 1024 : b);
}
int wg_rown_enabled() { static const int v = wg_env_int("SAN_WG_ROWN", 1); return v; }   // 0: A/B runs against the per-row form

bool wg_geometry(int H, int W, int Cin, int Cout, int K, WgGeom* g) {
  if (K != 1 && K != 3) return false;
  const int WG_SMEM_MAX = wg_smem_max();
  if (W + 2 < 18) return false;   // the 16-slot tail padding must stay inside the zero border row
  g->KGo = (Cout + 7) / 8; g->KGi = (Cin + 7) / 8;   // REAL channel groups = planes of the staged tensors (no all-zero groups)
  g->nmb = (g->KGo + 15) / 16;
  const int cip = pad16(Cin);
  g->nnc = 0;
  for (int k = 1; k <= 16; ++k)
    if (cip % k == 0 && (cip / k) % 16 == 0 && cip / k <= 160) { g->nnc = k; break; }
  if (!g->nnc) return false;
  g->Nn = cip / g->nnc; g->KGn = g->Nn / 8;
  // narrow layers (one chunk, M = 64 shape: N may be any multiple of 8): the MMA N covers the real groups only
  // (18 input channels: N = 24 per filter row instead of 32)
  if (g->nnc == 1 && g->nmb == 1 && 2 * g->KGo <= 8) { g->KGn = g->KGi; g->Nn = 8 * g->KGi; }
  g->rown = (K == 3 && 9 * g->Nn <= 512 && wg_rown_enabled()) ? 1 : 0;     // 3 dx accumulators of [M x 3 Nn] in 512 TMEM columns
  g->ncp = g->rown ? 3 : 1;
  g->ndy = g->rown ? 1 : K;
  g->Wp = W + 2; g->PS = (H + 2) * g->Wp;
  g->range0 = g->Wp;                                  // first slot of image row 0 (padded row 1)
  g->range_len = (H * g->Wp + 15) / 16 * 16;          // tail runs into the zero border row
  const int kga = g->KGo < 16 ? g->KGo : 16;          // groups actually loaded per M block (max)
  // The UMMA reads M/8 channel groups from the A start (and again from the lo half when hi/lo are not
  // stacked) whether or not they were loaded: the over-read must stay inside the allocation.
  const bool stacked = 2 * kga <= 16;
  const int mg = (2 * kga <= 8) ? 8 : 16;
  const int reach_groups = (stacked ? 0 : kga) + mg;  // groups spanned from the A start of a stage
  int best = 0, best_slack = 0;
  static const int kc_max = wg_env_int("SAN_WG_KC_MAX", WG_KC_MAX);   // tuning runs
  for (int kc = kc_max; kc >= 32; kc -= 16) {
    const int a = kga * 2 * kc * 16, b = g->ncp * g->KGn * 2 * (kc + 16) * 16;
    int slack = reach_groups * kc * 16 - (a + b);
    if (slack < 0) slack = 0;
    if (2 * (a + b) + slack + WG_HEADER <= WG_SMEM_MAX) { best = kc; best_slack = slack; break; }
  }
  if (!best) return false;
  g->KC = best; g->XS = best + 16;
  g->a_bytes = kga * 2 * g->KC * 16;
  g->b_bytes = g->ncp * g->KGn * 2 * g->XS * 16;
  g->stage_bytes = g->a_bytes + g->b_bytes;
  g->stages = (WG_SMEM_MAX - WG_HEADER - best_slack) / g->stage_bytes;
  if (g->stages > 4) g->stages = 4;
  g->smem_bytes = WG_HEADER + g->stages * g->stage_bytes + best_slack;
  if (g->smem_bytes < 120 * 1024) g->smem_bytes = 120 * 1024;
  g->nchunks = (g->range_len + g->KC - 1) / g->KC;
  return true;
}

struct WgParams {
  const __nv_bfloat16* dys;   // staged dY (past the lead-in)
  const __nv_bfloat16* xs;    // staged X
  float* dw;                  // [Cout][Cin][K][K], pre-zeroed
  int N, Cin, Cout, K;
  int ngroups, ctas_per_group;
  int fmt;                    // TC_FMT_* bits: A = dY pair format, B = X pair format
  float out_scale;
  const float* a_absmax;      // null, or device scalar max|dY| when dY was staged with the dynamic scale
  WgGeom g;
};

template <int NDX, bool STACKED>
__device__ __forceinline__ void wg_issue(uint32_t a_s, uint32_t b_s, uint32_t a_hiw, uint32_t b_hiw, uint32_t a_losplit,
                                         uint32_t b_losplit, uint32_t tmem_base, int Nn, uint32_t idesc, int kc,
                                         uint32_t started) {
  uint32_t st = started;
  for (int k = 0; k < kc; k += 16) {
    const uint32_t al = a_s + k;                          // 16 slots = 16 units of 16 B
    const uint64_t A_hi = ((uint64_t)a_hiw << 32) | al, A_lo = ((uint64_t)a_hiw << 32) | (al + a_losplit);
#pragma unroll
    for (int dx = 0; dx < NDX; ++dx) {
      const uint32_t bl = b_s + k + dx;
      const uint64_t B_hi = ((uint64_t)b_hiw << 32) | bl, B_lo = ((uint64_t)b_hiw << 32) | (bl + b_losplit);
      const uint32_t d = tmem_base + (uint32_t)(dx * Nn);
      tc_mma_bf16(d, A_hi, B_hi, idesc, st);
      if (!STACKED) tc_mma_bf16(d, A_lo, B_hi, idesc, 1u);
      tc_mma_bf16(d, A_hi, B_lo, idesc, 1u);
    }
    st = 1;
  }
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const WgParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const WgGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t hdr = smem_u32(smem);
  const uint32_t bar_full = hdr, bar_empty = hdr + 32, bar_done = hdr + 64;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + 96);
  const uint32_t stage0 = hdr + WG_HEADER;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((void*)tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform by construction

  // group of this CTA
  const int grp = blockIdx.x / p.ctas_per_group, cta = blockIdx.x - grp * p.ctas_per_group;
  const int dyi = grp % g.ndy;
  const int nc = (grp / g.ndy) % g.nnc;
  const int mb = grp / (g.ndy * g.nnc);
  const int kga = min(16, g.KGo - 16 * mb);          // channel groups of dY loaded for this M block
  const int ndx = p.K;                               // 3 taps per filter row (1 for 1x1)
  const int ncp = g.ncp;                             // row-shifted X copies per stage (3: filter rows in N)
  const int Nmma = ncp * g.Nn;                       // MMA N: input channels of the chunk x filter rows
  const int nunits = p.N * g.nchunks;
  const long long plane = (long long)g.PS * 8;
  // slot offset of the X span relative to the dY chunk start: (dy-1)*Wp - 1 for 3x3, 0 for 1x1
  const int xoff = (p.K == 3) ? (dyi - 1) * g.Wp - 1 : 0;

  if (warp == 0) {
    // TMA producer: the whole warp walks the units; lane 0 arms the barrier, then the lanes issue the bulk
    // copies of the stage in parallel (one thread issuing 16-70 copies per stage with their address
    // arithmetic was a serial bottleneck)
    int s = 0;
    uint32_t ph = 0;
    const int per_hl = kga + ncp * g.KGn;
    const int ncopy = 2 * per_hl;
    const int kgl = min(g.KGn, g.KGi - nc * g.KGn);      // real groups of this chunk (a padding group has no staged plane:
                                                         // its shared-memory plane keeps stale data, its dW columns are dropped)
    for (int u = cta; u < nunits; u += p.ctas_per_group) {
      const int n = u / g.nchunks, ch = u - n * g.nchunks;
      const int p0 = g.range0 + ch * g.KC;
      const int kc = min(g.KC, g.range_len - ch * g.KC);      // multiple of 16
      const uint32_t bytesA = (uint32_t)kc * 16, bytesB = (uint32_t)(kc + 2) * 16;   // X span: kc + 2 tap slots
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint32_t sbase = stage0 + (uint32_t)s * g.stage_bytes;
      if (lane == 0) mbar_expect_tx(bar_full + 8 * s, 2u * kga * bytesA + 2u * ncp * kgl * bytesB);
      __syncwarp();
      for (int c = lane; c < ncopy; c += 32) {
        const int hl = c / per_hl, r = c - hl * per_hl;
        if (r < kga) {
          const __nv_bfloat16* src = p.dys + ((long long)(n * 2 + hl) * g.KGo + 16 * mb + r) * plane + (long long)p0 * 8;
          bulk_g2s(sbase + (uint32_t)((hl * kga + r) * g.KC) * 16, src, bytesA, bar_full + 8 * s);
        } else {
          const int rr = r - kga;
          const int cp = rr / g.KGn, kg = rr - cp * g.KGn;        // cp: row-shifted copy = filter row (rown form)
          if (kg >= kgl) continue;
          const int xo = g.rown ? (cp - 1) * g.Wp - 1 : xoff;
          const __nv_bfloat16* src =
              p.xs + ((long long)(n * 2 + hl) * g.KGi + nc * g.KGn + kg) * plane + (long long)(p0 + xo) * 8;
          bulk_g2s(sbase + g.a_bytes + (uint32_t)(((hl * ncp + cp) * g.KGn + kg) * g.XS) * 16, src, bytesB, bar_full + 8 * s);
        }
      }
      __syncwarp();
      if (++s == g.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // whole warp runs the (warp-uniform) loops; one elected lane issues the tcgen05 instructions
    const bool stacked = 2 * kga <= 16;                // [hi; lo] of dY as one A operand
    const int M = (2 * kga <= 8) ? 64 : 128;
    const uint32_t idesc = umma_idesc_16(M, Nmma, p.fmt, /*mn_major=*/1);
    // MN-major, no swizzle: LBO = 128 B between 8-slot K groups, SBO = one staged plane between channel groups
    const uint64_t a_tmpl = umma_desc(0, 128, (uint32_t)g.KC * 16);
    const uint64_t b_tmpl = umma_desc(0, 128, (uint32_t)g.XS * 16);
    const uint32_t a_hiw = (uint32_t)(a_tmpl >> 32), b_hiw = (uint32_t)(b_tmpl >> 32);
    const uint32_t a_low = (uint32_t)a_tmpl, b_low = (uint32_t)b_tmpl;
    const uint32_t a_losplit = (uint32_t)(kga * g.KC);        // hi -> lo half, 16 B units
    const uint32_t b_losplit = (uint32_t)(ncp * g.KGn * g.XS);
    int s = 0;
    uint32_t ph = 0;
    uint32_t started = 0;
    for (int u = cta; u < nunits; u += p.ctas_per_group) {
      const int ch = u % g.nchunks;
      const int kc = min(g.KC, g.range_len - ch * g.KC);
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      const uint32_t a_s = ((stage0 + (uint32_t)s * g.stage_bytes) >> 4) + a_low;
      const uint32_t b_s = ((stage0 + (uint32_t)s * g.stage_bytes + g.a_bytes) >> 4) + b_low;
      if (elect_one_sync()) {
        // the issue loop is specialised on (taps per row, stacked) so that it is branch-free
        if (ndx == 3) {
          if (stacked) wg_issue<3, true>(a_s, b_s, a_hiw, b_hiw, a_losplit, b_losplit, tmem_base, Nmma, idesc, kc, started);
          else wg_issue<3, false>(a_s, b_s, a_hiw, b_hiw, a_losplit, b_losplit, tmem_base, Nmma, idesc, kc, started);
        } else {
          if (stacked) wg_issue<1, true>(a_s, b_s, a_hiw, b_hiw, a_losplit, b_losplit, tmem_base, Nmma, idesc, kc, started);
          else wg_issue<1, false>(a_s, b_s, a_hiw, b_hiw, a_losplit, b_losplit, tmem_base, Nmma, idesc, kc, started);
        }
        tc_commit(bar_empty + 8 * s);
        if (u + p.ctas_per_group >= nunits) tc_commit(bar_done);
      }
      __syncwarp();
      started = 1;
      if (++s == g.stages) { s = 0; ph ^= 1; }
    }
  } else {
    // ---- final epilogue: D[dx][co][ci] -> atomicAdd dW[co][ci][dy][dx]
    const int wq = warp & 3;
    if (cta < nunits) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      // accumulator row of this lane: M = 128 -> row = TMEM lane; M = 64 -> rows 16*w + (lane < 16)
      // (cute::UMMA tmem_frg "half subpartition" atom).  Stacked rows r and r + 8*kga both belong to
      // output channel r.
      const bool stacked = 2 * kga <= 16;
      const bool m64 = 2 * kga <= 8;
      int row = m64 ? (lane < 16 ? wq * 16 + lane : -1) : wq * 32 + lane;
      if (stacked && row >= 8 * kga) row = (row < 16 * kga) ? row - 8 * kga : -1;
      const int co = row < 0 ? p.Cout : mb * 128 + row;
      const int KK = p.K * p.K;
      const float oscale = p.a_absmax ? p.out_scale / tc_dyn_scale(__ldg(p.a_absmax)) : p.out_scale;
      for (int dx = 0; dx < ndx; ++dx) {
        for (int c0 = 0; c0 < Nmma; c0 += 8) {       // column = (filter row, input channel) in the rown form (Nn % 8 == 0)
          float v[8];
          tc_ld8(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(dx * Nmma + c0), v);
          const int dye = g.rown ? c0 / g.Nn : dyi;
          const int tap = (p.K == 3) ? dye * 3 + dx : 0;
          const int cb = nc * g.Nn + c0 - (g.rown ? dye * g.Nn : 0);
          if (co < p.Cout) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int ci = cb + j;
              if (ci < p.Cin) atomicAdd(p.dw + ((long long)co * p.Cin + ci) * KK + tap, v[j] * oscale);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// dbias[c] = sum over n, pixels of dy[n][c][:]   (grid: (C, N, chunks); float4 loads, 4 independent partial sums per thread)
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, int C, int P) {
  __shared__ float red[32];
  const float* src = dy + ((long long)blockIdx.y * C + blockIdx.x) * P;
  const int per = (P + gridDim.z - 1) / gridDim.z;
  const int i0 = blockIdx.z * per, i1 = min(P, i0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if ((P & 3) == 0 && (per & 3) == 0) {
    const float4* s4 = (const float4*)src;
    for (int i = (i0 >> 2) + threadIdx.x; i < (i1 >> 2); i += blockDim.x) {
      const float4 v = __ldg(s4 + i);
      s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    }
  } else {
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) s0 += __ldg(src + i);
  }
  const float s = block_sum((s0 + s1) + (s2 + s3), red);
  if (threadIdx.x == 0) atomicAdd(db + blockIdx.x, s);
}

}  // namespace

extern "C" {

// Host-only: the decomposition san_tc_wgrad would use.  out[0..15] = KGo, KGi, nmb, nnc, ndy, Nn, KGn, KC, XS, stages,
// smem_bytes, nchunks, Wp, PS, range0, range_len.
int san_tc_wgrad_describe(int H, int W, int Cin, int Cout, int K, int* out) {
  WgGeom g;
  if (!out || !wg_geometry(H, W, Cin, Cout, K, &g)) return SAN_ERR_UNSUPPORTED;
  const int v[16] = {g.KGo, g.KGi, g.nmb, g.nnc, g.ndy, g.Nn, g.KGn, g.KC, g.XS, g.stages, g.smem_bytes, g.nchunks,
                     g.Wp, g.PS, g.range0, g.range_len};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return SAN_OK;
}

// Host-only: out[0..1] = rown (1: the three filter rows sit in the MMA N dimension), ncp (row-shifted X copies per stage)
int san_tc_wgrad_describe_form(int H, int W, int Cin, int Cout, int K, int* out) {
  WgGeom g;
  if (!out || !wg_geometry(H, W, Cin, Cout, K, &g)) return SAN_ERR_UNSUPPORTED;
  out[0] = g.rown; out[1] = g.ncp;
  return SAN_OK;
}

int san_tc_wgrad_supported(int H, int W, int Cin, int Cout, int K) {
  WgGeom g;
  return wg_geometry(H, W, Cin, Cout, K, &g) ? 1 : 0;
}

int san_tc_wgrad(const void* dys, const void* xs, float* dw, float* dbias, const float* dy, int N, int H, int W, int Cin,
                 int Cout, int K, int fmt, const float* dy_absmax, void* stream) {
  SAN_CHECK_ARG(dys && xs && dw && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (fmt == 0 || fmt == 3),
                "san_tc_wgrad: bad args (fmt: 0 = bf16 pairs, 3 = fp16 pairs; the operands of an MMA share the format)");
  SAN_CHECK_ARG(!dy_absmax || fmt == 3, "san_tc_wgrad: the dynamic scale applies to fp16 pairs only");
  WgParams p{};
  p.fmt = fmt; p.a_absmax = dy_absmax;
  // A = staged dY, B = staged X; both were staged as ACTIVATIONS: static scale TC_SX, or the dynamic one for dY
  p.out_scale = ((fmt & TC_FMT_A_F16) && !dy_absmax ? 1.f / TC_SX : 1.f) * ((fmt & TC_FMT_B_F16) ? 1.f / TC_SX : 1.f);
  SAN_CHECK_ARG(wg_geometry(H, W, Cin, Cout, K, &p.g), "san_tc_wgrad: unsupported shape H=%d W=%d Cin=%d Cout=%d K=%d", H, W,
                Cin, Cout, K);
  cudaStream_t st = (cudaStream_t)stream;
  p.dys = (const __nv_bfloat16*)dys + TC_LEAD; p.xs = (const __nv_bfloat16*)xs + TC_LEAD; p.dw = dw;
  p.N = N; p.Cin = Cin; p.Cout = Cout; p.K = K;
  p.ngroups = p.g.nmb * p.g.nnc * p.g.ndy;
  const int nunits = N * p.g.nchunks;
  int cpg = san_num_sms() / p.ngroups;
  if (cpg < 1) cpg = 1;
  if (cpg > nunits) cpg = nunits;
  p.ctas_per_group = cpg;
  SAN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * K * K, st));
  static bool attr_set = false;
  if (!attr_set) {
    SAN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_LIMIT));
    attr_set = true;
  }
  wgrad_tc_kernel<<<p.ngroups * cpg, WG_THREADS, p.g.smem_bytes, st>>>(p);
  SAN_LAUNCH_CHECK();
  if (dbias) {
    SAN_CHECK_ARG(dy, "san_tc_wgrad: dbias needs the fp32 dy");
    SAN_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)Cout, st));
    {
      const int P = H * W;
      int chunks = P / 8192;            // ~8 K elements (32 per thread) per block
      chunks = chunks < 1 ? 1 : (chunks > 16 ? 16 : chunks);
      bias_grad_kernel<<<dim3(Cout, N, chunks), 256, 0, st>>>(dy, dbias, Cout, P);
    }
    SAN_LAUNCH_CHECK();
  }
  return SAN_OK;
}

}  // extern "C"
