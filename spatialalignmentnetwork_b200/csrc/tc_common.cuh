// PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, wgrad_tc.cu): mbarrier, TMA bulk copy,
// tcgen05 alloc / mma / commit / ld, UMMA descriptors.  sm_100a only.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "san_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Polling wait that yields the issue slots between polls (roles that idle next to warps doing real work: the converter
// warps of conv_rows_kernel share their schedulers with 18 waiting warps)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(100);
  }
}
// One lane polls (yielding between polls), the warp re-converges behind it: 32 lanes spinning on try_wait cost issue slots
// that the converter / MMA-issuing warps of conv_rows_kernel need (ncu: 508 M warp instructions, two thirds of them polls)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);      // (no back-off: these waits sit on the row -> MMA -> row latency chain)
  __syncwarp();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One elected lane of a converged warp (cute::elect_one_sync).  The MMA-issuing warp runs its loops
// warp-uniformly and guards only the tcgen05 instructions with this predicate: the compiler then knows
// exactly one thread is active and feeds UTCHMMA from uniform registers directly (under a plain
// `if (lane == 0)` it emitted an ELECT / R2UR / BRA.U.ANY loop around every MMA, ~60 cycles each).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// two 8-column loads in flight before ONE wait (the two column blocks of the HLS accumulator)
__device__ __forceinline__ void tc_ld8x2(uint32_t t0, uint32_t t1, float* a, float* b) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%17];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(t0), "r"(t1)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); }
}

// three 8-column loads (the dx blocks of the DXN accumulator) in flight before ONE wait
__device__ __forceinline__ void tc_ld8x3(uint32_t t0, uint32_t t1, uint32_t t2, float* a, float* b, float* c) {
  uint32_t r[24];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%24];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%25];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%26];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
      : "r"(t0), "r"(t1), "r"(t2)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); c[i] = __uint_as_float(r[16 + i]); }
}

// UMMA shared-memory descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor): start address,
// leading byte offset (between the two 8-element K groups of one K=16 step), stride byte offset
// (between 8-row groups), all in 16 B units; version 1 (bit 46); layout type 0 (bits 61..63).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bit 4), A / B format (bits 7..9 / 10..12: 0 = f16,
// 1 = bf16; on the B200 both operands must use the SAME format - a mixed f16 x bf16 MMA faults), A / B major (bits 15 / 16: 0 = K-major, 1 = MN-major), N>>3 at bit 17,
// M>>4 at bit 24.  fmt: bit 0 = A is an fp16 pair, bit 1 = B is an fp16 pair (else bf16 pairs), see TC_FMT_*.
__host__ __device__ inline uint32_t umma_idesc_16(int M, int N, int fmt, int mn_major = 0) {
  return (1u << 4) | ((fmt & 1) ? 0u : (1u << 7)) | ((fmt & 2) ? 0u : (1u << 10)) | ((uint32_t)(mn_major ? 3 : 0) << 15) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ inline uint32_t umma_idesc_bf16(int M, int N, int mn_major = 0) { return umma_idesc_16(M, N, 0, mn_major); }

// ---- split formats of the staged 16-bit operand pairs -------------------------------------------------------------
// Every fp32 operand value v is staged as TWO 16-bit numbers (hi, lo) and a product is accumulated in fp32 TMEM as
// hi*hi + lo*hi + hi*lo (3 tcgen05.mma per K-step).
//   bf16 pair  (fmt bit clear): hi = bf16(v), lo = bf16(v - hi).  Full fp32 exponent range, ~17 significant bits:
//              used for gradients (dY), whose magnitude is unbounded below.
//   fp16 pair  (fmt bit set)  : z = s*v with a static power-of-two scale s; hi = fp16(z), lo = fp16(z - hi): 22
//              significant bits (|z - hi - lo| <= max(2^-24 |z|, 2^-25)), i.e. fp32-class products, for operands of
//              known O(1) magnitude: normalised / activated activations and network inputs (s = TC_SX; the
//              InstanceNorm bound |x| <= sqrt(H*W) keeps s*x far below the fp16 maximum, values are clamped anyway)
//              and weights (s = TC_SW); gradients (dY) use a DYNAMIC power-of-two scale from the tensor's largest
//              magnitude (san_absmax -> tc_dyn_scale).  The kernel epilogue multiplies by the exact inverse scale.
// The two operands of one MMA must share the format: mixing a bf16 with an f16 operand in kind::f16 raises an
// illegal-instruction fault on the B200 (measured, round 2), so the backward GEMMs run on fp16 pairs as well.
// Why: at the headline configuration (12 cascades, 320x320) bf16 pairs put img_rec 4.9e-3 from the fp32 reference on
// ill-conditioned (fresh-init) weights, 26x the fp32-vs-fp64 floor (profiles/r2b_parity_*.json); fp16 pairs cost
// the same 3 MMAs and the same bytes.
constexpr int TC_FMT_A_F16 = 1, TC_FMT_B_F16 = 2;
constexpr float TC_SX = 16.f;       // activations: typical |x| ~ 0.5 -> 8; fp16 max 65504 -> |x| < 4094
constexpr float TC_SW = 256.f;      // weights: Kaiming bound 0.02 .. 0.2 -> 5 .. 50; |w| < 255
constexpr float TC_F16_MAX = 65504.f;

// Dynamic scale of an fp16-pair operand whose magnitude is not known statically (gradients): the power of two that maps
// the tensor's largest magnitude `amax` (device scalar from san_absmax) into [2^13, 2^14).  Elements down to 2^-28 of
// the maximum keep full precision; smaller ones are absolutely negligible in a dot product.  amax == 0 -> 1.
__device__ __forceinline__ float tc_dyn_scale(float amax) {
  if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;
  int e = (int)((__float_as_uint(amax) >> 23) & 0xff) - 127;       // floor(log2(amax)) for normal numbers
  e = 13 - e;
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return __uint_as_float((uint32_t)(e + 127) << 23);
}

// (hi, lo) bit patterns of one value in the given pair format
__device__ __forceinline__ void split16(float v, bool f16, float scale, unsigned short& hi, unsigned short& lo) {
  if (f16) {
    float z = v * scale;
    z = z > TC_F16_MAX ? TC_F16_MAX : (z < -TC_F16_MAX ? -TC_F16_MAX : z);   // saturate; NaN stays NaN
    const __half h = __float2half_rn(z);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(z - __half2float(h)));
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
  }
}
// two values at once -> packed (hi0 | hi1 << 16), (lo0 | lo1 << 16): one F2FP (packed, saturating) per pair and half
__device__ __forceinline__ void split16x2(float v0, float v1, bool f16, float scale, uint32_t& hw, uint32_t& lw) {
  if (f16) {
    const float z0 = v0 * scale, z1 = v1 * scale;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw) : "f"(z1), "f"(z0));     // saturates, NaN stays NaN
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lw) : "f"(z1 - hf.y), "f"(z0 - hf.x));
  } else {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hw) : "f"(v1), "f"(v0));
    const float h0 = __uint_as_float(hw << 16), h1 = __uint_as_float(hw & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lw) : "f"(v1 - h1), "f"(v0 - h0));
  }
}
__device__ __forceinline__ float unsplit16(unsigned short hi, unsigned short lo, bool f16, float inv_scale) {
  if (f16) return (__half2float(__ushort_as_half(hi)) + __half2float(__ushort_as_half(lo))) * inv_scale;
  return __bfloat162float(__ushort_as_bfloat16(hi)) + __bfloat162float(__ushort_as_bfloat16(lo));
}

inline int pad16(int c) { return (c + 15) / 16 * 16; }

// A staged activation buffer is [TC_LEAD zeros][N][2][KG][PS][8][TC_TRAIL zeros] (bf16 elements): the
// weight-gradient kernel reads one pixel slot before the first plane and up to 17 slots past the last
// one (always multiplied by zero border pixels of dY, but the values must be finite and in bounds).
constexpr int TC_LEAD = 8;
constexpr int TC_TRAIL = 256;


}  // namespace
