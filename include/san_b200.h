/* san_b200.h — C ABI of libsan_b200.so, the B200 (sm_100a) kernels behind the
 * VarNet + spatial-alignment hot path of woxuankai/SpatialAlignmentNetwork.
 *
 * The reference has no FFI: its arithmetic is PyTorch library calls made from
 * flat Python modules.  Each entry point below names the reference call site
 * (file:line under the reference repo) whose arithmetic it replaces; the
 * Python modules in spatialalignmentnetwork_b200/ bind them with ctypes and
 * re-expose the reference's module API (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer into caller-owned, contiguous memory
 *    (complex64 = interleaved float pairs; "planar" = [.., 2, H, W] float with
 *    the real plane first); the library allocates nothing per call except
 *    cached twiddle tables / repacked weights, never synchronises, and launches
 *    on `stream` (a cudaStream_t passed as void*);
 *  - return 0 on success, <0 on error (SAN_ERR_*); san_last_error() returns the
 *    message of the last failure on the calling thread;
 *  - `scratch` arguments are small caller-owned device buffers (size stated).
 */
#ifndef SAN_B200_H_
#define SAN_B200_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* san_last_error(void);
int san_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long san_launch_count(void);

/* ---- FFT / data consistency  (signal_utils.py:4-12, varnet.py:395-402,486,508-530) ---- */
size_t san_fft_workspace_bytes(int B, int H, int W);
/* out = (i)fft2(in) with norm='ortho' over the last two dims of [B,H,W].
 * in/out may be complex64 or planar ([B,2,H,W]); colmask_in multiplies column w of a
 * complex64 input by colmask_in[w] before the transform (ACS mask, varnet.py:395-402),
 * colmask_out multiplies the complex64 output (its adjoint).  tmp: B*H*W complex64
 * (may alias `out` when out is complex64). */
int san_fft2(const void* in, int in_planar, const float* colmask_in, void* out, int out_planar,
             const float* colmask_out, void* tmp, int B, int H, int W, int inverse, void* stream);
/* sens_reduce, varnet.py:511-512: x[n] = sign * sum_c ifft2(k[n,c]) * conj(S[n,c]); x planar
 * [N,2,H,W]; u_out (optional, complex64 [N,C,H,W]) receives sign*ifft2(k) for the backward. */
int san_fft_reduce(const void* k, const void* sens, float* x_planar, void* u_out, void* tmp, int N, int C,
                   int H, int W, int inverse, float sign, void* stream);
/* sens_expand + soft data consistency, varnet.py:508-509,525-530:
 *   out = k - where(mask[w], k - k0, 0) * dc_weight - fft2(x * S)
 * x planar [N,2,H,W]; mask = W bytes (bool); dc_weight = 1 float on device.
 * With k == NULL: out = fft2(x * S) (the adjoint of sens_reduce). */
int san_fft_expand_dc(const float* x_planar, const void* sens, const void* k, const void* k0,
                      const unsigned char* mask, const float* dc_weight, void* out, void* tmp, int N, int C,
                      int H, int W, int inverse, void* stream);
/* rss(ifft2(k)), varnet.py:486: out [N,H,W] float; u_out optional complex64 copy of ifft2(k). */
int san_fft_rss(const void* k, float* out, void* u_out, void* tmp, int N, int C, int H, int W, int inverse,
                void* stream);
/* out[b,p] = sign * u[b,p] * conj(planar[n = b / C, p])  (sensitivity-map gradients) */
int san_cmul_conj_planar(const void* u, const float* planar, void* out, int N, int C, long long P, float sign,
                         void* stream);
/* backward of the soft-DC term: dk = G - where(mask, G, 0)*w, dk0 = where(mask, G, 0)*w (either
 * may be NULL); d_dc_weight = -sum Re(conj(where(mask, k-k0, 0)) * G).  scratch: 1 double. */
int san_dc_bwd(const void* G, const void* k, const void* k0, const unsigned char* mask, const float* dc_weight,
               void* dk, void* dk0, float* d_dc_weight, double* scratch, long long rows, int W, void* stream);
/* signal_utils.rss (signal_utils.py:24-26) over dim 1 of [N,C,P] (complex64 or float) */
int san_rss_fwd(const float* x, float* out, int N, int C, long long P, int is_complex, void* stream);
int san_rss_bwd(const float* g, const float* x, const float* r, float* dx, int N, int C, long long P,
                int is_complex, void* stream);
/* S = s / (rss_c(s) + eps), varnet.py:419; s planar [N*C,2,P] -> S complex64 [N,C,P] */
int san_sens_normalize_fwd(const float* s_planar, void* S, int N, int C, long long P, float eps, void* stream);
int san_sens_normalize_bwd(const void* G, const float* s_planar, float* ds_planar, int N, int C, long long P,
                           float eps, void* stream);
/* out = a*x + b*y (y may be NULL): residual adds (unet.py:15-24) and gradient accumulation */
int san_axpby(const float* x, const float* y, float* out, float a, float b, long long n, void* stream);
/* out[0] = max |x[i]| over finite elements (device scalar): the dynamic scale of fp16-pair GRADIENT operands (dY of the
 * data- and weight-gradient GEMMs; no reference counterpart: cuDNN computes these in fp32) */
int san_absmax(const float* x, long long n, float* out, void* stream);

/* ---- convolutions (varnet.py:75-80,139-146,176-179; unet.py:119-140; cross.py:15) ---- */
/* repack OIHW weights for the kernels: dgrad=0 -> [Cin][K*K][Cout]; dgrad=1 -> flipped,
 * [Cout][K*K][Cin] (so the data-gradient is the same forward kernel run on dY). */
int san_conv_pack_weights(const float* w, float* packed, int Cout, int Cin, int K, int dgrad, void* stream);
/* y = conv2d(x, w) + bias, stride 1, padding K/2, K in {1,3}; NCHW fp32.
 * x_bs / y_bs: batch strides in floats (0 = dense), so outputs can land inside a
 * wider concat buffer. */
int san_conv2d_fwd(const float* x, const float* w_packed, const float* bias, float* y, int N, int Cin, int H, int W,
                   int Cout, int K, long long x_bs, long long y_bs, void* stream);
/* dw[Cout,Cin,K,K] (OIHW) and optional dbias[Cout] */
int san_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias, int N, int Cin, int H, int W, int Cout,
                     int K, long long x_bs, long long dy_bs, void* stream);
/* x [N,Co*4,H,W] (channel = co*4 + a*2 + b) <-> y [N,Co,2H,2W]: ConvTranspose2d(k=2,s=2) =
 * 1x1 conv to 4*Co channels + depth_to_space */
int san_depth_to_space2(const float* x, float* y, int N, int Co, int H, int W, void* stream);
int san_space_to_depth2(const float* y, float* x, int N, int Co, int H, int W, void* stream);

/* ---- tcgen05 tensor-core convolutions (same call sites as above; 16-bit pair split, fp32 TMEM accumulate) ----
 * Every fp32 operand value is staged as a PAIR of 16-bit numbers (hi, lo) and a product is accumulated as
 * hi*hi + lo*hi + hi*lo (3 tcgen05.mma per K-step).  Two pair formats (`fmt` arguments below):
 *   0  bf16 pair: hi = bf16(v), lo = bf16(v - hi): full fp32 exponent range, ~17 significant bits (gradients);
 *   1  fp16 pair: z = s*v (static power-of-two s: 16 for activations, 256 for weights), hi = fp16(z),
 *      lo = fp16(z - hi): 22 significant bits, fp32-class products, for operands of O(1) magnitude (normalised
 *      activations, network inputs, weights); the kernels undo the scale exactly in their epilogues.
 *      Gradient operands (dY) have no static magnitude: their fp16 pairs use a DYNAMIC power-of-two scale derived
 *      on the device from max|dY| (san_absmax); pass that device scalar as `absmax` / `a_absmax` / `dy_absmax`.
 * For san_tc_conv / san_tc_wgrad `fmt` is 0 (both operands bf16 pairs) or 3 (both fp16 pairs): the two operands of
 * one tcgen05 kind::f16 MMA must share the format on the B200 (a mixed f16 x bf16 MMA faults).
 * Activations are staged as Xs[n][hl][kg][(H+2)*(W+2)][8] 16-bit (hl = hi/lo halves of the fp32 value,
 * kg = ceil(C/8) groups of 8 channels - `Cpad` of the staging calls is C padded to 8 -, one-pixel zero border); weights as
 * Ws[nsplit][KS][taps][hl][2][Npad][8].  Element counts of the caller-allocated buffers: */
long long san_tc_staged_act_elems(int N, int H, int W, int C);
long long san_tc_staged_weight_elems(int H, int W, int Cout, int Cin, int K);   /* Cout/Cin of the LAUNCH */
int san_tc_supported(int H, int W, int Cin, int Cout, int K);
/* host-only: strip geometry of san_tc_conv for this shape; out[16] = Cin_pad, KG, KS, nsplit, Npad, Wp, Hp, R, T,
 * S_alloc, strips, stages, acc_stages, a_bytes, b_bytes, smem_bytes (host pointer) */
int san_tc_describe(int H, int W, int Cin, int Cout, int K, int* out);
/* host-only: out[7] = dxn (1 = "DXN" form, opt-in SAN_TC_DXN=1: the three horizontal taps sit in the MMA N dimension,
 * B row = dx * Np + co, and the epilogue adds the column blocks across neighbouring pixels), Np (Cout padded to 8),
 * wtaps (weight blocks per K-step: 9, 3 or 1), xchg_bytes, hls (1 = narrow 3x3 layers, <= 32 padded output channels:
 * [W_hi | W_lo] stacked along the MMA N dimension, 2 instead of 3 reads of the A tile per tap), Ncol (TMEM columns per
 * 128-pixel tile), pair (1 = odd number of 8-channel input groups: the half-empty last K = 16 step pairs filter taps, 5
 * MMA steps instead of 9, and never loads the all-zero group) (host pointer) */
int san_tc_describe_form(int H, int W, int Cin, int Cout, int K, int* out);
/* Fused operand producer: up to 3 channel-concatenated sources (varnet.py:116 concat order),
 * each out = leaky_relu(a[plane]*(y - mu[plane]) + b[plane], slope) (a NULL = identity), i.e. the
 * InstanceNorm / BatchNorm + LeakyReLU of varnet.py:141-145 / unet.py:124-126 applied on the fly;
 * mode 0 same resolution, 1 avg_pool2d(2) of the activated source (varnet.py:98), 2 depth-to-space
 * of a [N,4C,H/2,W/2] source (the ConvTranspose2d pixel shuffle), 3 nearest x2 (unet.py:130). */
int san_tc_stage_act(void* xs, int N, int H, int W, int Cpad,
                     const float* y0, const float* mu0, const float* a0, const float* b0, float slope0, int C0, int mode0,
                     const float* y1, const float* mu1, const float* a1, const float* b1, float slope1, int C1, int mode1,
                     const float* y2, const float* mu2, const float* a2, const float* b2, float slope2, int C2, int mode2,
                     int fmt, void* stream);
/* General form: an ordered list of terms; a term with accumulate = 0 starts a new channel range right after
 * the previous range (concatenation), accumulate = 1 ADDS onto the previous term's range (the residual sums of
 * unet.py:15-24: x + subnet(x) of two activated tensors).  At most 6 terms; host memory. */
typedef struct san_stage_term {
  const float* y;   /* fp32 NCHW source at its own resolution */
  const float* mu;  /* per-plane [N*C] centre / scale / shift; a == NULL: identity */
  const float* a;
  const float* b;
  float slope;      /* LeakyReLU slope, 1 = none */
  int C;            /* channels contributed (after the pixel shuffle for mode 2) */
  int mode;         /* 0 direct, 1 avg-pool 2x2, 2 depth-to-space, 3 nearest x2 */
  int accumulate;
} san_stage_term;
int san_tc_stage_terms(void* xs, int N, int H, int W, int Cpad, const san_stage_term* terms, int nterms, int fmt,
                       const float* absmax, void* stream);
/* x[N,C,H,W] = hi + lo of a staged tensor (input of the fp32 weight-gradient kernel) */
int san_tc_unstage_act(const void* xs, float* x, int N, int C, int H, int W, int fmt, void* stream);
/* OIHW fp32 -> staged hi/lo for images of H x W (the output-channel split depends on the strip geometry);
 * dgrad = 1 stages the flipped, transposed filter so that the data gradient is san_tc_conv run on the
 * staged dY (Cout/Cin below are always the ORIGINAL ones of the OIHW tensor) */
int san_tc_stage_weights(const float* w, void* ws, int H, int W, int Cout, int Cin, int K, int dgrad, int fmt,
                         void* stream);
/* y[N,Cout,H,W] (+ bias) = conv2d(staged x, staged w), stride 1, padding K/2, K in {1,3};
 * Cin/Cout are the channel counts of THIS launch (for dgrad: Cin = original Cout, Cout = original Cin) */
int san_tc_conv(const void* xs, const void* ws, const float* bias, float* y, int N, int H, int W, int Cin, int Cout,
                int K, long long y_bs, int fmt, const float* a_absmax, void* stream);
/* The same convolution with a STATISTICS EPILOGUE: sums[N][Cout][2] (fp64, zeroed by the call) receives the per-plane sum
 * and sum of squares of y, accumulated by the epilogue warps while they store y (per-warp partials in shared memory,
 * flushed with fp64 atomicAdd when a CTA moves to another image: reproducible) - what the reference's InstanceNorm2d (varnet.py:141) would
 * otherwise re-read the whole tensor for.  san_in_stats_from_sums turns them into the coefficient table of
 * san_plane_stats_in (mean, m2, a = rstd, b = 0); group = 4 for the pixel-shuffled ConvTranspose2d output (four
 * sub-planes of a channel normalise together, varnet.py:176-181), P = elements per sub-plane. */
int san_tc_conv_stats_supported(int H, int W, int Cin, int Cout, int K);
int san_tc_conv_stats(const void* xs, const void* ws, const float* bias, float* y, int N, int H, int W, int Cin, int Cout,
                      int K, long long y_bs, int fmt, const float* a_absmax, double* sums, void* stream);
int san_in_stats_from_sums(const double* sums, float* mean, float* m2, float* a, float* b, int planes, int group, int P,
                           float eps, void* stream);
/* Row-ring form of the 3x3 conv for the narrow full-resolution layers (<= 24 input, <= 32 output channels): the kernel reads
 * the RAW fp32 input x[N,Cin,H,W] and applies the producing layer's per-plane normalisation + LeakyReLU
 * (act(a*(x-mu)+b), a null = identity: network inputs, gradients) and the fp16-pair split ITSELF - what
 * san_tc_stage_terms + san_tc_conv do in two passes over HBM (reference varnet.py:141-145 followed by :140,143).  absmax: the
 * dynamic scale of a gradient operand (identity only).  xs_out (may be null): the staged form of the operand
 * (san_tc_staged_act_elems), written row by row with TMA bulk stores for the weight-gradient GEMM.  sums (may be null): the
 * statistics epilogue of san_tc_conv_stats.  Weights: san_tc_stage_weights_rows (fp16 pairs, fmt = 1) into
 * san_tc_rows_weight_elems elements; Cout / Cin of the _elems / _supported / conv calls are those of the LAUNCH. */
int san_tc_conv_rows_supported(int H, int W, int Cin, int Cout, int K);
/* host-only: geometry of the row-ring kernel; out[12] = KG, KS, Npad, Ncol (TMEM columns per tile), Wp, T (128-pixel tiles
 * per row), RS (slots per row plane), pair (2: taps of the last K-step paired inside a filter row), NR (ring slots),
 * row_bytes, w_bytes, smem_bytes */
int san_tc_conv_rows_describe(int H, int W, int Cin, int Cout, int stats, int* out);
long long san_tc_rows_weight_elems(int H, int W, int Cout, int Cin);
int san_tc_stage_weights_rows(const float* w, void* ws, int H, int W, int Cout, int Cin, int dgrad, int fmt, void* stream);
int san_tc_conv_rows(const float* x, const float* mu, const float* a, const float* b, float slope, const float* absmax,
                     void* xs_out, const void* ws, const float* bias, float* y, double* sums, int N, int H, int W, int Cin,
                     int Cout, void* stream);

/* dW[Cout,Cin,K,K] (and dbias[Cout] from the fp32 dy, both optional-bias) from the staged dY and the staged
 * input of the forward conv: tcgen05 GEMM over the pixel dimension, MN-major operands, BF16x3 */
int san_tc_wgrad_supported(int H, int W, int Cin, int Cout, int K);
/* host-only: decomposition of san_tc_wgrad; out[16] = KGo, KGi, nmb, nnc, ndy, Nn, KGn, KC, XS, stages, smem_bytes,
 * nchunks, Wp, PS, range0, range_len (host pointer) */
int san_tc_wgrad_describe(int H, int W, int Cin, int Cout, int K, int* out);
/* host-only: out[2] = rown (1 = the three filter rows sit in the MMA N dimension: three row-shifted copies of the X
 * span per stage, 3x3 layers with <= 48 padded input channels), ncp (X copies per stage) */
int san_tc_wgrad_describe_form(int H, int W, int Cin, int Cout, int K, int* out);
int san_tc_wgrad(const void* dys, const void* xs, float* dw, float* dbias, const float* dy, int N, int H, int W, int Cin,
                 int Cout, int K, int fmt, const float* dy_absmax, void* stream);

/* ---- normalisation / activation / resampling (varnet.py:98,139-146,235,257-273; unet.py:119-140) ---- */
/* per-plane mean and centred sum of squares (two-pass) */
int san_plane_stats(const float* x, float* mean, float* m2, int planes, int P, void* stream);
/* the same + the InstanceNorm2d coefficients (a = rstd = 1/sqrt(m2/P + eps), b = 0; varnet.py:141, biased variance) in
 * one launch: san_plane_stats followed by san_in_finalize_fwd */
int san_plane_stats_in(const float* x, float* mean, float* m2, float* a, float* b, int planes, int P, float eps,
                       void* stream);
/* InstanceNorm coefficients: a = rstd, b = 0 (the centre mu is the `mean` array itself) */
int san_in_finalize_fwd(const float* mean, const float* m2, float* a, float* b, int planes, int P, float eps,
                        void* stream);
/* BatchNorm coefficients per plane [N*C]: mu, a = gamma*rstd, b = beta, sa = rstd; updates the
 * running buffers in training mode (momentum, unbiased variance) */
int san_bn_finalize_fwd(const float* mean, const float* m2, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float* mu, float* a, float* b, float* sa, int N,
                        int C, int P, float eps, float momentum, int training, void* stream);
/* out = leaky_relu(a[plane] * (y - mu[plane]) + b[plane], slope); mu, b may be NULL (= 0) */
int san_affine_act_fwd(const float* y, const float* mu, const float* a, const float* b, float slope, float* out,
                       int planes, int P, void* stream);
/* s1 = sum g', s2 = sum g' * sa*(y - mu) per plane, g' = g * lrelu'(a*(y - mu) + b); sa NULL = 1 */
int san_act_bwd_reduce(const float* g, const float* y, const float* mu, const float* a, const float* b,
                       const float* sa, float slope, float* s1, float* s2, int planes, int P, void* stream);
int san_in_finalize_bwd(const float* s1, const float* s2, const float* a, float* p, float* q,
                        float* r, int planes, int P, void* stream);
int san_bn_finalize_bwd(const float* s1, const float* s2, const float* gamma, const float* sa,
                        float* p, float* q, float* r, float* dgamma, float* dbeta, int N, int C, int P, int training,
                        void* stream);
/* dy = p*g' + q*(y - mu) + r (q, r may be NULL) */
int san_act_bwd_apply(const float* g, const float* y, const float* mu, const float* a, const float* b, float slope,
                      const float* p, const float* q, const float* r, float* dy, int planes, int P, void* stream);
/* the same two passes with g read IN PLACE from the data gradient dx[N, Ctot, Hd, Wd] of the consuming conv:
 * channel offset c0 of a concatenated source, and the adjoint of the resampling applied while staging
 * (mode 0 direct: Hd x Wd = Hy x Wy; 1 avg-pool: dx is Hy/2 x Wy/2; 2 pixel shuffle: y is [N, 4*Cy, Hy, Wy], dx is
 * 2Hy x 2Wy; 3 nearest x2: dx is 2Hy x 2Wy).  No slice copies, no up2 / space_to_depth2 / pool2 temporaries. */
int san_act_bwd_reduce_map(const float* g, int Ctot, int c0, int mode, const float* y, const float* mu, const float* a,
                           const float* b, const float* sa, float slope, float* s1, float* s2, int N, int Cy, int Hy,
                           int Wy, void* stream);
/* absmax (optional device scalar): receives max |dy| - the producing layer's backward stages dy as an fp16 pair with
 * the dynamic scale derived from it, so no separate san_absmax pass over dy is needed */
int san_act_bwd_apply_map(const float* g, int Ctot, int c0, int mode, const float* y, const float* mu, const float* a,
                          const float* b, float slope, const float* p, const float* q, const float* r, float* dy, int N,
                          int Cy, int Hy, int Wy, float* absmax, void* stream);
/* InstanceNorm2d + LeakyReLU backward of varnet.py:141-145 in ONE kernel per tensor (the three calls above fused for
 * per-plane statistics): one CTA per (n, c) plane reduces sum(g'), sum(g' xhat), forms the coefficients and applies
 * them, re-reading the plane from L2 in reverse order.  mu / a = mean and rstd per plane (the 'in' coefficient table). */
int san_in_bwd_fused_map(const float* g, int Ctot, int c0, int mode, const float* y, const float* mu, const float* a,
                         float slope, float* dy, int N, int Cy, int Hy, int Wy, float* absmax, void* stream);
/* y = scale * (2x2 block sum of x): avg_pool2d (scale .25) and the adjoint of nearest up-sampling (scale 1) */
int san_pool2(const float* x, float* y, long long planes, int H, int W, float scale, void* stream);
/* y[2h+a,2w+b] = scale * x[h,w]: nearest x2 (scale 1) and the adjoint of avg_pool2d (scale .25) */
int san_up2(const float* x, float* y, long long planes, int H, int W, float scale, void* stream);

/* ---- spatial alignment (cross.py:23-38, model.py:21-28) ---- */
int san_grid_from_offset(const float* x_nchw, float* grid, int N, int H, int W, void* stream);
int san_grid_to_nchw(const float* g_nhwc, float* dx_nchw, int N, int H, int W, void* stream);
int san_warp_fwd(const float* img, const float* grid, float* out, int N, int C, int H, int W, int Ho, int Wo,
                 void* stream);
int san_warp_bwd(const float* gout, const float* img, const float* grid, float* dimg, float* dgrid, int N, int C,
                 int H, int W, int Ho, int Wo, void* stream);
/* s[n,h,w,c] at n*sn + h*sh + w*sw + c*sc (floats), c in {0,1}.  scratch: 2 doubles. */
int san_grad_loss_fwd(const float* s, long long sn, long long sh, long long sw, long long sc, int N, int H, int W,
                      float* out, double* scratch, void* stream);
int san_grad_loss_bwd(const float* s, long long sn, long long sh, long long sw, long long sc, int N, int H, int W,
                      const float* gout, float* ds, void* stream);

/* ---- losses (ssimloss.py:11-40, lnccloss.py:7-56, miloss.py:6-57) ---- */
/* images [N,1,H,W] float; out = 1 scalar on device; scratch: 1 double */
int san_ssim_loss_fwd(const float* X, const float* Y, int N, int H, int W, float* out, double* scratch, void* stream);
int san_ssim_loss_bwd(const float* X, const float* Y, const float* gout, int N, int H, int W, float* dX, float* dY,
                      void* stream);
int san_lncc_loss_fwd(const float* I, const float* J, int N, int H, int W, float* out, double* scratch, void* stream);
int san_lncc_loss_bwd(const float* I, const float* J, const float* gout, int N, int H, int W, float* dI, float* dJ,
                      void* stream);
/* Parzen soft histograms: joint[N,64,64] = p_I p_J^T, mI/mJ[N,64] = row sums (miloss.py:26-42) */
int san_mi_hist_fwd(const float* I, const float* J, float* joint, float* mI, float* mJ, int N, int P, int bins,
                    float sigma, float minv, float maxv, void* stream);
int san_mi_hist_bwd(const float* I, const float* J, const float* gjoint, const float* gmI, const float* gmJ, float* dI,
                    float* dJ, int N, int P, int bins, float sigma, float minv, float maxv, void* stream);
/* single-channel KxK correlation, zero padding K/2 (gaussian_smooth, miloss.py:13-24) */
int san_filter2d(const float* x, const float* w, float* y, long long planes, int H, int W, int K, void* stream);

/* ---- misalignment augmentation of the auxiliary modality (augment.py:7-66, train.py:207-212) ---- */
/* grid[n,h,w,:] = affine_grid(theta[n] (2x3), align_corners=False) + bicubic up-sampling (align_corners=False)
 * of the control displacements ctrl[n] ([2,G,G], channel 0 = x; NULL: rigid only); grid [N,H,W,2] */
int san_augment_grid(const float* theta, const float* ctrl, int G, float* grid, int N, int H, int W, void* stream);
/* grid_sample(img, grid, bilinear, padding_mode='reflection', align_corners=False) on every plane of
 * img [N,C,H,W] with `interleave` float components per pixel (1 = float, 2 = complex64: real and imaginary
 * parts are sampled separately, augment.py:57-60); out [N,C,Ho,Wo] in the same format */
int san_warp_reflect(const float* img, const float* grid, float* out, int N, int C, int H, int W, int Ho, int Wo,
                     int interleave, void* stream);

/* ---- GAN branch: spectral norm + point-wise losses (gan.py:24,131-137; model.py:138-139) ---- */
/* torch.nn.utils.spectral_norm (gan.py:24; one power iteration, dim 0) on W [rows, cols]:
 * power_iteration != 0 (training): v = normalize(W^T u, eps), u = normalize(W v, eps) IN PLACE;
 * then sigma[0] = u . (W v).  tmp: max(rows, cols) floats. */
int san_sn_sigma(const float* w, float* u, float* v, float* tmp, float* sigma, int rows, int cols, float eps,
                 int power_iteration, void* stream);
/* out = w / sigma[0] */
int san_sn_scale(const float* w, const float* sigma, float* out, long long n, void* stream);
/* dW = (G - <G, W_sn> u v^T) / sigma   (u, v constants, as in torch).  scratch: 1 double. */
int san_sn_bwd(const float* g, const float* w_sn, const float* u, const float* v, const float* sigma, double* scratch,
               float* dw, int rows, int cols, void* stream);
/* out[0] = mean_i f(x_i, y_i): mode 0 |x - y| (F.l1_loss, model.py:138), 1 max(sign*x, -1) (hinge,
 * gan.py:133), 2 sign*x (gan.py:135).  y only for mode 0.  scratch: 1 double. */
int san_pair_loss_fwd(const float* x, const float* y, long long n, int mode, float sign, float* out, double* scratch,
                      void* stream);
int san_pair_loss_bwd(const float* x, const float* y, const float* gout, long long n, int mode, float sign, float* dx,
                      float* dy, void* stream);

/* ---- evaluation metrics (metrics.py:23-68, model.py:265-286) ---- */
/* out3 = { sum (a-b)^2, sum |a-b|, sum a^2 } in fp64 over n floats */
int san_error_sums(const float* a, const float* b, long long n, double* out3, void* stream);
/* per image n: np.histogram2d(x[n], y[n], bins, range [minv, maxv]) -> plug-in mutual information; out [N] doubles */
int san_mi_metric(const float* x, const float* y, int N, int P, int bins, float minv, float maxv, double* out,
                  void* stream);

/* ---- optimiser (model.py:72-81: torch.optim.AdamW, one instance per network) ---- */
/* One AdamW step (no amsgrad) on `ntensors` fp32 tensors.  params / grads / exp_avg / exp_avg_sq / numel are HOST
 * arrays of device pointers (and element counts); `step` is the 1-based step count of the bias corrections.
 * step_dev / sched_dev (optional, device int / float[2]): the step counter lives on the DEVICE instead - it is incremented
 * and the bias-correction scalars are formed there (same fp64 arithmetic), which makes the call CUDA-graph capturable. */
int san_adamw_step(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                   const long long* numel, int ntensors, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int step, int* step_dev, float* sched_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAN_B200_H_ */
