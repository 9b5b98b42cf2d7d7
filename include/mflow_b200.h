/*
 * mflow_b200.h - C ABI of libmflow_b200.so: the B200 (sm_100a) kernels behind the
 * manifold-flow neural-spline coupling hot path.
 *
 * The reference (johannbrehmer/manifold-flow) has no FFI: its boundary is the Python
 * class API of the `manifold_flow` package.  The drop-in `manifold_flow` package in
 * manifold-flow_b200/ keeps that API and calls the entry points below through ctypes
 * (manifold_flow/b200/cabi.py).  Each entry point names the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 (or int32/int64 where stated) unless it is a
 *    `const mflow_*_desc*`, which is a HOST pointer read before the launch;
 *  - nothing allocates, synchronises or touches the default stream: the caller provides
 *    outputs and workspaces and a `cudaStream_t` (passed as void*), so every call is
 *    CUDA-graph capturable;
 *  - return value 0 = launched; non-zero = MFLOW_E_* and mflow_last_error_string() explains;
 *  - no C++ types, no torch types, no exceptions cross this boundary.
 */
#ifndef MFLOW_B200_H
#define MFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFLOW_ABI_VERSION 1
#if defined(__GNUC__)
#define MFLOW_API __attribute__((visibility("default")))
#else
#define MFLOW_API
#endif

enum {
  MFLOW_OK = 0,
  MFLOW_E_BADARG = 1,   /* unsupported shape / null pointer / misaligned */
  MFLOW_E_LAUNCH = 2,   /* cudaGetLastError() after the launch           */
  MFLOW_E_NODEVICE = 3, /* no sm_100 device visible                       */
  MFLOW_E_WORKSPACE = 4 /* workspace too small                            */
};

MFLOW_API int mflow_version(void);
MFLOW_API const char* mflow_last_error_string(void);
/* 0 if a CUDA device with compute capability 10.x is current, else MFLOW_E_NODEVICE. */
MFLOW_API int mflow_check_device(void);

/* ------------------------------------------------------------------------------------
 * (a) Rational-quadratic spline with linear tails.
 * Replaces manifold_flow/transforms/splines/rational_quadratic.py:16-168
 * (unconstrained_rational_quadratic_spline + rational_quadratic_spline), the bin search
 * manifold_flow/utils/various.py:472-474, the parameter slicing / pre-scale of
 * manifold_flow/transforms/coupling.py:501-511, the per-sample reduction :279 and the
 * split / merge of CouplingTransform.forward/inverse :58-185.
 *
 * Elements are indexed (o, c, i): o = outer index (sample for images, sample for vectors),
 * c = transformed channel / feature, i = inner contiguous index (pixel; 1 for vectors).
 *   x  element  at x  + o*x_so + c*x_sc + i
 *   y  element  at y  + o*y_so + c*y_sc + i
 *   parameter p (0 <= p < 3K-1: K widths, K heights, K-1 derivatives) of the element at
 *              params + o*p_so + c*p_sc + p*p_sp + i*p_si
 * Three layouts are implemented:
 *   "planes" (image, conditioner output [B, C_t*P, H, W]): p_si == 1, p_sp == n_inner
 *   "rows"   (vector, conditioner output [B, C_t*P]):       n_inner == 1, p_sp == 1
 *   "rows x planes" (p_hw > 0): vectors x / y [O, D] (n_inner == 1) whose parameters are the NCHW output
 *            [O / p_hw, C_t*P, p_hw] of a ResidualNet (nn/resnet.py:43-72) that was evaluated on the convolution
 *            kernels with the batch rows as the pixels of p_hw-pixel images: parameter p of element (o, c) at
 *            params + ((o / p_hw) * C_t*P + c*P + p) * p_hw + o % p_hw;  p_so / p_sc / p_sp / p_si are ignored.
 *            x / y (and the gradients) are dense rows of D = x_so = y_so <= 92 features and point at feature 0 of
 *            row 0; transformed feature c sits at t_off + c*x_sc.  The launch writes whole rows of y (g_in): features
 *            that are neither transformed nor listed as identity features are copied through as well.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t n_outer;  /* O */
  int64_t n_chan;   /* C_t */
  int64_t n_inner;  /* I */
  int64_t x_so, x_sc;
  int64_t y_so, y_sc;
  int64_t p_so, p_sc, p_sp, p_si;
  int32_t num_bins; /* K in {2..16, 20, 24, 32} */
  int32_t inverse;  /* 0 forward, 1 inverse */
  float tail_bound; /* T: spline on [-T, T]^2, identity outside (closed interval) */
  float min_bin_width, min_bin_height, min_derivative;
  float wh_divisor; /* widths/heights logits are divided by this (sqrt(hidden)); 1 = off */
  /* identity part copied through by the same launch (coupling merge); n_id_chan == 0 = off.
   *   src at x + o*x_so + id_off + c*id_sc + i,  dst at y + o*y_so + id_off + c*id_sc + i */
  int64_t n_id_chan, id_off, id_sc;
  int64_t p_hw;     /* 0, or the pixels per parameter image of the "rows x planes" layout (a power of two dividing O) */
  int64_t t_off;    /* "rows x planes" only: index of the first transformed feature */
} mflow_rqs_desc;

/* Forward or inverse evaluation.
 *   lad_elem   : optional [O*C_t*I] per-element logabsdet, same indexing as a dense (o,c,i) array
 *   lad_sample : optional [O]; receives lad_base[o] (or 0 when lad_base is NULL) + sum over (c,i)
 *   bins       : optional int32 [O*C_t*I] bin index per element (-1 in the tails)
 *   workspace  : device scratch of mflow_rqs_workspace_bytes(desc) bytes: 256 KiB of per-sample
 *                arrival counters (must be zero on first use; every launch leaves them zeroed)
 *                followed by per-CTA partial sums.  One workspace per stream. */
MFLOW_API int64_t mflow_rqs_workspace_bytes(const mflow_rqs_desc* desc);
MFLOW_API int mflow_rqs_apply(const mflow_rqs_desc* desc, const float* x, const float* params, float* y,
                    float* lad_elem, float* lad_sample, const float* lad_base, int32_t* bins,
                    void* workspace, void* stream);

/* Backward of mflow_rqs_apply.
 *   x_in       : the tensor that was the INPUT of the apply call
 *   x_out      : its OUTPUT (only read when desc->inverse; may be NULL otherwise)
 *   g_out      : gradient w.r.t. the output tensor (same indexing as y), may be NULL (= 0)
 *   g_lad_sample / g_lad_elem : gradient w.r.t. lad_sample [O] / lad_elem (either may be NULL)
 *   g_in       : gradient w.r.t. the input tensor (transform part and, if n_id_chan, identity part)
 *   g_params   : gradient w.r.t. params, same indexing as params; every entry is written
 *   g_params_amax : optional device scalar (zero-initialised by the caller) that receives max |g_params| of the
 *                launch - the operand scale of the fp16-split convolution that back-propagates it (section (b)) */
MFLOW_API int mflow_rqs_backward(const mflow_rqs_desc* desc, const float* x_in, const float* x_out,
                       const float* params, const float* g_out, const float* g_lad_sample,
                       const float* g_lad_elem, float* g_in, float* g_params, float* g_params_amax, void* stream);

/* Spline on a box: rational_quadratic_spline(inputs, W, H, D, inverse, left, right, bottom, top, ...)
 * of manifold_flow/transforms/splines/rational_quadratic.py:67-168 - the bounded spline the unconstrained
 * wrapper calls, as a function of its own: all K+1 knot derivatives are parameters (no padding constant),
 * the x knots span [left, right] and the y knots [bottom, top].
 *   params : [n, 3K+1] rows = [K widths | K heights | K+1 derivatives] (unnormalised)
 *   x, y, lad (per-element logabsdet) : [n];  bins : optional int32 [n]
 * The caller checks the domain (the reference raises InputOutsideDomain, :85-86); an element outside the
 * searched axis' range passes through unchanged with logabsdet 0 and bin -1. */
typedef struct {
  int64_t n;
  int32_t num_bins;
  int32_t inverse;
  float left, right, bottom, top;
  float min_bin_width, min_bin_height, min_derivative;
} mflow_rqs_box_desc;
MFLOW_API int mflow_rqs_box_apply(const mflow_rqs_box_desc* desc, const float* x, const float* params, float* y,
                        float* lad, int32_t* bins, void* stream);
MFLOW_API int mflow_rqs_box_backward(const mflow_rqs_box_desc* desc, const float* x_in, const float* x_out,
                           const float* params, const float* g_out, const float* g_lad, float* g_in,
                           float* g_params, void* stream);

/* Test entry: the kernel's own bin search on caller-provided knots (bit-exactness pin,
 * utils/various.py:472-474).  knots [n, K+1] row-major, values [n] -> bins int32 [n]. */
MFLOW_API int mflow_rqs_search(const float* knots, const float* values, int32_t* bins, int64_t n,
                     int32_t num_bins, void* stream);

/* ------------------------------------------------------------------------------------
 * (c) Linear glue.
 * ---------------------------------------------------------------------------------- */
/* y[o, :, i] = M x[o, :, i] + m  for a dense C x C matrix (C <= 128): the fused form of
 * ActNorm (transforms/normalization.py:99-131), the channel permutation and LU map of
 * OneByOneConvolution (transforms/conv.py:16-55, transforms/lu.py:56-95) and, for vectors,
 * RandomPermutation + LULinear (architectures/vector_transforms.py:39-46).
 *   element (o, c, i) at ptr + o*so + c*sc + i*si;  M row-major [C, C] (y_c = sum_k M[c,k] x_k). */
typedef struct {
  int64_t n_outer, n_chan, n_inner;
  int64_t x_so, x_sc, x_si;
  int64_t y_so, y_sc, y_si;
} mflow_mix_desc;
MFLOW_API int mflow_chanmix_apply(const mflow_mix_desc* desc, const float* x, const float* mat, const float* bias,
                        float* y, void* stream);
/* Weight gradients of the above: g_mat[c,k] = sum_{o,i} g_y[o,c,i] x[o,k,i], g_bias[c] = sum g_y.
 * Deterministic two-pass reduction.  workspace: device scratch of mflow_chanmix_wgrad_workspace_bytes(desc)
 * bytes, no initialisation needed. */
MFLOW_API int64_t mflow_chanmix_wgrad_workspace_bytes(const mflow_mix_desc* desc);
MFLOW_API int mflow_chanmix_wgrad(const mflow_mix_desc* desc, const float* x, const float* g_y, float* g_mat,
                        float* g_bias, void* workspace, void* stream);

/* out[c] = sum_{b,i} x[b, c, i] over a contiguous [batch, chan, inner] tensor: the bias gradient of the
 * conditioner's convolutions (autograd of nn.Conv2d in nn/resnet.py:85-142; ATen sum over dims (0,2,3)).
 * Deterministic two-pass reduction; workspace: mflow_channel_sum_workspace_bytes() bytes, no initialisation. */
MFLOW_API int64_t mflow_channel_sum_workspace_bytes(int64_t batch, int64_t chan, int64_t inner);
MFLOW_API int mflow_channel_sum(const float* x, float* out, int64_t batch, int64_t chan, int64_t inner, void* workspace,
                      void* stream);

/* 2x2 space-to-depth with a fused scalar affine map: y = squeeze(x) * scale + shift (forward) or
 * y = unsqueeze((x - shift) / scale) (inverse).  SqueezeTransform (transforms/reshape.py:29-57) and
 * AffineScalarTransform (transforms/standard.py:41-57).  x [B, C, H, W] contiguous. */
MFLOW_API int mflow_squeeze2x2(const float* x, float* y, int64_t batch, int64_t chan, int64_t height, int64_t width,
                     int32_t inverse, float scale, float shift, void* stream);

/* y = x * scale + shift (forward) / (x - shift) / scale (inverse) over n floats:
 * AffineScalarTransform (transforms/standard.py:41-57). */
MFLOW_API int mflow_affine_scalar(const float* x, float* y, int64_t n, float scale, float shift, int32_t inverse,
                        void* stream);

/* y[o, j] = x[o, perm[j]] over rows of length n: Permutation._permute (transforms/permutations.py:73). */
MFLOW_API int mflow_permute_rows(const float* x, const int64_t* perm, float* y, int64_t rows, int64_t n, void* stream);

/* LogTanh (transforms/nonlinearities.py:38-95) on [rows, n]; lad_sample[o] = sum_j logabsdet. */
MFLOW_API int mflow_logtanh_apply(const float* x, float* y, float* lad_sample, int64_t rows, int64_t n, int32_t inverse,
                        float cut_point, void* stream);
MFLOW_API int mflow_logtanh_backward(const float* x, const float* g_y, const float* g_lad_sample, float* g_x,
                           int64_t rows, int64_t n, int32_t inverse, float cut_point, void* stream);

/* Gated residual of a ResidualBlock with context, manifold_flow/nn/resnet.py:36-40:
 *   out = h + t * sigmoid(g)          (F.glu(cat(t, g)) = t * sigmoid(g), then the skip connection)
 * for the large-batch vector conditioners that run on the convolution kernels (h, t, g: same shape, n floats, n % 4 == 0,
 * 16-byte aligned).  Backward: g_h = g_out (nothing to compute), g_t = g_out * sigmoid(g), g_g = g_out * t * s (1 - s).
 * The optional *_amax device scalars (zero on entry) receive the absolute maxima of the written tensors: the operand
 * scales of the fp16-split convolutions that consume them. */
MFLOW_API int mflow_glu_residual(const float* h, const float* t, const float* g, float* out, float* out_amax, int64_t n,
                       void* stream);
MFLOW_API int mflow_glu_residual_backward(const float* t, const float* g, const float* g_out, float* g_t, float* g_g,
                                float* g_t_amax, float* g_g_amax, int64_t n, void* stream);

/* out = a + b (n floats, n % 4 == 0, 16-byte aligned) with max |out| merged into the optional device scalar out_amax:
 * the gradient accumulation of a residual block's skip connection (nn/resnet.py:27-40, :91-104 backward), fused with the
 * operand-scale pass of the convolution that consumes the sum. */
MFLOW_API int mflow_add_amax(const float* a, const float* b, float* out, float* out_amax, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * (b') Vector conditioner: the whole ResidualNet (manifold_flow/nn/resnet.py:9-72) in one launch.
 *   h = W_0 [x | ctx] + b_0;  per block: h += gate * (W_b relu(W_a relu(h) + b_a) + b_b),
 *   gate = sigmoid(W_c ctx + b_c) when n_ctx > 0 (the GLU of :38-39), else 1;  out = W_f h + b_f.
 *   x: input feature c of row r at x[r * x_ld + c * x_step] (the coupling layer's identity features, read in
 *   place); ctx [rows, n_ctx] with row pitch ctx_ld or NULL; out [rows, n_out] contiguous.
 *   weights / biases: HOST arrays of device pointers in the order first, (lin_a, lin_b[, gate]) per block, last -
 *   nn.Linear layout weight [out_features, in_features].  hidden <= 128, n_in + n_ctx <= 128, n_blocks <= 4.
 * Backward: recomputes the forward per row tile (only x and ctx are kept), writes g_x [rows, n_in] (may be NULL)
 * and every weight / bias gradient (same order as weights / biases; deterministic row-ordered reduction);
 * workspace: mflow_mlp_workspace_floats(desc) floats.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t rows;
  int32_t n_in, n_ctx, hidden, n_out, n_blocks;
} mflow_mlp_desc;
MFLOW_API int64_t mflow_mlp_workspace_floats(const mflow_mlp_desc* desc);
MFLOW_API int mflow_mlp_forward(const mflow_mlp_desc* desc, const float* x, int64_t x_ld, int64_t x_step, const float* ctx,
                      int64_t ctx_ld, const float* const* weights, const float* const* biases, float* out,
                      void* stream);
MFLOW_API int mflow_mlp_backward(const mflow_mlp_desc* desc, const float* x, int64_t x_ld, int64_t x_step, const float* ctx,
                       int64_t ctx_ld, const float* const* weights, const float* const* biases, const float* g_out,
                       float* g_x, float* const* g_weights, float* const* g_biases, float* workspace, void* stream);

/* ------------------------------------------------------------------------------------
 * (b) Conditioner convolutions on tcgen05 tensor cores (error-compensated 3-product split: fp32-accurate).
 * Replaces nn.Conv2d 3x3 (pad 1) / 1x1 + bias (+ pre-activation ReLU, + residual, + ReLU) of
 * ConvResidualNet / ConvResidualBlock (manifold_flow/nn/resnet.py:75-142).
 *   x [batch, c_in, height, width] fp32 NCHW contiguous, height == width in {4, 8, 16, 32}
 *   w_hi / w_lo [taps, n_pad, k_pad] fp32: w[tap][co][ci] split as w = w_hi + w_lo with w_hi the TF32
 *     truncation (low 13 mantissa bits zero), zero padded to n_pad %% 128 == 0, k_pad %% 32 == 0;
 *     taps = 9 in (dy, dx) row-major order for kernel == 3
 *   y [batch, c_out, height, width] = conv(relu_in ? relu(x) : x) + bias (+ residual), optional ReLU,
 *     then zeroed where gate <= 0 (gate: optional tensor shaped like y - the ReLU mask of the data
 *     gradient, which is the same convolution with the weights transposed and the taps flipped)
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t batch, c_in, c_out, height, width;
  int32_t kernel;   /* 1 or 3 */
  int32_t relu_in;  /* apply ReLU to x on the fly */
  int32_t relu_out; /* apply ReLU to the result */
  int64_t n_pad, k_pad;
} mflow_conv_desc;
MFLOW_API int mflow_conv2d_tf32x3(const mflow_conv_desc* desc, const float* x, const float* w_hi, const float* w_lo,
                        const float* bias, const float* residual, const float* gate, float* y, void* stream);

/* The same convolution with fp16 operand halves (tcgen05 kind::f16: the 11 significant bits of TF32 at twice the
 * tensor-core rate; the default path of the drop-in package).  fp16 has a 5-bit exponent, so both operands carry
 * power-of-two scales:
 *   w_hi / w_lo [taps, n_pad, k_pad] fp16: halves of w[tap][co][ci] * 2^s_co with s_co placing the largest magnitude
 *     of output channel co in [2^13, 2^14); w_unscale [n_pad] fp32 = 2^-s_co (1 for the padding rows)
 *   in_amax : device scalar max |x| (mflow_absmax, or the out_amax of the launch that produced x); the kernel scales
 *     x by the power of two that places it in [2^13, 2^14).  NULL = unscaled (|x| must stay within fp16 range)
 *   out_amax: optional device scalar, ZERO on entry, receives max |y| (bit pattern compared as unsigned)
 * Everything else as mflow_conv2d_tf32x3; the result agrees with it to the last few ulps of fp32. */
MFLOW_API int mflow_conv2d_f16x3(const mflow_conv_desc* desc, const float* x, const void* w_hi, const void* w_lo,
                       const float* w_unscale, const float* in_amax, const float* bias, const float* residual,
                       const float* gate, float* y, float* out_amax, void* stream);

/* The conditioner's FINAL 1x1 convolution of an image coupling layer with the coupling layer's spline in its epilogue
 * (manifold_flow/transforms/coupling.py:108-110 + :264-279 + splines/rational_quadratic.py:16-168):
 * C_out = 32 C_t, so the 32 spline parameters of a transformed channel are one 32-column chunk of the pixel's
 * accumulator row; the epilogue evaluates the spline on them and writes y and partial log-dets - the
 * [B, C_t * 32, H, W] parameter tensor (128 of the 136 bytes the spline moves per element) never exists.
 * Evaluation / sampling only (no backward: the training path keeps the parameters for mflow_rqs_backward).
 *   desc        : the 1x1 convolution (x = the conditioner's last hidden activation [B, c_in, H, W])
 *   sd          : the spline as for mflow_rqs_apply in the planes layout (num_bins == 11, n_chan == c_out / 32,
 *                 n_inner == H * W; the p_* fields are unused); x_full / y_full as x / y there (shifted to the first
 *                 transformed channel, identity channels copied through)
 *   lad_sample  : [B] per-sample log-dets
 *   workspace   : mflow_conv2d_rqs_workspace_bytes(desc) bytes of device scratch (the partial log-dets) */
MFLOW_API int64_t mflow_conv2d_rqs_workspace_bytes(const mflow_conv_desc* desc);
MFLOW_API int mflow_conv2d_rqs_f16x3(const mflow_conv_desc* desc, const float* x, const void* w_hi, const void* w_lo,
                           const float* w_unscale, const float* in_amax, const float* bias, const mflow_rqs_desc* sd,
                           const float* x_full, float* y_full, float* lad_sample, void* workspace, void* stream);

/* max |x| over n floats (x 16-byte aligned) merged into *out, which the caller zero-initialises: the operand
 * scale of the fp16-split kernels for tensors whose producer did not publish it. */
MFLOW_API int mflow_absmax(const float* x, int64_t n, float* out, void* stream);

/* Weight gradient of the same convolutions (autograd of nn.Conv2d in nn/resnet.py:85-142; the reference runs
 * cuDNN / cuBLAS fp32 wgrad kernels):
 *   g_w[co, ci, ky, kx] = sum_{b,h,w} g_y[b, co, h, w] * (relu_in ? relu(x) : x)[b, ci, h + ky - 1, w + kx - 1]
 * written in the reference's weight layout [c_out, c_in, kernel, kernel].  3xTF32 on tcgen05 with the pixels as the
 * reduction dimension, split over CTAs and reduced deterministically.  height == width in {4, 8, 16, 32};
 * kernel == 3 needs c_out <= 112.  desc->relu_out, n_pad, k_pad are ignored.
 * g_bias (optional, [c_out]): the bias gradient sum_{b,h,w} g_y[b, co, h, w], accumulated by the warps that
 * already touch every element of g_y (plain fp32 sums, deterministic).
 * workspace: mflow_conv2d_wgrad_workspace_bytes(desc) bytes, no initialisation needed. */
MFLOW_API int64_t mflow_conv2d_wgrad_workspace_bytes(const mflow_conv_desc* desc);
MFLOW_API int mflow_conv2d_wgrad_tf32x3(const mflow_conv_desc* desc, const float* x, const float* g_y, float* g_w,
                              float* g_bias, void* workspace, void* stream);
/* fp16-operand variant (see mflow_conv2d_f16x3): x_amax / g_amax are device scalars with max |x| / max |g_y|
 * (NULL = unscaled). */
MFLOW_API int mflow_conv2d_wgrad_f16x3(const mflow_conv_desc* desc, const float* x, const float* g_y, const float* x_amax,
                             const float* g_amax, float* g_w, float* g_bias, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------
 * Likelihood epilogues.
 * ---------------------------------------------------------------------------------- */
/* log_prob[o] = logN(u[o,:]; 0, 1) + logN(clamp(v[o,:], +-clip); 0, eps^2) + sum of the addends.
 * StandardNormal / RescaledNormal (distributions/normal.py:18-23, 52-59) and
 * ManifoldFlow._log_prob "pie"/"mf-fixed-manifold" (flows/manifold_flow.py:125-128,149-152).
 * v may be NULL (n_v = 0); clip <= 0 disables clipping; add0/add1 optional [O]. */
MFLOW_API int mflow_gauss_logprob(const float* u, int64_t n_u, int64_t u_stride, const float* v, int64_t n_v,
                        int64_t v_stride, float eps, float clip, const float* add0, const float* add1,
                        float* log_prob, int64_t rows, void* stream);
MFLOW_API int mflow_gauss_logprob_backward(const float* u, int64_t n_u, int64_t u_stride, const float* v, int64_t n_v,
                                 int64_t v_stride, float eps, float clip, const float* g_log_prob,
                                 float* g_u, float* g_v, int64_t rows, void* stream);

/* (d) logdet[o] = log det (J_o^T J_o) for J [O, D, n] row-major, n <= 32, by Cholesky of the Gram
 * matrix: ManifoldFlow._log_prob "mf" (flows/manifold_flow.py:140-147: bmm + slogdet). */
MFLOW_API int mflow_jtj_logdet(const float* jac, float* logdet, int64_t rows, int32_t d, int32_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFLOW_B200_H */
