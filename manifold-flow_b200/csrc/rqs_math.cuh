// Per-element rational-quadratic spline arithmetic, shared by every layout's kernels.
//
// The value path follows the reference's fp32 operation ORDER
// (manifold_flow/transforms/splines/rational_quadratic.py:102-168): every
// multiply/add that the reference performs as a separate ATen op is written with __fmul_rn /
// __fadd_rn / __fsub_rn so nvcc cannot contract it into an FMA, divisions and square roots are the
// IEEE-rounded ones, and exp/log/log1p are the same libdevice functions torch's own CUDA kernels
// call.  The same functions are used by the forward and the backward kernels, so both always agree
// on the bin an element falls in.
//
// Two arithmetic modes (compile-time):
//   default               - same formulas, but the knot construction lets the compiler fuse multiply-adds
//                           (min + c*p, span*c + lo), scales the logits with one multiply and evaluates the
//                           softmax exponentials with ex2.approx: knots move by at most ~2 ulp, which changes
//                           an output only if x sits within ~1e-7 of a knot (and then by ~1e-7: the spline is
//                           C^1).  About 40% fewer instructions; measured accuracy against fp64 is
//                           indistinguishable from the reference's own fp32 path (tools/spline_error_stats.py).
//   -DMFLOW_RQS_EXACT_ORDER - every ATen op of the reference as a separately rounded operation (no
//                           contraction, IEEE divisions, compensated exponentials): ~55% of the outputs are
//                           bit-identical to the CPU reference.
#pragma once
#include "mflow_common.cuh"

namespace mflow {

struct RqsConsts {
  float tail;        // T
  float lo;          // -T
  float span;        // float(T - (-T))
  float min_w, min_h, min_d;
  float cfac_w;      // float(1 - min_w * K)   (Python double arithmetic, rounded once)
  float cfac_h;      // float(1 - min_h * K)
  float edge;        // float(log(exp(1 - min_d) - 1)), the boundary derivative logit (:39)
  float divisor;     // widths / heights logits are divided by this (coupling.py:506-511)
  float rcp_divisor; // float(1 / divisor)
  int use_divisor;
  // box mode only (rational_quadratic_spline on [left, right] x [bottom, top], :67-168): the x axis uses lo / tail /
  // span above as left / right / right - left, the y axis these
  float lo_y, hi_y, span_y;
};

// Correctly rounded x / d for a constant d with pre-rounded reciprocal r (one Newton step on the
// quotient; exact IEEE result for every normal x when d is a small integer such as 10).
__device__ __forceinline__ float div_const(float x, float d, float r) {
  float q = __fmul_rn(x, r);
  float rem = __fmaf_rn(-q, d, x);
  return __fmaf_rn(rem, r, q);
}

// x / s for a per-element denominator with r = 1/s already rounded (softmax normalisation).
__device__ __forceinline__ float div_by(float x, float s, float r) {
  float q = __fmul_rn(x, r);
  float rem = __fmaf_rn(-q, s, x);
  return __fmaf_rn(rem, r, q);
}

// exp(x) for x <= 0 (softmax arguments after the max subtraction): ex2.approx on a compensated product.
// x * log2(e) is split as t + r with t = fl(x * L2E_hi), r the rounding error of that product plus the
// low part of log2(e); exp2(t + r) = exp2(t) * (1 + r ln 2 + ...).  About 3 ulp, 6 instructions instead of
// libdevice expf's ~14 (no special-case handling is needed on this argument range).
__device__ __forceinline__ float exp_nonpos(float x) {
  const float l2e_hi = 1.4426950216293335f, l2e_lo = 1.9259629911266175e-8f;
  const float t = __fmul_rn(x, l2e_hi);
  const float r = __fmaf_rn(x, l2e_lo, __fmaf_rn(x, l2e_hi, -t));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return __fmaf_rn(e, __fmul_rn(r, 0.6931471805599453f), e);
}

// softplus with torch's default beta = 1, threshold = 20 (F.softplus)
__device__ __forceinline__ float softplus20(float u) { return u > 20.f ? u : log1pf(expf(u)); }
__device__ __forceinline__ float sigmoid20(float u) { return u > 20.f ? 1.f : 1.f / (1.f + expf(-u)); }

// Knots of one axis: softmax -> floor -> cumulative sum -> rescale (mul, then add) -> exact ends.
//   rational_quadratic.py:102-108 / :113-119.  frac[] receives the softmax probabilities.
template <int K>
__device__ __forceinline__ void rqs_axis(const float (&u)[K], float lo, float hi, float span, float min_size,
                                         float cfac, float (&knots)[K + 1], float (&frac)[K]) {
  float m = u[0];
#pragma unroll
  for (int j = 1; j < K; ++j) m = fmaxf(m, u[j]);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
#ifdef MFLOW_RQS_EXACT_ORDER
    frac[j] = exp_nonpos(__fsub_rn(u[j], m));
#else
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(frac[j]) : "f"((u[j] - m) * 1.4426950408889634f));
#endif
    s = __fadd_rn(s, frac[j]);
  }
  const float r = __frcp_rn(s);
  float c = 0.f;
  knots[0] = lo;
#pragma unroll
  for (int j = 0; j < K; ++j) {
#ifdef MFLOW_RQS_EXACT_ORDER
    frac[j] = div_by(frac[j], s, r);
    c = __fadd_rn(c, __fadd_rn(min_size, __fmul_rn(cfac, frac[j])));
    knots[j + 1] = __fadd_rn(__fmul_rn(span, c), lo);
#else
    frac[j] = frac[j] * r;
    c += fmaf(cfac, frac[j], min_size);
    knots[j + 1] = fmaf(span, c, lo);
#endif
  }
  knots[K] = hi;
}

// Bin search, utils/various.py:472-474: bin = #(v >= knot_j) - 1 with the last knot bumped by 1e-6.
// For lo <= v <= hi the bumped last knot is never reached, and knots are non-decreasing, so the bin
// is the largest j <= K-1 with v >= knot_j.
template <int K>
__device__ __forceinline__ int rqs_search(const float (&knots)[K + 1], float v, float& k_lo, float& k_hi) {
  int k = 0;
  k_lo = knots[0];
  k_hi = knots[1];
#pragma unroll
  for (int j = 1; j < K; ++j) {
    if (v >= knots[j]) {
      k = j;
      k_lo = knots[j];
      k_hi = knots[j + 1];
    }
  }
  return k;
}

template <int K>
__device__ __forceinline__ void rqs_pick(const float (&knots)[K + 1], int k, float& k_lo, float& k_hi) {
  k_lo = knots[0];
  k_hi = knots[1];
#pragma unroll
  for (int j = 1; j < K; ++j) {
    if (k == j) {
      k_lo = knots[j];
      k_hi = knots[j + 1];
    }
  }
}

// Everything an evaluation needs about the bin the element fell in.
template <int K>
struct RqsBin {
  int k;
  float xk, wk, yk, hk, sk;  // left knot, bin width, bottom knot, bin height, slope h/w
  float uk, uk1;             // derivative logits at both knots (edge constant at the ends)
  float dk, dk1;             // knot derivatives min_d + softplus(logit)
  float sig_w[K], sig_h[K];  // softmax probabilities (backward only)
};

// p holds the element's raw parameters [widths | heights | derivatives]: NP = 3K-1 for the unconstrained spline
// with linear tails (K-1 interior derivative logits, the two boundary ones are the constant c.edge, :38-41), or
// NP = 3K+1 for the spline on a box ("BOX": all K+1 knot derivatives given, separate x and y ranges, :67-168).
template <int K, bool INV, int NP = 3 * K - 1>
__device__ __forceinline__ void rqs_locate(float v, const float (&p)[NP], const RqsConsts& c, RqsBin<K>& b) {
  constexpr bool BOX = NP == 3 * K + 1;
  static_assert(NP == 3 * K - 1 || NP == 3 * K + 1, "parameter count");
  float uw[K], uh[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
#ifdef MFLOW_RQS_EXACT_ORDER
    uw[j] = c.use_divisor ? div_const(p[j], c.divisor, c.rcp_divisor) : p[j];
    uh[j] = c.use_divisor ? div_const(p[K + j], c.divisor, c.rcp_divisor) : p[K + j];
#else
    uw[j] = c.use_divisor ? p[j] * c.rcp_divisor : p[j];
    uh[j] = c.use_divisor ? p[K + j] * c.rcp_divisor : p[K + j];
#endif
  }
  float kw[K + 1], kh[K + 1];
  rqs_axis<K>(uw, c.lo, c.tail, c.span, c.min_w, c.cfac_w, kw, b.sig_w);
  rqs_axis<K>(uh, BOX ? c.lo_y : c.lo, BOX ? c.hi_y : c.tail, BOX ? c.span_y : c.span, c.min_h, c.cfac_h, kh, b.sig_h);
  float x_lo, x_hi, y_lo, y_hi;
  if (INV) {
    b.k = rqs_search<K>(kh, v, y_lo, y_hi);
    rqs_pick<K>(kw, b.k, x_lo, x_hi);
  } else {
    b.k = rqs_search<K>(kw, v, x_lo, x_hi);
    rqs_pick<K>(kh, b.k, y_lo, y_hi);
  }
  b.xk = x_lo;
  b.wk = __fsub_rn(x_hi, x_lo);  // widths re-derived from the rounded knots (:109)
  b.yk = y_lo;
  b.hk = __fsub_rn(y_hi, y_lo);
  b.sk = __fdiv_rn(b.hk, b.wk);  // :131
  if constexpr (BOX) {
    b.uk = p[2 * K];
    b.uk1 = p[2 * K + 1];
#pragma unroll
    for (int j = 1; j < K; ++j) {
      if (b.k == j) { b.uk = p[2 * K + j]; b.uk1 = p[2 * K + j + 1]; }
    }
  } else {
    b.uk = c.edge;
    b.uk1 = c.edge;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) {  // padded logits: [edge, D_0 .. D_{K-2}, edge]
      if (b.k == j + 1) b.uk = p[2 * K + j];
      if (b.k == j) b.uk1 = p[2 * K + j];
    }
  }
  b.dk = __fadd_rn(c.min_d, softplus20(b.uk));  // :111
  b.dk1 = __fadd_rn(c.min_d, softplus20(b.uk1));
}

__device__ __forceinline__ float rqs_logdet_terms(float sk, float dk, float dk1, float t, float tt, float den) {
  // log(s^2 (d1 t^2 + 2 s t(1-t) + d0 (1-t)^2)) - 2 log(den)     (:165-166 / :153-154)
  const float omt = __fsub_rn(1.f, t);
  const float q = __fadd_rn(__fadd_rn(__fmul_rn(dk1, __fmul_rn(t, t)), __fmul_rn(__fmul_rn(2.f, sk), tt)),
                            __fmul_rn(dk, __fmul_rn(omt, omt)));
  const float dnum = __fmul_rn(__fmul_rn(sk, sk), q);
  return __fsub_rn(logf(dnum), __fmul_rn(2.f, logf(den)));
}

// Evaluate one element.  Outside [-T, T] (closed): identity, logabsdet 0, bin -1 (:31,:43-44).
// (BOX: outside the domain of the searched axis; the host wrapper has raised InputOutsideDomain before.)
template <int K, bool INV, int NP = 3 * K - 1>
__device__ __forceinline__ void rqs_eval(float v, const float (&p)[NP], const RqsConsts& c, float& out, float& lad,
                                         int& bin) {
  constexpr bool BOX = NP == 3 * K + 1;
  if (!(v >= ((BOX && INV) ? c.lo_y : c.lo) && v <= ((BOX && INV) ? c.hi_y : c.tail))) {
    out = v;
    lad = 0.f;
    bin = -1;
    return;
  }
  RqsBin<K> b;
  rqs_locate<K, INV, NP>(v, p, c, b);
  bin = b.k;
  const float a_term = __fsub_rn(__fadd_rn(b.dk, b.dk1), __fmul_rn(2.f, b.sk));  // d_k + d_{k+1} - 2 s
  if (INV) {  // :139-156
    const float dy = __fsub_rn(v, b.yk);
    const float qa = __fadd_rn(__fmul_rn(dy, a_term), __fmul_rn(b.hk, __fsub_rn(b.sk, b.dk)));
    const float qb = __fsub_rn(__fmul_rn(b.hk, b.dk), __fmul_rn(dy, a_term));
    const float qc = __fmul_rn(-b.sk, dy);
    float disc = __fsub_rn(__fmul_rn(qb, qb), __fmul_rn(__fmul_rn(4.f, qa), qc));
    disc = fmaxf(disc, 0.f);
    const float root = __fdiv_rn(__fmul_rn(2.f, qc), __fsub_rn(-qb, __fsqrt_rn(disc)));
    out = __fadd_rn(__fmul_rn(root, b.wk), b.xk);
    const float tt = __fmul_rn(root, __fsub_rn(1.f, root));
    const float den = __fadd_rn(b.sk, __fmul_rn(a_term, tt));
    lad = -rqs_logdet_terms(b.sk, b.dk, b.dk1, root, tt, den);
  } else {  // :158-168
    const float t = __fdiv_rn(__fsub_rn(v, b.xk), b.wk);
    const float tt = __fmul_rn(t, __fsub_rn(1.f, t));
    const float num = __fmul_rn(b.hk, __fadd_rn(__fmul_rn(b.sk, __fmul_rn(t, t)), __fmul_rn(b.dk, tt)));
    const float den = __fadd_rn(b.sk, __fmul_rn(a_term, tt));
    out = __fadd_rn(b.yk, __fdiv_rn(num, den));
    lad = rqs_logdet_terms(b.sk, b.dk, b.dk1, t, tt, den);
  }
}

// Backward of one element: gradient w.r.t. the input value and all 3K-1 raw parameters, given the
// gradients g_out (w.r.t. the output value) and g_lad (w.r.t. the element's logabsdet).
//   v_in  : input of the apply call (x for forward, y for inverse)
//   v_out : output of the apply call (only used when INV: the x the inverse produced)
// Inverse direction by implicit differentiation of y = F(x; params), lad_inv = -lad_F(x; params).
// store(j, value) is called exactly once for every parameter index j in [0, 3K-1).
template <int K, bool INV, int NP = 3 * K - 1, typename Store>
__device__ __forceinline__ void rqs_grad(float v_in, float v_out, const float (&p)[NP], const RqsConsts& c,
                                         float g_out, float g_lad, float& g_in, Store&& store) {
  constexpr bool BOX = NP == 3 * K + 1;
  if (!(v_in >= ((BOX && INV) ? c.lo_y : c.lo) && v_in <= ((BOX && INV) ? c.hi_y : c.tail))) {
    g_in = g_out;
#pragma unroll
    for (int j = 0; j < NP; ++j) store(j, 0.f);
    return;
  }
  RqsBin<K> b;
  rqs_locate<K, INV, NP>(v_in, p, c, b);
  const float xdom = INV ? v_out : v_in;
  const float inv_w = 1.f / b.wk;
  const float t = (xdom - b.xk) * inv_w, u = 1.f - t, tt = t * u, umt = u - t;
  const float A = b.dk + b.dk1 - 2.f * b.sk;
  const float R = b.sk * t * t + b.dk * tt;
  const float num = b.hk * R;
  const float den = b.sk + A * tt;
  const float Q = b.dk1 * t * t + 2.f * b.sk * tt + b.dk * u * u;
  const float dnum = b.sk * b.sk * Q;
  const float inv_den = 1.f / den;
  const float Q_t = 2.f * (b.dk1 * t + b.sk * umt - b.dk * u);
  const float den_t = A * umt;
  float ga, gb;  // weights of grad(y) and grad(lad) in the total gradient
  if (INV) {
    const float lad_t = Q_t / Q - 2.f * den_t * inv_den;
    const float fprime = dnum * inv_den * inv_den;  // dy/dx
    gb = -g_lad;
    ga = -(g_out + gb * lad_t * inv_w) / fprime;
  } else {
    ga = g_out;
    gb = g_lad;
  }
  const float g_num = ga * inv_den;
  const float g_den = -ga * num * inv_den * inv_den - 2.f * gb * inv_den;
  const float g_dnum = gb / dnum;
  const float g_Q = g_dnum * b.sk * b.sk;
  const float g_A = g_den * tt;
  const float g_R = g_num * b.hk;
  float g_sk = g_dnum * 2.f * b.sk * Q + g_Q * 2.f * tt + g_den - 2.f * g_A + g_R * t * t;
  const float g_dk = g_Q * u * u + g_A + g_R * tt;
  const float g_dk1 = g_Q * t * t + g_A;
  const float g_t = g_Q * Q_t + g_den * den_t + g_R * (2.f * b.sk * t + b.dk * umt);
  const float g_x = g_t * inv_w;
  const float g_hk = g_num * R + g_sk * inv_w;
  const float g_wk = -g_sk * b.sk * inv_w - g_x * t;
  const float g_xk = -g_x;
  const float g_yk = ga;
  g_in = INV ? -ga : g_x;

  const float inv_div = c.use_divisor ? c.rcp_divisor : 1.f;
  {  // widths: knot_k = lo + span * sum_{i<k} w~_i, w_k = span * w~_k, w~ = min + cfac * softmax
    const float lt = c.cfac_w * c.span * g_xk, eq = c.cfac_w * c.span * g_wk;
    float below = 0.f, at = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      below += (j < b.k) ? b.sig_w[j] : 0.f;
      at = (j == b.k) ? b.sig_w[j] : at;
    }
    const float dot = lt * below + eq * at;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float gs = (j < b.k ? lt : 0.f) + (j == b.k ? eq : 0.f);
      store(j, b.sig_w[j] * (gs - dot) * inv_div);
    }
  }
  {  // heights
    const float span_y = BOX ? c.span_y : c.span;
    const float lt = c.cfac_h * span_y * g_yk, eq = c.cfac_h * span_y * g_hk;
    float below = 0.f, at = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      below += (j < b.k) ? b.sig_h[j] : 0.f;
      at = (j == b.k) ? b.sig_h[j] : at;
    }
    const float dot = lt * below + eq * at;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float gs = (j < b.k ? lt : 0.f) + (j == b.k ? eq : 0.f);
      store(K + j, b.sig_h[j] * (gs - dot) * inv_div);
    }
  }
  const float gd0 = g_dk * sigmoid20(b.uk), gd1 = g_dk1 * sigmoid20(b.uk1);
  if constexpr (BOX) {
#pragma unroll
    for (int j = 0; j <= K; ++j) store(2 * K + j, (b.k == j) ? gd0 : ((b.k + 1 == j) ? gd1 : 0.f));
  } else {
#pragma unroll
    for (int j = 0; j < K - 1; ++j) store(2 * K + j, (b.k == j + 1) ? gd0 : ((b.k == j) ? gd1 : 0.f));
  }
}

}  // namespace mflow
