// Fused rational-quadratic spline kernels (kernel (a) of BASELINE.json:north_star).
//
// One launch replaces ~60 ATen launches of the reference
// (manifold_flow/transforms/splines/rational_quadratic.py:16-168, utils/various.py:472-474,
// transforms/coupling.py:264-279,501-511): parameter slicing and 1/sqrt(hidden) pre-scale, both
// softmaxes, knot construction, bin search, the RQ map (or its quadratic-root inverse), linear
// tails, the per-sample log-det reduction and the coupling layer's identity pass-through.
//
// Two data layouts (see include/mflow_b200.h):
//   "planes": parameters [O, C_t*P, I] - a warp reads 32 consecutive pixels of one parameter
//             plane per load instruction: 128 B, fully coalesced, P independent loads in flight
//             per thread.  This is also the generic strided kernel.
//   "rows"  : parameters [O*C_t, P] contiguous - a CTA stages its rows through shared memory with
//             coalesced loads and an odd row pitch (conflict-free per-thread row reads).
// HBM-bound by design: 4*(P+2) algorithmic bytes per element forward, 4*(2P+3) backward.
#include <stdarg.h>

#include "rqs_math.cuh"

namespace mflow {

struct RqsArgs {
  const float* x;
  const float* x_out;  // backward + inverse only
  const float* params;
  float* y;            // forward: output; backward: g_in
  float* lad_elem;     // forward: optional; backward: unused
  float* lad_sample;
  const float* lad_base;
  int32_t* bins;
  float* partial;      // [O, chunks]
  unsigned int* counters;  // [O]
  // backward
  const float* g_out;
  const float* g_lad_sample;
  const float* g_lad_elem;
  float* g_params;
  float* g_params_amax;   // optional device scalar receiving max |g_params| of this launch
  int64_t n_outer, n_chan, n_inner;
  int64_t x_so, x_sc, y_so, y_sc, p_so, p_sc, p_sp, p_si;
  int64_t n_id_chan, id_off, id_sc;
  int64_t p_hw;        // rows x planes layout: pixels per parameter image (power of two), else 0
  int64_t t_off;       // rows x planes layout: index of the first transformed feature
  int tile_shift;      // rows x planes layout: log2(rows per CTA)
  uint32_t d_mul, d_shift;   // idx / x_so as a multiply-shift (rows x planes layout)
  int p_hw_shift;      // log2(p_hw)
  int chunks;
  uint32_t inner_mul, inner_shift;  // e / n_inner == umulhi(e, inner_mul) >> inner_shift for 0 <= e < 2^31 (planes kernels)
  RqsConsts c;
};

// exact unsigned division by an invariant divisor (Granlund-Montgomery, 32-bit, valid for dividends < 2^31)
__device__ __forceinline__ uint32_t fast_div(uint32_t e, uint32_t mul, uint32_t shift) {
  return mul == 0 ? (e >> shift) : (__umulhi(e, mul) >> shift);
}

// max |value| of a kernel's parameter gradients -> device scalar (zero-initialised by the caller; non-negative floats
// order like unsigned integers): the operand scale of the fp16-split convolution that consumes them, without an extra
// pass over the tensor.  All 32 lanes of the warp call.
__device__ __forceinline__ void publish_amax(float* slot, float m) {
  const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  if ((threadIdx.x & 31) == 0 && w > *reinterpret_cast<volatile uint32_t*>(slot)) atomicMax(reinterpret_cast<unsigned int*>(slot), w);
}

constexpr int kPlaneThreads = 256;
constexpr int64_t kCounterBytes = 65536 * 4;  // one arrival counter per sample (planes layout: O <= 65535)

// ---------------------------------------------------------------------------------------------
// planes / generic layout, forward or inverse
// grid = (chunks, O); CTA (chunk, o) walks elements [chunk*256, ...) of sample o with stride
// chunks*256, so consecutive threads touch consecutive pixels of every parameter plane.
// ---------------------------------------------------------------------------------------------
// INNER > 0: the standard planes layout with a compile-time plane size (p_si == 1, p_sp == INNER, p_sc ==
// P * INNER): parameter loads become LDG [base + immediate] with no per-load address arithmetic.
template <int K, bool INV, int INNER>
__global__ void __launch_bounds__(kPlaneThreads, 6) rqs_planes_apply_kernel(const RqsArgs a) {
  constexpr int P = 3 * K - 1;
  __shared__ float red[kPlaneThreads / 32];
  __shared__ int s_last;
  const int64_t o = blockIdx.y;
  // per-sample extents and strides fit in 32 bits (checked on the host): 64-bit math only for the sample base
  const uint32_t n_inner = INNER ? INNER : (uint32_t)a.n_inner, per_sample = (uint32_t)a.n_chan * n_inner;
  const uint32_t p_sc = INNER ? P * INNER : (uint32_t)a.p_sc, p_si = INNER ? 1 : (uint32_t)a.p_si;
  const uint32_t p_sp = INNER ? INNER : (uint32_t)a.p_sp;
  const uint32_t x_sc = (uint32_t)a.x_sc, y_sc = (uint32_t)a.y_sc;
  const float* xo = a.x + o * a.x_so;
  float* yo = a.y + o * a.y_so;
  const float* po = a.params + o * a.p_so;
  float lad_acc = 0.f;
  for (uint32_t e = blockIdx.x * kPlaneThreads + threadIdx.x; e < per_sample; e += gridDim.x * kPlaneThreads) {
    const uint32_t ch = INNER ? e / INNER : fast_div(e, a.inner_mul, a.inner_shift), i = e - ch * n_inner;
    const float* pe = po + (ch * p_sc + i * p_si);
    float p[P];
#pragma unroll
    for (int j = 0; j < P; ++j) p[j] = ldg_stream(pe + j * p_sp);
    const float v = ldg_stream(xo + (ch * x_sc + i));
    float out, lad;
    int bin;
    rqs_eval<K, INV>(v, p, a.c, out, lad, bin);
    stg_stream(yo + (ch * y_sc + i), out);
    if (a.lad_elem) stg_stream(a.lad_elem + o * per_sample + e, lad);
    if (a.bins) a.bins[o * per_sample + e] = bin;
    lad_acc += lad;
  }
  // identity half of the coupling layer: plain copy (coupling.py:116-118)
  const uint32_t n_id = (uint32_t)(a.n_id_chan * a.n_inner);
  for (uint32_t e = blockIdx.x * kPlaneThreads + threadIdx.x; e < n_id; e += gridDim.x * kPlaneThreads) {
    const uint32_t ch = INNER ? e / INNER : fast_div(e, a.inner_mul, a.inner_shift), i = e - ch * n_inner;
    const int64_t off = a.id_off + (int64_t)(ch * (uint32_t)a.id_sc + i);
    stg_stream(yo + off, ldg_stream(xo + off));
  }
  if (a.lad_sample == nullptr) return;
  // deterministic per-sample reduction: CTA partial -> last-arriving CTA sums partials in order
  const float bs = block_sum<kPlaneThreads>(lad_acc, red);
  if (gridDim.x == 1) {
    if (threadIdx.x == 0) a.lad_sample[o] = (a.lad_base ? a.lad_base[o] : 0.f) + bs;
    return;
  }
  if (threadIdx.x == 0) {
    a.partial[o * gridDim.x + blockIdx.x] = bs;
    __threadfence();
    const unsigned int prev = atomicAdd(&a.counters[o], 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    float s = 0.f;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += 32) s += __ldcg(a.partial + o * gridDim.x + j);
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      a.lad_sample[o] = (a.lad_base ? a.lad_base[o] : 0.f) + s;
      a.counters[o] = 0u;  // leave the workspace ready for the next launch
    }
  }
}

// planes / generic layout, backward
template <int K, bool INV, int INNER>
__global__ void __launch_bounds__(kPlaneThreads, 2) rqs_planes_backward_kernel(const RqsArgs a) {
  constexpr int P = 3 * K - 1;
  const int64_t o = blockIdx.y;
  const uint32_t n_inner = INNER ? INNER : (uint32_t)a.n_inner, per_sample = (uint32_t)a.n_chan * n_inner;
  const uint32_t p_sc = INNER ? P * INNER : (uint32_t)a.p_sc, p_si = INNER ? 1 : (uint32_t)a.p_si;
  const uint32_t x_sc = (uint32_t)a.x_sc, y_sc = (uint32_t)a.y_sc;
  const float* xo = a.x + o * a.x_so;
  const float* go = a.g_out ? a.g_out + o * a.y_so : nullptr;
  float* gio = a.y + o * a.x_so;  // g_in uses the input tensor's indexing
  const float* po = a.params + o * a.p_so;
  float* gpo = a.g_params + o * a.p_so;
  const float gl_s = a.g_lad_sample ? a.g_lad_sample[o] : 0.f;
  const uint32_t p_sp = INNER ? INNER : (uint32_t)a.p_sp;
  float gp_max = 0.f;
  for (uint32_t e = blockIdx.x * kPlaneThreads + threadIdx.x; e < per_sample; e += gridDim.x * kPlaneThreads) {
    const uint32_t ch = INNER ? e / INNER : fast_div(e, a.inner_mul, a.inner_shift), i = e - ch * n_inner;
    const uint32_t poff = ch * p_sc + i * p_si;
    const float* pe = po + poff;
    float p[P];
#pragma unroll
    for (int j = 0; j < P; ++j) p[j] = ldg_stream(pe + j * p_sp);
    const float v = ldg_stream(xo + (ch * x_sc + i));
    const float vo = INV ? ldg_stream(a.x_out + o * a.y_so + (ch * y_sc + i)) : 0.f;
    const float g = go ? ldg_stream(go + (ch * y_sc + i)) : 0.f;
    const float gl = gl_s + (a.g_lad_elem ? ldg_stream(a.g_lad_elem + o * per_sample + e) : 0.f);
    float g_in;
    float* gpe = gpo + poff;
    const uint32_t sp = p_sp;
    rqs_grad<K, INV>(v, vo, p, a.c, g, gl, g_in, [&](int j, float val) { stg_stream(gpe + j * sp, val); gp_max = fmaxf(gp_max, fabsf(val)); });
    stg_stream(gio + (ch * x_sc + i), g_in);
  }
  if (a.g_params_amax != nullptr) publish_amax(a.g_params_amax, gp_max);
  const uint32_t n_id = (uint32_t)(a.n_id_chan * a.n_inner);
  for (uint32_t e = blockIdx.x * kPlaneThreads + threadIdx.x; e < n_id; e += gridDim.x * kPlaneThreads) {
    const uint32_t ch = INNER ? e / INNER : fast_div(e, a.inner_mul, a.inner_shift), i = e - ch * n_inner;
    const int64_t off = a.id_off + (int64_t)(ch * (uint32_t)a.id_sc + i);
    stg_stream(gio + off, go ? ldg_stream(go + off) : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// rows layout (vectors): parameters [O*C_t, P] contiguous.  A CTA owns G = 128 / C_t whole samples
// (G*C_t <= 128 elements, one per thread), stages their G*C_t*P floats through shared memory with
// coalesced loads, and reduces each sample's logabsdet sequentially in a fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int kRowThreads = 128;

template <int K, bool INV, bool BWD>
__global__ void __launch_bounds__(kRowThreads) rqs_rows_kernel(const RqsArgs a, int samples_per_cta) {
  constexpr int P = 3 * K - 1;
  constexpr int PS = P | 1;  // odd pitch: thread t reads smem[t*PS + j] without bank conflicts
  __shared__ float s_par[kRowThreads * PS];
  __shared__ float s_lad[kRowThreads];
  const int ct = (int)a.n_chan;
  const int64_t o0 = (int64_t)blockIdx.x * samples_per_cta;
  const int n_samp = (int)min((int64_t)samples_per_cta, a.n_outer - o0);
  const int n_elem = n_samp * ct;
  const int total = n_elem * P;
  const float* src = a.params + o0 * a.p_so;
  for (int idx = threadIdx.x; idx < total; idx += kRowThreads) {
    const int r = idx / P, col = idx - r * P;
    s_par[r * PS + col] = ldg_stream(src + idx);
  }
  __syncthreads();
  const int t = threadIdx.x;
  const bool active = t < n_elem;
  const int so = t / ct, ch = t - so * ct;
  const int64_t o = o0 + so;
  float p[P];
  float lad = 0.f, gp_max = 0.f;
  if (active) {
#pragma unroll
    for (int j = 0; j < P; ++j) p[j] = s_par[t * PS + j];
    const float v = a.x[o * a.x_so + ch * a.x_sc];
    if (!BWD) {
      float out;
      int bin;
      rqs_eval<K, INV>(v, p, a.c, out, lad, bin);
      a.y[o * a.y_so + ch * a.y_sc] = out;
      if (a.lad_elem) a.lad_elem[o * ct + ch] = lad;
      if (a.bins) a.bins[o * ct + ch] = bin;
    } else {
      const float vo = INV ? a.x_out[o * a.y_so + ch * a.y_sc] : 0.f;
      const float g = a.g_out ? a.g_out[o * a.y_so + ch * a.y_sc] : 0.f;
      const float gl = (a.g_lad_sample ? a.g_lad_sample[o] : 0.f) + (a.g_lad_elem ? a.g_lad_elem[o * ct + ch] : 0.f);
      float g_in;
      float* row = s_par + t * PS;  // the row is private to this thread: reuse it for the gradients
      rqs_grad<K, INV>(v, vo, p, a.c, g, gl, g_in, [&](int j, float val) { row[j] = val; gp_max = fmaxf(gp_max, fabsf(val)); });
      a.y[o * a.x_so + ch * a.x_sc] = g_in;
    }
  }
  // identity features of these samples (coupling merge / its gradient)
  for (int idx = t; idx < n_samp * (int)a.n_id_chan; idx += kRowThreads) {
    const int s2 = idx / (int)a.n_id_chan, c2 = idx - s2 * (int)a.n_id_chan;
    if (!BWD) {
      a.y[(o0 + s2) * a.y_so + a.id_off + c2 * a.id_sc] = a.x[(o0 + s2) * a.x_so + a.id_off + c2 * a.id_sc];
    } else {
      a.y[(o0 + s2) * a.x_so + a.id_off + c2 * a.id_sc] =
          a.g_out ? a.g_out[(o0 + s2) * a.y_so + a.id_off + c2 * a.id_sc] : 0.f;
    }
  }
  if (!BWD) {
    if (a.lad_sample == nullptr) return;
    s_lad[t] = lad;
    __syncthreads();
    if (active && ch == 0) {
      float s = 0.f;
      for (int j = 0; j < ct; ++j) s += s_lad[t + j];
      a.lad_sample[o] = (a.lad_base ? a.lad_base[o] : 0.f) + s;
    }
  } else {
    if (a.g_params_amax != nullptr) publish_amax(a.g_params_amax, gp_max);
    __syncthreads();
    float* dst = a.g_params + o0 * a.p_so;
    for (int idx = threadIdx.x; idx < total; idx += kRowThreads) {
      const int r = idx / P, col = idx - r * P;
      stg_stream(dst + idx, s_par[r * PS + col]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// rows x planes layout: x / y are vectors [O, D] (n_inner == 1) but the parameters are the conditioner's NCHW
// output [O / HW, C_t * P, HW] - ResidualNet evaluated on the tensor-core convolution kernels keeps its batch rows
// as the pixels of HW-pixel images (manifold_flow/b200/conditioner.py), and transposing its [O, C_t * P] result
// back to rows (and the parameter gradient to planes again) cost more HBM traffic than the spline itself.
// thread = (row, channel group), a CTA = 32 rows x 8 groups (one warp per group): a warp reads 32 consecutive pixels of
// every parameter plane (128-byte coalesced, P independent loads in flight per thread), and a thread handles every 8th
// transformed feature of its row.  (First version: 128 rows x 2 groups - at 65 536 rows that is a single wave of 24 warps
// per SM with ten elements per thread in sequence: 1.9 TB/s, latency bound, ncu: 30 % occupancy, 45 % long-scoreboard.)  The tensor the launch WRITES (y, or g_in) is the tile [128 rows, D] of
// its source (x, or g_out) with the transformed features replaced, so the CTA stages that tile through shared memory:
// whole rows in and out with coalesced accesses instead of 4-byte stores 4 D bytes apart (the identity features and
// their gradient pass through on the way).  A row's log-det is the sum of its channel groups' partial sums in a
// fixed order.  x, y, g_out, g_in point at feature 0 of row 0; transformed feature c sits at t_off + c * x_sc.
// ---------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256, kTileMaxD = 92;
// Rows per CTA (a power of two, 32 .. 256; the CTA's 256 threads are rows x channel groups): as many channel groups as
// there are transformed features to go round (at most 8, so that a warp still reads 32 consecutive pixels of a plane),
// fewer when the row tile would not fit 48 KB of shared memory... which cannot happen for D <= 92 at 32 rows.
static int tile_rows_shift(int64_t n_chan, int64_t d_full) {
  int groups = 1;
  while (groups < 8 && groups * 2 <= n_chan) groups *= 2;
  while (groups < 8 && (size_t)(kTileThreads / groups) * ((size_t)d_full | 1) * sizeof(float) > 48 * 1024) groups *= 2;
  int shift = 0;
  while ((kTileThreads / groups) >> (shift + 1)) ++shift;
  return shift;   // log2(rows per CTA)
}

template <int K, bool INV, bool BWD>
__global__ void __launch_bounds__(kTileThreads, BWD ? 2 : 6) rqs_rowtile_kernel(const RqsArgs a) {
  constexpr int P = 3 * K - 1;
  extern __shared__ float s_tile[];                     // [rows][DS]
  __shared__ float s_lad[kTileThreads];                 // [group][row]
  const int tr = 1 << a.tile_shift, groups = kTileThreads >> a.tile_shift;
  const int r = threadIdx.x & (tr - 1), cg = threadIdx.x >> a.tile_shift;
  const int ct = (int)a.n_chan, D = (int)a.x_so, DS = D | 1;
  const int t_off = (int)a.t_off, x_sc = (int)a.x_sc;
  const int64_t o0 = (int64_t)blockIdx.x * tr;
  const int n_rows = (int)min((int64_t)tr, a.n_outer - o0);
  {
    // the tile's rows are one contiguous range of the source: flat, fully coalesced copy; idx -> (row, column) by a
    // multiply-shift division
    const float* src = BWD ? a.g_out : a.x;
    const float* base = src ? src + o0 * D : nullptr;
    for (uint32_t idx = threadIdx.x; idx < (uint32_t)(n_rows * D); idx += kTileThreads) {
      const uint32_t rr = fast_div(idx, a.d_mul, a.d_shift), c = idx - rr * (uint32_t)D;
      s_tile[rr * DS + c] = base ? ldg_stream(base + idx) : 0.f;
    }
  }
  __syncthreads();
  const int64_t o = o0 + r;
  const bool active = r < n_rows;
  float lad_acc = 0.f, gp_max = 0.f;
  if (active) {
    const int64_t img = o >> a.p_hw_shift, pix = o & (a.p_hw - 1);
    const int64_t pbase = (img * ct * P) * a.p_hw + pix;          // feature f of this row at pbase + f * HW
    const float* po = a.params + pbase;
    const uint32_t hw = (uint32_t)a.p_hw;
    const float gl_s = (BWD && a.g_lad_sample) ? a.g_lad_sample[o] : 0.f;
    float* cell = s_tile + r * DS + t_off;
    for (int ch = cg; ch < ct; ch += groups) {
      const uint32_t f0 = (uint32_t)(ch * P) * hw;
      float p[P];
#pragma unroll
      for (int j = 0; j < P; ++j) p[j] = ldg_stream(po + f0 + j * hw);
      if (!BWD) {
        float out, lad;
        int bin;
        rqs_eval<K, INV>(cell[ch * x_sc], p, a.c, out, lad, bin);
        cell[ch * x_sc] = out;
        if (a.lad_elem) a.lad_elem[o * ct + ch] = lad;
        if (a.bins) a.bins[o * ct + ch] = bin;
        lad_acc += lad;
      } else {
        const float v = a.x[o * D + t_off + ch * x_sc];
        const float vo = INV ? a.x_out[o * D + t_off + ch * x_sc] : 0.f;
        const float gl = gl_s + (a.g_lad_elem ? a.g_lad_elem[o * ct + ch] : 0.f);
        float g_in;
        float* gpe = a.g_params + pbase + f0;
        rqs_grad<K, INV>(v, vo, p, a.c, cell[ch * x_sc], gl, g_in, [&](int j, float val) { stg_stream(gpe + j * hw, val); gp_max = fmaxf(gp_max, fabsf(val)); });
        cell[ch * x_sc] = g_in;
      }
    }
  }
  if (!BWD && a.lad_sample != nullptr) s_lad[threadIdx.x] = lad_acc;
  if (BWD && a.g_params_amax != nullptr) publish_amax(a.g_params_amax, gp_max);
  __syncthreads();
  {
    float* base = a.y + o0 * D;
    for (uint32_t idx = threadIdx.x; idx < (uint32_t)(n_rows * D); idx += kTileThreads) {
      const uint32_t rr = fast_div(idx, a.d_mul, a.d_shift), c = idx - rr * (uint32_t)D;
      stg_stream(base + idx, s_tile[rr * DS + c]);
    }
  }
  if (!BWD && a.lad_sample != nullptr && cg == 0 && active) {
    float s = 0.f;
    for (int g2 = 0; g2 < groups; ++g2) s += s_lad[g2 * tr + r];   // fixed order: deterministic
    a.lad_sample[o] = (a.lad_base ? a.lad_base[o] : 0.f) + s;
  }
}

// ---------------------------------------------------------------------------------------------
// spline on a box [left, right] x [bottom, top] with all K+1 knot derivatives given
// (rational_quadratic_spline, rational_quadratic.py:67-168): function-level API, one thread per element,
// parameter rows [n, 3K+1] = [widths | heights | derivatives].
// ---------------------------------------------------------------------------------------------
struct RqsBoxArgs {
  const float* x;
  const float* x_out;
  const float* params;
  float* y;
  float* lad;
  int32_t* bins;
  const float* g_out;
  const float* g_lad;
  float* g_params;
  int64_t n;
  RqsConsts c;
};

template <int K, bool INV, bool BWD>
__global__ void __launch_bounds__(128) rqs_box_kernel(const RqsBoxArgs a) {
  constexpr int NP = 3 * K + 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  float p[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) p[j] = a.params[i * NP + j];
  const float v = a.x[i];
  if (!BWD) {
    float out, lad;
    int bin;
    rqs_eval<K, INV, NP>(v, p, a.c, out, lad, bin);
    a.y[i] = out;
    a.lad[i] = lad;
    if (a.bins) a.bins[i] = bin;
  } else {
    float g_in;
    float* gp = a.g_params + i * NP;
    rqs_grad<K, INV, NP>(v, INV ? a.x_out[i] : 0.f, p, a.c, a.g_out ? a.g_out[i] : 0.f, a.g_lad ? a.g_lad[i] : 0.f, g_in,
                         [&](int j, float val) { gp[j] = val; });
    a.y[i] = g_in;
  }
}

template <int K>
__global__ void rqs_search_kernel(const float* knots, const float* values, int32_t* bins, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float kn[K + 1];
#pragma unroll
  for (int j = 0; j <= K; ++j) kn[j] = knots[i * (K + 1) + j];
  float lo, hi;
  const float v = values[i];
  // same closed-interval rule as the spline kernels; values outside the knots get -1
  bins[i] = (v >= kn[0] && v <= kn[K]) ? rqs_search<K>(kn, v, lo, hi) : -1;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool fill_args(const mflow_rqs_desc* d, RqsArgs& a) {
  a.n_outer = d->n_outer;
  a.n_chan = d->n_chan;
  a.n_inner = d->n_inner;
  a.x_so = d->x_so; a.x_sc = d->x_sc; a.y_so = d->y_so; a.y_sc = d->y_sc;
  a.p_so = d->p_so; a.p_sc = d->p_sc; a.p_sp = d->p_sp; a.p_si = d->p_si;
  a.n_id_chan = d->n_id_chan; a.id_off = d->id_off; a.id_sc = d->id_sc;
  a.p_hw = d->p_hw; a.t_off = d->t_off; a.p_hw_shift = 0;
  while (d->p_hw > 0 && (1ll << a.p_hw_shift) < d->p_hw) ++a.p_hw_shift;
  const int K = d->num_bins;
  const double T = (double)d->tail_bound;
  RqsConsts& c = a.c;
  c.tail = (float)T;
  c.lo = (float)(-T);
  c.span = (float)(T - (-T));
  c.min_w = d->min_bin_width;
  c.min_h = d->min_bin_height;
  c.min_d = d->min_derivative;
  // Python evaluates (1 - min * K) and log(exp(1 - min_d) - 1) in double, torch rounds them once
  c.cfac_w = (float)(1.0 - (double)d->min_bin_width * K);
  c.cfac_h = (float)(1.0 - (double)d->min_bin_height * K);
  c.edge = (float)log(exp(1.0 - (double)d->min_derivative) - 1.0);
  c.divisor = d->wh_divisor;
  c.rcp_divisor = 1.0f / d->wh_divisor;
  c.use_divisor = d->wh_divisor != 1.0f;
  // magic number for e / dv, e < 2^31: power of two -> shift; else ceil(2^(32+s) / d) with s = ceil(log2 d) - 1...
  auto magic = [](uint64_t dv, uint32_t& mul, uint32_t& shift) {
    if ((dv & (dv - 1)) == 0) {
      mul = 0; shift = 0;
      while ((1ull << shift) < dv) ++shift;
    } else {
      uint32_t s = 0;
      while ((1ull << s) < dv) ++s;            // s = ceil(log2 d)
      mul = (uint32_t)(((1ull << (31 + s)) + dv - 1) / dv);  // ceil(2^(31+s) / d) < 2^32, exact for e < 2^31
      shift = s - 1;
    }
  };
  magic((uint64_t)d->n_inner, a.inner_mul, a.inner_shift);
  if (d->p_hw > 0 && d->x_so >= 1) {
    magic((uint64_t)d->x_so, a.d_mul, a.d_shift);
    a.tile_shift = tile_rows_shift(d->n_chan, d->x_so);
  }
  return true;
}

static int plane_chunks(const mflow_rqs_desc* d) {
  // CTAs per sample: one per 256 elements, capped so the grid stays near 16 CTAs per SM (each
  // thread then loops); never more than 1024 partial sums per sample.
  const int64_t per_sample = d->n_chan * d->n_inner;
  int64_t chunks = (per_sample + kPlaneThreads - 1) / kPlaneThreads;
  const int64_t cap = ((int64_t)kNumSMs * 16 + d->n_outer - 1) / (d->n_outer > 0 ? d->n_outer : 1);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  return (int)chunks;
}

// plane size for which a compile-time specialisation exists (K = 11, standard planes layout), else 0
static int planes_inner_const(const mflow_rqs_desc* d) {
  const int P = 3 * d->num_bins - 1;
  if (d->num_bins != 11 || d->p_si != 1 || d->p_sp != d->n_inner || d->p_sc != P * d->n_inner) return 0;
  const int64_t i = d->n_inner;
  return (i == 16 || i == 64 || i == 256 || i == 1024) ? (int)i : 0;
}

static bool rows_fast_path(const mflow_rqs_desc* d) {
  const int P = 3 * d->num_bins - 1;
  return d->n_inner == 1 && d->p_sp == 1 && d->p_sc == P && d->p_so == d->n_chan * P && d->n_chan <= kRowThreads &&
         d->n_id_chan <= 4 * kRowThreads;
}

// --splinebins is a free integer in the reference (experiments/train.py:60); every K from 2 to 16 plus 20, 24 and
// 32 is compiled in (the parameters of an element live in registers, so K is a template argument)
#define MFLOW_K_CASE(N_, ...) case N_: { constexpr int K = N_; __VA_ARGS__; break; }
#define MFLOW_DISPATCH_K(K_, ...)                    \
  switch (K_) {                                      \
    MFLOW_K_CASE(2, __VA_ARGS__) MFLOW_K_CASE(3, __VA_ARGS__) MFLOW_K_CASE(4, __VA_ARGS__) MFLOW_K_CASE(5, __VA_ARGS__)     \
    MFLOW_K_CASE(6, __VA_ARGS__) MFLOW_K_CASE(7, __VA_ARGS__) MFLOW_K_CASE(8, __VA_ARGS__) MFLOW_K_CASE(9, __VA_ARGS__)     \
    MFLOW_K_CASE(10, __VA_ARGS__) MFLOW_K_CASE(11, __VA_ARGS__) MFLOW_K_CASE(12, __VA_ARGS__) MFLOW_K_CASE(13, __VA_ARGS__) \
    MFLOW_K_CASE(14, __VA_ARGS__) MFLOW_K_CASE(15, __VA_ARGS__) MFLOW_K_CASE(16, __VA_ARGS__) MFLOW_K_CASE(20, __VA_ARGS__) \
    MFLOW_K_CASE(24, __VA_ARGS__) MFLOW_K_CASE(32, __VA_ARGS__)                                                             \
    default:                                         \
      set_error("num_bins=%d is not compiled in (supported: 2..16, 20, 24, 32)", (int)(K_)); \
      return MFLOW_E_BADARG;                         \
  }

static int validate(const mflow_rqs_desc* d) {
  MFLOW_REQUIRE(d != nullptr, "null descriptor");
  MFLOW_REQUIRE(d->n_outer >= 0 && d->n_chan >= 1 && d->n_inner >= 1, "bad extents");
  MFLOW_REQUIRE(d->n_outer <= 65535 * 1024LL, "n_outer too large");
  MFLOW_REQUIRE(d->tail_bound > 0.f, "tail_bound must be positive");
  MFLOW_REQUIRE(d->wh_divisor > 0.f, "wh_divisor must be positive");
  MFLOW_REQUIRE(d->n_chan * d->n_inner < (1ll << 30) && d->n_inner * d->p_sp < (1ll << 26) && d->n_chan * d->p_sc < (1ll << 31) &&
                    d->n_chan * d->x_sc < (1ll << 31) && d->n_chan * d->y_sc < (1ll << 31) && d->n_id_chan * d->id_sc < (1ll << 31),
                "per-sample extents must fit in 32 bits");
  MFLOW_REQUIRE((double)d->min_bin_width * d->num_bins <= 1.0, "Minimal bin width too large for the number of bins");
  MFLOW_REQUIRE((double)d->min_bin_height * d->num_bins <= 1.0, "Minimal bin height too large for the number of bins");
  MFLOW_REQUIRE(d->p_hw >= 0 && (d->p_hw & (d->p_hw - 1)) == 0, "p_hw must be 0 or a power of two");
  MFLOW_REQUIRE(d->p_hw == 0 || (d->n_inner == 1 && d->n_outer % d->p_hw == 0),
                "rows x planes layout: vectors (n_inner == 1) with n_outer a multiple of p_hw");
  MFLOW_REQUIRE(d->p_hw == 0 || (d->x_so == d->y_so && d->x_sc == d->y_sc && d->x_so >= 1 && d->x_so <= kTileMaxD && d->t_off >= 0 &&
                                 d->t_off + (d->n_chan - 1) * d->x_sc < d->x_so && d->n_chan * (3 * d->num_bins - 1) * d->p_hw < (1ll << 31)),
                "rows x planes layout: dense rows of at most 92 features, x and y alike");
  return MFLOW_OK;
}

}  // namespace mflow

using namespace mflow;

extern "C" MFLOW_API int64_t mflow_rqs_workspace_bytes(const mflow_rqs_desc* d) {
  if (d == nullptr || d->n_outer <= 0) return 16;
  const int chunks = plane_chunks(d);
  return kCounterBytes + (int64_t)d->n_outer * chunks * 4 + 16;
}

extern "C" MFLOW_API int mflow_rqs_apply(const mflow_rqs_desc* d, const float* x, const float* params, float* y, float* lad_elem,
                               float* lad_sample, const float* lad_base, int32_t* bins, void* workspace, void* stream) {
  if (int rc = validate(d)) return rc;
  if (d->n_outer == 0) return MFLOW_OK;
  MFLOW_REQUIRE(x && params && y, "null tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  RqsArgs a{};
  fill_args(d, a);
  a.x = x; a.params = params; a.y = y; a.lad_elem = lad_elem; a.lad_sample = lad_sample; a.lad_base = lad_base;
  a.bins = bins;
  if (d->p_hw > 0) {
    const int64_t tile_rows = 1ll << a.tile_shift;
    const unsigned grid = (unsigned)((d->n_outer + tile_rows - 1) / tile_rows);
    const size_t tile_smem = (size_t)tile_rows * ((size_t)d->x_so | 1) * sizeof(float);
    MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_rowtile_kernel<K, true, false><<<grid, kTileThreads, tile_smem, st>>>(a);
                     else rqs_rowtile_kernel<K, false, false><<<grid, kTileThreads, tile_smem, st>>>(a));
    return check_launch("rqs_rowtile_kernel");
  }
  if (rows_fast_path(d)) {
    const int spc = kRowThreads / (int)d->n_chan;
    const unsigned grid = (unsigned)((d->n_outer + spc - 1) / spc);
    MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_rows_kernel<K, true, false><<<grid, kRowThreads, 0, st>>>(a, spc);
                     else rqs_rows_kernel<K, false, false><<<grid, kRowThreads, 0, st>>>(a, spc));
    return check_launch("rqs_rows_kernel");
  }
  MFLOW_REQUIRE(d->n_outer <= 65535, "planes layout: n_outer must be <= 65535 (split the batch)");
  const int chunks = plane_chunks(d);
  a.chunks = chunks;
  if (lad_sample && chunks > 1) {
    MFLOW_REQUIRE(workspace != nullptr, "workspace required for the per-sample reduction");
    a.counters = (unsigned int*)workspace;  // fixed-size counter region first: stays zero between launches
    a.partial = (float*)((char*)workspace + kCounterBytes);
  }
  dim3 grid(chunks, (unsigned)d->n_outer);
  const int inner = planes_inner_const(d);
  if (inner) {
#define MFLOW_APPLY_INNER(I_)                                                                                      \
  case I_:                                                                                                         \
    if (d->inverse) rqs_planes_apply_kernel<11, true, I_><<<grid, kPlaneThreads, 0, st>>>(a);                       \
    else rqs_planes_apply_kernel<11, false, I_><<<grid, kPlaneThreads, 0, st>>>(a);                                \
    break;
    switch (inner) { MFLOW_APPLY_INNER(16) MFLOW_APPLY_INNER(64) MFLOW_APPLY_INNER(256) MFLOW_APPLY_INNER(1024) }
#undef MFLOW_APPLY_INNER
    return check_launch("rqs_planes_apply_kernel");
  }
  MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_planes_apply_kernel<K, true, 0><<<grid, kPlaneThreads, 0, st>>>(a);
                   else rqs_planes_apply_kernel<K, false, 0><<<grid, kPlaneThreads, 0, st>>>(a));
  return check_launch("rqs_planes_apply_kernel");
}

extern "C" MFLOW_API int mflow_rqs_backward(const mflow_rqs_desc* d, const float* x_in, const float* x_out, const float* params,
                                  const float* g_out, const float* g_lad_sample, const float* g_lad_elem, float* g_in,
                                  float* g_params, float* g_params_amax, void* stream) {
  if (int rc = validate(d)) return rc;
  if (d->n_outer == 0) return MFLOW_OK;
  MFLOW_REQUIRE(x_in && params && g_in && g_params, "null tensor pointer");
  MFLOW_REQUIRE(!d->inverse || x_out, "inverse backward needs the output of the apply call");
  cudaStream_t st = (cudaStream_t)stream;
  RqsArgs a{};
  fill_args(d, a);
  a.x = x_in; a.x_out = x_out; a.params = params; a.y = g_in; a.g_out = g_out; a.g_lad_sample = g_lad_sample;
  a.g_lad_elem = g_lad_elem; a.g_params = g_params; a.g_params_amax = g_params_amax;
  if (d->p_hw > 0) {
    const int64_t tile_rows = 1ll << a.tile_shift;
    const unsigned grid = (unsigned)((d->n_outer + tile_rows - 1) / tile_rows);
    const size_t tile_smem = (size_t)tile_rows * ((size_t)d->x_so | 1) * sizeof(float);
    MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_rowtile_kernel<K, true, true><<<grid, kTileThreads, tile_smem, st>>>(a);
                     else rqs_rowtile_kernel<K, false, true><<<grid, kTileThreads, tile_smem, st>>>(a));
    return check_launch("rqs_rowtile_kernel(bwd)");
  }
  if (rows_fast_path(d)) {
    const int spc = kRowThreads / (int)d->n_chan;
    const unsigned grid = (unsigned)((d->n_outer + spc - 1) / spc);
    MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_rows_kernel<K, true, true><<<grid, kRowThreads, 0, st>>>(a, spc);
                     else rqs_rows_kernel<K, false, true><<<grid, kRowThreads, 0, st>>>(a, spc));
    return check_launch("rqs_rows_kernel(bwd)");
  }
  MFLOW_REQUIRE(d->n_outer <= 65535, "planes layout: n_outer must be <= 65535 (split the batch)");
  dim3 grid(plane_chunks(d), (unsigned)d->n_outer);
  const int inner = planes_inner_const(d);
  if (inner) {
#define MFLOW_BWD_INNER(I_)                                                                                        \
  case I_:                                                                                                         \
    if (d->inverse) rqs_planes_backward_kernel<11, true, I_><<<grid, kPlaneThreads, 0, st>>>(a);                    \
    else rqs_planes_backward_kernel<11, false, I_><<<grid, kPlaneThreads, 0, st>>>(a);                             \
    break;
    switch (inner) { MFLOW_BWD_INNER(16) MFLOW_BWD_INNER(64) MFLOW_BWD_INNER(256) MFLOW_BWD_INNER(1024) }
#undef MFLOW_BWD_INNER
    return check_launch("rqs_planes_backward_kernel");
  }
  MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_planes_backward_kernel<K, true, 0><<<grid, kPlaneThreads, 0, st>>>(a);
                   else rqs_planes_backward_kernel<K, false, 0><<<grid, kPlaneThreads, 0, st>>>(a));
  return check_launch("rqs_planes_backward_kernel");
}

extern "C" MFLOW_API int mflow_rqs_search(const float* knots, const float* values, int32_t* bins, int64_t n, int32_t num_bins,
                                void* stream) {
  MFLOW_REQUIRE(knots && values && bins, "null tensor pointer");
  if (n == 0) return MFLOW_OK;
  const unsigned grid = (unsigned)((n + 127) / 128);
  MFLOW_DISPATCH_K(num_bins, rqs_search_kernel<K><<<grid, 128, 0, (cudaStream_t)stream>>>(knots, values, bins, n));
  return check_launch("rqs_search_kernel");
}

// rational_quadratic_spline on a box (rational_quadratic.py:67-168); see include/mflow_b200.h
static int fill_box(const mflow_rqs_box_desc* d, RqsBoxArgs& a) {
  MFLOW_REQUIRE(d != nullptr, "null descriptor");
  MFLOW_REQUIRE(d->n >= 0, "bad extents");
  MFLOW_REQUIRE(d->right > d->left && d->top > d->bottom, "empty box");
  MFLOW_REQUIRE((double)d->min_bin_width * d->num_bins <= 1.0, "Minimal bin width too large for the number of bins");
  MFLOW_REQUIRE((double)d->min_bin_height * d->num_bins <= 1.0, "Minimal bin height too large for the number of bins");
  a.n = d->n;
  RqsConsts& c = a.c;
  const int K = d->num_bins;
  c.lo = d->left; c.tail = d->right; c.span = (float)((double)d->right - (double)d->left);
  c.lo_y = d->bottom; c.hi_y = d->top; c.span_y = (float)((double)d->top - (double)d->bottom);
  c.min_w = d->min_bin_width; c.min_h = d->min_bin_height; c.min_d = d->min_derivative;
  c.cfac_w = (float)(1.0 - (double)d->min_bin_width * K);
  c.cfac_h = (float)(1.0 - (double)d->min_bin_height * K);
  c.edge = 0.f; c.divisor = 1.f; c.rcp_divisor = 1.f; c.use_divisor = 0;
  return MFLOW_OK;
}

extern "C" MFLOW_API int mflow_rqs_box_apply(const mflow_rqs_box_desc* d, const float* x, const float* params, float* y, float* lad,
                                             int32_t* bins, void* stream) {
  RqsBoxArgs a{};
  if (int rc = fill_box(d, a)) return rc;
  if (d->n == 0) return MFLOW_OK;
  MFLOW_REQUIRE(x && params && y && lad, "null tensor pointer");
  a.x = x; a.params = params; a.y = y; a.lad = lad; a.bins = bins;
  const unsigned grid = (unsigned)((d->n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_box_kernel<K, true, false><<<grid, 128, 0, st>>>(a);
                   else rqs_box_kernel<K, false, false><<<grid, 128, 0, st>>>(a));
  return check_launch("rqs_box_kernel");
}

extern "C" MFLOW_API int mflow_rqs_box_backward(const mflow_rqs_box_desc* d, const float* x_in, const float* x_out,
                                                const float* params, const float* g_out, const float* g_lad, float* g_in,
                                                float* g_params, void* stream) {
  RqsBoxArgs a{};
  if (int rc = fill_box(d, a)) return rc;
  if (d->n == 0) return MFLOW_OK;
  MFLOW_REQUIRE(x_in && params && g_in && g_params, "null tensor pointer");
  MFLOW_REQUIRE(!d->inverse || x_out, "inverse backward needs the output of the apply call");
  a.x = x_in; a.x_out = x_out; a.params = params; a.y = g_in; a.g_out = g_out; a.g_lad = g_lad; a.g_params = g_params;
  const unsigned grid = (unsigned)((d->n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  MFLOW_DISPATCH_K(d->num_bins, if (d->inverse) rqs_box_kernel<K, true, true><<<grid, 128, 0, st>>>(a);
                   else rqs_box_kernel<K, false, true><<<grid, 128, 0, st>>>(a));
  return check_launch("rqs_box_kernel(bwd)");
}
