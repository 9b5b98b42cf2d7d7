// Vector conditioner (ResidualNet) as ONE forward launch and TWO backward launches per coupling layer.
//
//   h  = W_0 [x | ctx] + b_0
//   h += gate_i * (W_ib relu(W_ia relu(h) + b_ia) + b_ib),   gate_i = sigmoid(W_ic ctx + b_ic) or 1      i = 0 .. blocks-1
//   out = W_f h + b_f
//
// (reference: manifold_flow/nn/resnet.py:9-72 - ResidualBlock / ResidualNet with ReLU, no batch norm, no dropout,
// `glu(cat(t, context_layer(ctx)))` = t * sigmoid(context_layer(ctx)), :38-39).  The reference issues six small GEMMs
// and ~10 pointwise launches per evaluation (~55 with the backward bookkeeping), each moving a [B, 100] activation through
// HBM; at the batch sizes the BASELINE configs run (100 .. a few thousand rows) that is pure launch latency.  Here a CTA
// owns R rows, keeps every activation of those rows in shared memory and walks the layers; thread j owns output
// feature j and streams its weight row (<= 100 floats per hidden layer, L1 / L2 resident) with 16-byte loads while the
// activations arrive as broadcast 16-byte shared loads.  fp32 FMA throughout: with M = R rows per CTA this
// is not tensor-core shaped (from 32768 rows on the conditioner runs on the tcgen05 convolution kernels instead,
// manifold_flow/b200/conditioner.py).
//
// Backward: one kernel recomputes the forward of its rows (nothing but x and ctx is saved), back-propagates to the
// input and leaves, per layer, the pair (A_l = layer input, G_l = gradient at the layer output) in a workspace;
// a second kernel reduces dW_l = G_l^T A_l and db_l = sum_rows G_l over all rows in a fixed order (deterministic).
#include <stdarg.h>

#include "mflow_common.cuh"

namespace mflow {

constexpr int kMlpFeat = 128;      // hidden features (and output features per pass of the final layer) per CTA
constexpr int kMlpSplit = 4;       // threads per feature: they split the reduction index and add up by warp shuffles
constexpr int kMlpThreads = kMlpFeat * kMlpSplit;
constexpr int kMlpMaxBlocks = 4;

struct MlpLayer {
  const float* w;   // [J][K] row-major (nn.Linear.weight)
  const float* b;   // [J]
  int J, K;
};

struct MlpArgs {
  const float* x; int64_t x_ld; int64_t x_step;      // input feature c of row r at x[r * x_ld + c * x_step]
  const float* ctx; int64_t ctx_ld;                  // [rows, n_ctx] or null
  float* out; int64_t out_ld;                        // [rows, n_out]
  int rows, n_in, n_ctx, hidden, n_out, n_blocks, gated;
  MlpLayer first, lin_a[kMlpMaxBlocks], lin_b[kMlpMaxBlocks], gate[kMlpMaxBlocks], last;
  // backward
  const float* g_out;                                // [rows, n_out]
  float* g_x;                                        // [rows, n_in] compact
  float* ws;                                         // per-layer (A_l, G_l) pairs, see mlp_ws_layout
};

__device__ __forceinline__ float quad_sum(float v) {   // sum over the kMlpSplit = 4 adjacent lanes of a feature
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// dense layer for the CTA's R rows: acc[r] = bias[j] + sum_k W[j][k] act[r][k] for output feature j; the four threads
// (j, q = 0..3) of the feature take every fourth 16-byte piece of the weight row - a warp reads 8 rows x 64 contiguous
// bytes per load - and the activations arrive as broadcast 16-byte shared loads (act: shared [R][ld], ld % 4 == 0).
// With one thread per feature a CTA was four warps, one per scheduler, and every L1 / L2 latency of the weight stream was
// exposed (100 us per launch at 256 rows); sixteen warps hide it.  Rows whose length is not a multiple of 4 (the first
// layer: in + context features; the context gates) take the scalar loop.  All lanes of a warp must call (shuffles); the
// result is valid in every lane of the quad.  The caller synchronises around writes to act.
template <int R>
__device__ __forceinline__ void dense_rows(const MlpLayer& L, int j, int q, const float* act, int ld, float (&acc)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  if (j < L.J) {
    const float* wr = L.w + (int64_t)j * L.K;
    if ((L.K & 3) == 0 && ((reinterpret_cast<uintptr_t>(wr) | reinterpret_cast<uintptr_t>(act)) & 15) == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(wr);
#pragma unroll 2
      for (int k4 = q; k4 < L.K / 4; k4 += kMlpSplit) {
        const float4 w = __ldg(w4 + k4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 a4 = *reinterpret_cast<const float4*>(act + r * ld + 4 * k4);
          acc[r] = fmaf(w.x, a4.x, fmaf(w.y, a4.y, fmaf(w.z, a4.z, fmaf(w.w, a4.w, acc[r]))));
        }
      }
    } else {
      for (int k = q; k < L.K; k += kMlpSplit) {
        const float w = __ldg(wr + k);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(w, act[r * ld + k], acc[r]);
      }
    }
  }
  const float bias = (j < L.J) ? __ldg(L.b + j) : 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = quad_sum(acc[r]) + bias;
}

// transposed product for the backward pass: acc[r] = sum_j W[j][k] g[r][j] for INPUT feature k; the quad's threads take
// every fourth group of four rows j (threads read W[j][k] for consecutive k: coalesced in the original layout);
// g: shared [R][ld], ld % 4 == 0
template <int R>
__device__ __forceinline__ void dense_rows_t(const MlpLayer& L, int k, int q, const float* g, int ld, float (&acc)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  if (k < L.K) {
    const float* wc = L.w + k;
    const int j_full = L.J & ~3;
#pragma unroll 2
    for (int j = 4 * q; j < j_full; j += 4 * kMlpSplit) {
      const float w0 = __ldg(wc + (int64_t)j * L.K), w1 = __ldg(wc + (int64_t)(j + 1) * L.K), w2 = __ldg(wc + (int64_t)(j + 2) * L.K),
                  w3 = __ldg(wc + (int64_t)(j + 3) * L.K);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 g4 = *reinterpret_cast<const float4*>(g + r * ld + j);
        acc[r] = fmaf(w0, g4.x, fmaf(w1, g4.y, fmaf(w2, g4.z, fmaf(w3, g4.w, acc[r]))));
      }
    }
    for (int j = j_full + q; j < L.J; j += kMlpSplit) {
      const float w = __ldg(wc + (int64_t)j * L.K);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(w, g[r * ld + j], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = quad_sum(acc[r]);
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

constexpr int kLd = kMlpFeat;      // row pitch of the shared activation arrays

// the forward of the CTA's rows up to (not including) the final layer; optionally records the intermediates the
// backward needs: s_keep holds, per block, {h_in, t (pre-ReLU output of lin_a), u (pre-gate output of lin_b), gate}
template <int R, bool KEEP>
__device__ __forceinline__ void mlp_trunk(const MlpArgs& a, int row0, int nrows, float* s_in, float* s_h, float* s_act, float* s_keep) {
  const int j = threadIdx.x / kMlpSplit, q = threadIdx.x % kMlpSplit;   // feature and the thread's share of its reductions
  const bool wr = q == 0;                                                // one lane of the quad writes
  const int n0 = a.n_in + a.n_ctx;
  // inputs [x | ctx] -> shared
  for (int idx = threadIdx.x; idx < R * kLd; idx += kMlpThreads) {
    const int r = idx / kLd, c = idx - r * kLd;
    float v = 0.f;
    if (r < nrows) {
      if (c < a.n_in) v = a.x[(int64_t)(row0 + r) * a.x_ld + (int64_t)c * a.x_step];
      else if (c < n0) v = a.ctx[(int64_t)(row0 + r) * a.ctx_ld + (c - a.n_in)];
    }
    s_in[idx] = v;
  }
  float acc[R];
  __syncthreads();
  dense_rows<R>(a.first, j, q, s_in, kLd, acc);
  if (wr) {
#pragma unroll
    for (int r = 0; r < R; ++r) s_h[r * kLd + j] = (j < a.hidden) ? acc[r] : 0.f;
  }
  for (int blk = 0; blk < a.n_blocks; ++blk) {
    float* keep = KEEP ? s_keep + blk * 4 * R * kLd : nullptr;
    __syncthreads();
    if (wr) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float h = s_h[r * kLd + j];
        if (KEEP) keep[r * kLd + j] = h;
        s_act[r * kLd + j] = fmaxf(h, 0.f);
      }
    }
    __syncthreads();
    dense_rows<R>(a.lin_a[blk], j, q, s_act, kLd, acc);
    __syncthreads();
    if (wr) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (KEEP) keep[(R + r) * kLd + j] = acc[r];
        s_act[r * kLd + j] = fmaxf(acc[r], 0.f);
      }
    }
    float u[R];
    __syncthreads();
    dense_rows<R>(a.lin_b[blk], j, q, s_act, kLd, u);
    float gt[R];
#pragma unroll
    for (int r = 0; r < R; ++r) gt[r] = 1.f;
    if (a.gated) {
      float gp[R];
      dense_rows<R>(a.gate[blk], j, q, s_in + a.n_in, kLd, gp);   // the context sits behind x in s_in
#pragma unroll
      for (int r = 0; r < R; ++r) gt[r] = sigmoidf_(gp[r]);
    }
    if (wr) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (KEEP) { keep[(2 * R + r) * kLd + j] = u[r]; keep[(3 * R + r) * kLd + j] = gt[r]; }
        if (j < a.hidden) s_h[r * kLd + j] += gt[r] * u[r];
      }
    }
  }
  __syncthreads();
}

template <int R>
__global__ void __launch_bounds__(kMlpThreads) mlp_forward_kernel(const MlpArgs a) {
  extern __shared__ float smem_f[];
  float* s_in = smem_f;
  float* s_h = s_in + R * kLd;
  float* s_act = s_h + R * kLd;
  const int row0 = blockIdx.x * R, nrows = min(R, a.rows - row0);
  mlp_trunk<R, false>(a, row0, nrows, s_in, s_h, s_act, nullptr);
  // final layer: output features j, j + 128, ... for this thread
  const int j = threadIdx.x / kMlpSplit, q = threadIdx.x % kMlpSplit;
  for (int j0 = 0; j0 < a.n_out; j0 += kMlpFeat) {
    MlpLayer L = a.last;
    L.w += (int64_t)j0 * L.K;
    L.b += j0;
    L.J = min(kMlpFeat, a.n_out - j0);
    float acc[R];
    dense_rows<R>(L, j, q, s_h, kLd, acc);
    if (q == 0 && j < L.J) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < nrows) a.out[(int64_t)(row0 + r) * a.out_ld + j0 + j] = acc[r];
    }
  }
}

// Workspace of the backward pass (floats), per layer l in the order first, (a_0, b_0, gate_0), ..., last:
//   A_l [rows][K_l] then G_l [rows][J_l]   (G_last is g_out itself and is not copied)
struct MlpWsLayout {
  int64_t a_off[3 * kMlpMaxBlocks + 2], g_off[3 * kMlpMaxBlocks + 2];
  int n_layers;
  int64_t total;
};

static MlpWsLayout mlp_ws_layout(int rows, int n_in, int n_ctx, int hidden, int n_out, int n_blocks, int gated) {
  MlpWsLayout w{};
  int64_t off = 0;
  auto add = [&](int K, int J, bool with_g) {
    w.a_off[w.n_layers] = off; off += (int64_t)rows * K;
    w.g_off[w.n_layers] = with_g ? off : -1; if (with_g) off += (int64_t)rows * J;
    ++w.n_layers;
  };
  add(n_in + n_ctx, hidden, true);
  for (int b = 0; b < n_blocks; ++b) {
    add(hidden, hidden, true);
    add(hidden, hidden, true);
    if (gated) add(n_ctx, hidden, true);
  }
  add(hidden, n_out, false);
  w.total = off;
  return w;
}

struct MlpBwdOffsets { int64_t a_off[3 * kMlpMaxBlocks + 2], g_off[3 * kMlpMaxBlocks + 2]; };

template <int R>
__global__ void __launch_bounds__(kMlpThreads) mlp_backward_kernel(const MlpArgs a, const MlpBwdOffsets o) {
  extern __shared__ float smem_f[];
  float* s_in = smem_f;
  float* s_h = s_in + R * kLd;
  float* s_act = s_h + R * kLd;
  float* s_keep = s_act + R * kLd;                    // n_blocks x {h_in, t, u, gate} x [R][kLd]
  float* s_g = s_keep + a.n_blocks * 4 * R * kLd;     // gradient vectors [R][max(n_out, kLd)]
  const int j = threadIdx.x / kMlpSplit, q = threadIdx.x % kMlpSplit;
  const bool wr = q == 0;
  const int row0 = blockIdx.x * R, nrows = min(R, a.rows - row0);
  const int n0 = a.n_in + a.n_ctx, H = a.hidden;
  const int g_ld = (max(a.n_out, kLd) + 3) & ~3;
  mlp_trunk<R, true>(a, row0, nrows, s_in, s_h, s_act, s_keep);
  auto store = [&](int64_t off, int ld, int col, int r, float v) {   // workspace tensor [rows][ld]
    if (r < nrows && col < ld) a.ws[off + (int64_t)(row0 + r) * ld + col] = v;
  };
  int layer = 1 + a.n_blocks * (a.gated ? 3 : 2);     // index of the last layer
  // last layer: A = h_final; g_h = W_f^T g_out
  for (int idx = threadIdx.x; idx < R * a.n_out; idx += kMlpThreads) {
    const int r = idx / a.n_out, c = idx - r * a.n_out;
    s_g[r * g_ld + c] = r < nrows ? a.g_out[(int64_t)(row0 + r) * a.n_out + c] : 0.f;
  }
  if (wr) {
#pragma unroll
    for (int r = 0; r < R; ++r) store(o.a_off[layer], H, j, r, s_h[r * kLd + j]);
  }
  __syncthreads();
  float gh[R];   // gradient at the trunk state h (this quad's feature; valid in every lane of the quad)
  dense_rows_t<R>(a.last, j, q, s_g, g_ld, gh);
  for (int blk = a.n_blocks - 1; blk >= 0; --blk) {
    const float* keep = s_keep + blk * 4 * R * kLd;
    const int l_a = 1 + blk * (a.gated ? 3 : 2), l_b = l_a + 1, l_g = l_a + 2;
    __syncthreads();   // s_g / s_act free
    // h_out = h_in + gate * u:  g_u = gate * g_h ;  g_gatepre = g_h * u * gate * (1 - gate)
    if (wr) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float u = keep[(2 * R + r) * kLd + j], gt = keep[(3 * R + r) * kLd + j], t = keep[(R + r) * kLd + j];
        const float g_u = (j < H) ? gh[r] * gt : 0.f;
        s_g[r * g_ld + j] = g_u;                                         // G_b
        store(o.g_off[l_b], H, j, r, g_u);
        store(o.a_off[l_b], H, j, r, fmaxf(t, 0.f));                     // A_b = relu(t)
        if (a.gated) {
          store(o.g_off[l_g], H, j, r, (j < H) ? gh[r] * u * gt * (1.f - gt) : 0.f);
          if (j < a.n_ctx) store(o.a_off[l_g], a.n_ctx, j, r, s_in[r * kLd + a.n_in + j]);
        }
      }
    }
    __syncthreads();
    float ga[R];
    dense_rows_t<R>(a.lin_b[blk], j, q, s_g, g_ld, ga);                  // gradient at relu(t)
    __syncthreads();
    if (wr) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float t = keep[(R + r) * kLd + j], hin = keep[r * kLd + j];
        const float g_t = (j < H && t > 0.f) ? ga[r] : 0.f;
        s_g[r * g_ld + j] = g_t;                                         // G_a
        store(o.g_off[l_a], H, j, r, g_t);
        store(o.a_off[l_a], H, j, r, fmaxf(hin, 0.f));                   // A_a = relu(h_in)
      }
    }
    __syncthreads();
    dense_rows_t<R>(a.lin_a[blk], j, q, s_g, g_ld, ga);                  // gradient at relu(h_in)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float hin = keep[r * kLd + j];
      gh[r] += (j < H && hin > 0.f) ? ga[r] : 0.f;
    }
  }
  __syncthreads();
  // first layer: G = g_h0, A = [x | ctx]; g_x = (W_0^T g_h0)[:n_in]
  if (wr) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s_g[r * g_ld + j] = (j < H) ? gh[r] : 0.f;
      store(o.g_off[0], H, j, r, (j < H) ? gh[r] : 0.f);
      if (j < n0) store(o.a_off[0], n0, j, r, s_in[r * kLd + j]);
    }
  }
  __syncthreads();
  float gx[R];
  dense_rows_t<R>(a.first, j, q, s_g, g_ld, gx);
  if (wr && a.g_x != nullptr && j < a.n_in) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r < nrows) a.g_x[(int64_t)(row0 + r) * a.n_in + j] = gx[r];
  }
}

// dW_l[j][k] = sum_rows G_l[r][j] A_l[r][k], db_l[j] = sum_rows G_l[r][j]: one CTA per 32 x 32 tile of one layer, rows in
// order (deterministic).  blockIdx.x enumerates (layer, j tile, k tile) through the prefix table.
struct MlpGradLayer { const float* A; const float* G; float* dW; float* db; int J, K, tiles_k, tile0; };
struct MlpGradArgs { MlpGradLayer l[3 * kMlpMaxBlocks + 2]; int n_layers, rows; };

__global__ void __launch_bounds__(256) mlp_param_grad_kernel(const MlpGradArgs a) {
  __shared__ float s_g[32][33], s_a[32][33];
  int li = 0;
  while (li + 1 < a.n_layers && (int)blockIdx.x >= a.l[li + 1].tile0) ++li;
  const MlpGradLayer L = a.l[li];
  const int t = blockIdx.x - L.tile0, tj = t / L.tiles_k, tk = t - tj * L.tiles_k;
  const int j0 = tj * 32, k0 = tk * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty 0..7: output rows ty, ty+8, ty+16, ty+24 of the tile
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, bacc = 0.f;
  for (int r0 = 0; r0 < a.rows; r0 += 32) {
    // stage G[r0 .. r0+32)[j0 .. j0+32) and A[..][k0 .. k0+32): thread (ty, tx) loads rows ty, ty+8, ...
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + ty + 8 * q;
      s_g[ty + 8 * q][tx] = (r < a.rows && j0 + tx < L.J) ? L.G[(int64_t)r * L.J + j0 + tx] : 0.f;
      s_a[ty + 8 * q][tx] = (r < a.rows && k0 + tx < L.K) ? L.A[(int64_t)r * L.K + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float av = s_a[r][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(s_g[r][ty + 8 * q], av, acc[q]);
      if (tk == 0 && ty == 0) bacc += s_g[r][tx];
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = j0 + ty + 8 * q, k = k0 + tx;
    if (j < L.J && k < L.K) L.dW[(int64_t)j * L.K + k] = acc[q];
  }
  if (tk == 0 && ty == 0 && j0 + tx < L.J && L.db != nullptr) L.db[j0 + tx] = bacc;
}

}  // namespace mflow

using namespace mflow;

static int fill_mlp(const mflow_mlp_desc* d, const float* const* w, const float* const* b, MlpArgs& a) {
  MFLOW_REQUIRE(d && w && b, "null pointer");
  MFLOW_REQUIRE(d->rows >= 0 && d->n_in >= 1 && d->n_out >= 1, "bad extents");
  MFLOW_REQUIRE(d->hidden >= 1 && d->hidden <= kMlpFeat, "hidden features must be <= 128");
  MFLOW_REQUIRE(d->n_in + d->n_ctx <= kMlpFeat, "in_features + context_features must be <= 128");
  MFLOW_REQUIRE(d->n_blocks >= 0 && d->n_blocks <= kMlpMaxBlocks, "at most 4 residual blocks");
  a.rows = (int)d->rows; a.n_in = d->n_in; a.n_ctx = d->n_ctx; a.hidden = d->hidden; a.n_out = d->n_out;
  a.n_blocks = d->n_blocks; a.gated = d->n_ctx > 0 ? 1 : 0;
  int li = 0;
  auto layer = [&](int J, int K) { MlpLayer L{w[li], b[li], J, K}; ++li; return L; };
  a.first = layer(d->hidden, d->n_in + d->n_ctx);
  for (int i = 0; i < d->n_blocks; ++i) {
    a.lin_a[i] = layer(d->hidden, d->hidden);
    a.lin_b[i] = layer(d->hidden, d->hidden);
    if (a.gated) a.gate[i] = layer(d->hidden, d->n_ctx);
  }
  a.last = layer(d->n_out, d->hidden);
  for (int i = 0; i < li; ++i) MFLOW_REQUIRE(w[i] && b[i], "null weight / bias pointer");
  return MFLOW_OK;
}

// rows per CTA: few rows at small batches so that the grid covers the machine, more at large ones (weight reuse)
static int mlp_rows_per_cta(int64_t rows) { return rows <= 512 ? 2 : (rows <= 8192 ? 4 : 8); }

extern "C" MFLOW_API int64_t mflow_mlp_workspace_floats(const mflow_mlp_desc* d) {
  if (!d) return 0;
  return mlp_ws_layout((int)d->rows, d->n_in, d->n_ctx, d->hidden, d->n_out, d->n_blocks, d->n_ctx > 0).total;
}

extern "C" MFLOW_API int mflow_mlp_forward(const mflow_mlp_desc* d, const float* x, int64_t x_ld, int64_t x_step, const float* ctx,
                                           int64_t ctx_ld, const float* const* weights, const float* const* biases, float* out,
                                           void* stream) {
  MlpArgs a{};
  if (int rc = fill_mlp(d, weights, biases, a)) return rc;
  if (d->rows == 0) return MFLOW_OK;
  MFLOW_REQUIRE(x && out && (d->n_ctx == 0 || ctx), "null tensor pointer");
  a.x = x; a.x_ld = x_ld; a.x_step = x_step; a.ctx = ctx; a.ctx_ld = ctx_ld; a.out = out; a.out_ld = d->n_out;
  const int R = mlp_rows_per_cta(d->rows);
  const unsigned grid = (unsigned)((d->rows + R - 1) / R);
  const size_t smem = (size_t)(3 * R * kLd) * sizeof(float);
  if (R == 2) mlp_forward_kernel<2><<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(a);
  else if (R == 4) mlp_forward_kernel<4><<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(a);
  else mlp_forward_kernel<8><<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(a);
  return check_launch("mlp_forward_kernel");
}

extern "C" MFLOW_API int mflow_mlp_backward(const mflow_mlp_desc* d, const float* x, int64_t x_ld, int64_t x_step, const float* ctx,
                                            int64_t ctx_ld, const float* const* weights, const float* const* biases,
                                            const float* g_out, float* g_x, float* const* g_weights, float* const* g_biases,
                                            float* workspace, void* stream) {
  MlpArgs a{};
  if (int rc = fill_mlp(d, weights, biases, a)) return rc;
  MFLOW_REQUIRE(g_out && g_weights && g_biases && workspace, "null pointer");
  if (d->rows == 0) return MFLOW_OK;
  a.x = x; a.x_ld = x_ld; a.x_step = x_step; a.ctx = ctx; a.ctx_ld = ctx_ld;
  a.g_out = g_out; a.g_x = g_x; a.ws = workspace;
  const MlpWsLayout lay = mlp_ws_layout(a.rows, a.n_in, a.n_ctx, a.hidden, a.n_out, a.n_blocks, a.gated);
  MlpBwdOffsets o{};
  for (int i = 0; i < lay.n_layers; ++i) { o.a_off[i] = lay.a_off[i]; o.g_off[i] = lay.g_off[i]; }
  const int R = a.rows <= 512 ? 2 : 4;   // the recompute keeps 4 x blocks activation sets per row in shared memory
  const unsigned grid = (unsigned)((a.rows + R - 1) / R);
  const int g_ld = ((a.n_out > kLd ? a.n_out : kLd) + 3) & ~3;
  const size_t smem = (size_t)(3 * R * kLd + a.n_blocks * 4 * R * kLd + R * g_ld) * sizeof(float);
  MFLOW_REQUIRE(smem <= 200 * 1024, "out_features too large for the fused backward");
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(mlp_backward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(mlp_backward_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set[dev] = true;
  }
  if (R == 2) mlp_backward_kernel<2><<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(a, o);
  else mlp_backward_kernel<4><<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(a, o);
  if (int rc = check_launch("mlp_backward_kernel")) return rc;
  // parameter gradients of every layer in one launch
  MlpGradArgs g{};
  g.rows = a.rows;
  g.n_layers = lay.n_layers;
  int tile0 = 0, li = 0;
  auto add = [&](const MlpLayer& L, const float* G) {
    MlpGradLayer& q = g.l[li];
    q.A = workspace + lay.a_off[li];
    q.G = G;
    q.dW = g_weights[li];
    q.db = g_biases[li];
    q.J = L.J; q.K = L.K;
    q.tiles_k = (L.K + 31) / 32;
    q.tile0 = tile0;
    tile0 += ((L.J + 31) / 32) * q.tiles_k;
    ++li;
  };
  add(a.first, workspace + lay.g_off[0]);
  for (int i = 0; i < a.n_blocks; ++i) {
    add(a.lin_a[i], workspace + lay.g_off[li]);
    add(a.lin_b[i], workspace + lay.g_off[li]);
    if (a.gated) add(a.gate[i], workspace + lay.g_off[li]);
  }
  add(a.last, g_out);
  for (int i = 0; i < li; ++i) MFLOW_REQUIRE(g.l[i].dW != nullptr, "null weight-gradient pointer");
  mlp_param_grad_kernel<<<tile0, 256, 0, (cudaStream_t)stream>>>(g);
  return check_launch("mlp_param_grad_kernel");
}
