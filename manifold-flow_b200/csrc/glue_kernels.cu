// Linear-glue and likelihood-epilogue kernels (kernels (c) and (d) of BASELINE.json:north_star).
// All of them are HBM-bound copies / reductions; the point is to do each in ONE coalesced pass
// where the reference issues several ATen launches with intermediate tensors.
#include <math.h>
#include <stdarg.h>

#include "mflow_common.cuh"

namespace mflow {

// ---------------------------------------------------------------------------------------------
// y[o, :, i] = M x[o, :, i] + m          (ActNorm + channel permutation + LU 1x1 conv, fused)
// reference: transforms/normalization.py:99-131, transforms/conv.py:16-55, transforms/lu.py:56-95
// A CTA owns 32 consecutive "pixels" q = o*I + i; 4 warps split the output channels; M sits in
// shared memory and is read as a broadcast, the x tile is read with lane == pixel.
// ---------------------------------------------------------------------------------------------
constexpr int kMixPix = 32;
constexpr int kMixThreads = 128;
constexpr int kMixFwdThreads = 256;   // chanmix_kernel: 8 warps x 4 output channels = one 32-channel pass per tile
// pitch of the transposed matrix in shared memory: a multiple of 4 (16-byte row reads) that is 4 mod 8, so that the
// transposing stores of 32 consecutive k hit 8 different banks instead of one (C = 64, 96)
__host__ __device__ constexpr int mix_cp(int C) { return ((C + 3) & ~3) | 4; }

__global__ void __launch_bounds__(kMixFwdThreads) chanmix_kernel(const float* __restrict__ x, const float* __restrict__ mat,
                                                              const float* __restrict__ bias, float* __restrict__ y,
                                                              int64_t n_pix, int C, int64_t I, int64_t x_so, int64_t x_sc,
                                                              int64_t x_si, int64_t y_so, int64_t y_sc, int64_t y_si) {
  // The CTA keeps the matrix (transposed, [k][c] with c padded to a multiple of 4) in shared memory and walks
  // pixel tiles grid-stride, so the matrix is fetched once per CTA, not once per 32 pixels.  A thread produces
  // four output channels of one pixel at a time: per k one activation load and one 16-byte matrix load
  // (a broadcast) feed four FMAs.
  extern __shared__ __align__(16) float smem[];
  const int Cp = mix_cp(C);
  float* s_matT = smem;                 // [C][Cp]
  float* s_x = smem + C * Cp;           // [C][33]
  float* s_b = s_x + C * (kMixPix + 1);  // [Cp]
  // coalesced read of mat[c][k] along k, transposing store (the strided read - 32 lines per warp load, 72 rounds for a
  // 96 x 96 matrix - was most of the 22 us a launch on a 1.5 MB activation took)
  for (int idx = threadIdx.x; idx < C * C; idx += kMixFwdThreads) {
    const int c = idx / C, k = idx - c * C;
    s_matT[k * Cp + c] = mat[idx];
  }
  for (int idx = threadIdx.x; idx < C * (Cp - C); idx += kMixFwdThreads) {   // zero the padding columns
    const int k = idx / (Cp - C), c = C + idx - k * (Cp - C);
    s_matT[k * Cp + c] = 0.f;
  }
  for (int idx = threadIdx.x; idx < Cp; idx += kMixFwdThreads) s_b[idx] = (bias && idx < C) ? bias[idx] : 0.f;
  const bool chan_fast = (x_sc == 1);  // rows layout: channels contiguous -> walk channels fastest
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (n_pix + kMixPix - 1) / kMixPix;
  __shared__ float s_y[kMixPix][33];
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t q0 = tile * kMixPix;
    const int n_here = (int)min((int64_t)kMixPix, n_pix - q0);
    __syncthreads();   // the previous tile's readers are done with s_x (and the matrix is in place)
    for (int idx = threadIdx.x; idx < C * kMixPix; idx += kMixFwdThreads) {
      int k, pq;
      if (chan_fast) { pq = idx / C; k = idx - pq * C; } else { k = idx / kMixPix; pq = idx - k * kMixPix; }
      float v = 0.f;
      if (pq < n_here) {
        const int64_t q = q0 + pq, o = q / I, i = q - o * I;
        v = x[o * x_so + k * x_sc + i * x_si];
      }
      s_x[k * (kMixPix + 1) + pq] = v;
    }
    __syncthreads();
    if (y_sc == 1) {
      // rows layout: transpose through shared memory so that stores are contiguous
      for (int c0 = 0; c0 < C; c0 += 32) {
        for (int cc = warp * 4; cc < 32 && c0 + cc < C; cc += kMixFwdThreads / 8) {
          const int c = c0 + cc;
          float a0 = s_b[c], a1 = s_b[c + 1], a2 = s_b[c + 2], a3 = s_b[c + 3];
          for (int k = 0; k < C; ++k) {
            const float xv = s_x[k * (kMixPix + 1) + lane];
            const float4 m = *reinterpret_cast<const float4*>(s_matT + k * Cp + c);
            a0 = fmaf(m.x, xv, a0); a1 = fmaf(m.y, xv, a1); a2 = fmaf(m.z, xv, a2); a3 = fmaf(m.w, xv, a3);
          }
          s_y[lane][cc] = a0; s_y[lane][cc + 1] = a1; s_y[lane][cc + 2] = a2; s_y[lane][cc + 3] = a3;
        }
        __syncthreads();
        const int nc = min(32, C - c0);
        for (int idx = threadIdx.x; idx < kMixPix * nc; idx += kMixFwdThreads) {
          const int pq = idx / nc, cc = idx - pq * nc;
          if (pq < n_here) {
            const int64_t q = q0 + pq, o = q / I, i = q - o * I;
            y[o * y_so + (c0 + cc) * y_sc + i * y_si] = s_y[pq][cc];
          }
        }
        __syncthreads();
      }
    } else if (lane < n_here) {
      const int64_t q = q0 + lane, o = q / I, i = q - o * I;
      float* yq = y + o * y_so + i * y_si;
      for (int c = warp * 4; c < C; c += kMixFwdThreads / 8) {
        float a0 = s_b[c], a1 = s_b[c + 1], a2 = s_b[c + 2], a3 = s_b[c + 3];
        for (int k = 0; k < C; ++k) {
          const float xv = s_x[k * (kMixPix + 1) + lane];
          const float4 m = *reinterpret_cast<const float4*>(s_matT + k * Cp + c);
          a0 = fmaf(m.x, xv, a0); a1 = fmaf(m.y, xv, a1); a2 = fmaf(m.z, xv, a2); a3 = fmaf(m.w, xv, a3);
        }
        yq[c * y_sc] = a0;
        if (c + 1 < C) yq[(c + 1) * y_sc] = a1;
        if (c + 2 < C) yq[(c + 2) * y_sc] = a2;
        if (c + 3 < C) yq[(c + 3) * y_sc] = a3;
      }
    }
  }
}

// Rows layout at large batch (vector flows: [rows, C] with C <= 64 dense rows, 10^5 .. 10^6 rows): the generic kernel
// above walks 32-row tiles with three barriers and a 64-bit division per element (28 us for a 10 MB activation, 80 us
// for [2^20, 3]).  Here a CTA stages 128 whole rows through shared memory with flat, fully coalesced copies; thread
// (row, half) produces every other group of four output features of its row from the transposed matrix (broadcast
// 16-byte loads) and the results leave as whole rows again.
constexpr int kMixRowTile = 128;
__global__ void __launch_bounds__(256) chanmix_rows_kernel(const float* __restrict__ x, const float* __restrict__ mat,
                                                           const float* __restrict__ bias, float* __restrict__ y, int64_t rows, int C) {
  extern __shared__ __align__(16) float smem[];
  const int Cp = mix_cp(C), DS = C | 1;
  float* s_matT = smem;                       // [C][Cp]
  float* s_b = s_matT + C * Cp;               // [Cp]
  float* s_in = s_b + Cp;                     // [128][DS]
  float* s_out = s_in + kMixRowTile * DS;     // [128][DS]
  for (int idx = threadIdx.x; idx < C * C; idx += 256) {
    const int c = idx / C, k = idx - c * C;
    s_matT[k * Cp + c] = mat[idx];
  }
  for (int idx = threadIdx.x; idx < C * (Cp - C); idx += 256) {
    const int k = idx / (Cp - C), c = C + idx - k * (Cp - C);
    s_matT[k * Cp + c] = 0.f;
  }
  for (int idx = threadIdx.x; idx < Cp; idx += 256) s_b[idx] = (bias && idx < C) ? bias[idx] : 0.f;
  const int r = threadIdx.x & (kMixRowTile - 1), half = threadIdx.x >> 7;
  const int64_t n_tiles = (rows + kMixRowTile - 1) / kMixRowTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kMixRowTile;
    const int n_here = (int)min((int64_t)kMixRowTile, rows - r0);
    const float* xb = x + r0 * C;
    __syncthreads();   // the previous tile's stores have read s_out; the matrix is in place
    for (int idx = threadIdx.x; idx < n_here * C; idx += 256) {
      const int rr = idx / C, k = idx - rr * C;
      s_in[rr * DS + k] = ldg_stream(xb + idx);
    }
    __syncthreads();
    if (r < n_here) {
      const float* xr = s_in + r * DS;
      for (int c = half * 4; c < C; c += 8) {
        float a0 = s_b[c], a1 = s_b[c + 1], a2 = s_b[c + 2], a3 = s_b[c + 3];
        for (int k = 0; k < C; ++k) {
          const float xv = xr[k];
          const float4 m = *reinterpret_cast<const float4*>(s_matT + k * Cp + c);
          a0 = fmaf(m.x, xv, a0); a1 = fmaf(m.y, xv, a1); a2 = fmaf(m.z, xv, a2); a3 = fmaf(m.w, xv, a3);
        }
        float* o = s_out + r * DS + c;
        o[0] = a0;
        if (c + 1 < C) o[1] = a1;
        if (c + 2 < C) o[2] = a2;
        if (c + 3 < C) o[3] = a3;
      }
    }
    __syncthreads();
    float* yb = y + r0 * C;
    for (int idx = threadIdx.x; idx < n_here * C; idx += 256) {
      const int rr = idx / C, k = idx - rr * C;
      stg_stream(yb + idx, s_out[rr * DS + k]);
    }
  }
}

// Planes layout (image levels: [B, 12, 32, 32] ... [B, 96, 4, 4]): a thread owns PX consecutive pixels of one sample,
// keeps all C input channels of them in registers (C coalesced loads in flight per thread) and produces C / CG of
// the outputs from the matrix in shared memory (broadcast 16-byte loads); the CG threads of a pixel group re-read the
// same inputs through L1.  One read and one write of the activation, nothing staged: the HBM-bound form.
template <int PX> struct MixVec;
template <> struct MixVec<4> { using T = float4; };
template <> struct MixVec<2> { using T = float2; };
template <> struct MixVec<1> { using T = float; };
template <int C, int PX, int CG>
__global__ void __launch_bounds__(256) chanmix_planes_kernel(const float* __restrict__ x, const float* __restrict__ mat,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             int64_t n_outer, int64_t I, int64_t x_so, int64_t y_so) {
  using V = typename MixVec<PX>::T;
  __shared__ __align__(16) float s_mat[C * C];
  __shared__ float s_b[C];
  for (int idx = threadIdx.x; idx < C * C; idx += 256) s_mat[idx] = mat[idx];
  for (int idx = threadIdx.x; idx < C; idx += 256) s_b[idx] = bias ? bias[idx] : 0.f;
  __syncthreads();
  const int64_t ivn = I / PX, total = n_outer * ivn * CG;
  for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < total; q += (int64_t)gridDim.x * 256) {
    const int cg = (int)(q % CG);
    const int64_t qq = q / CG, o = qq / ivn, iv = qq - o * ivn;
    const V* xp = reinterpret_cast<const V*>(x + o * x_so) + iv;
    V* yp = reinterpret_cast<V*>(y + o * y_so) + iv;
    float v[C][PX];
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const V t = __ldg(xp + k * ivn);
      const float* tf = reinterpret_cast<const float*>(&t);
#pragma unroll
      for (int p = 0; p < PX; ++p) v[k][p] = tf[p];
    }
#pragma unroll 2
    for (int cc = 0; cc < C / CG; ++cc) {
      const int c = cg * (C / CG) + cc;
      float acc[PX];
#pragma unroll
      for (int p = 0; p < PX; ++p) acc[p] = s_b[c];
#pragma unroll
      for (int k4 = 0; k4 < C / 4; ++k4) {
        const float4 m = *reinterpret_cast<const float4*>(s_mat + c * C + 4 * k4);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          acc[p] = fmaf(m.x, v[4 * k4][p], acc[p]);
          acc[p] = fmaf(m.y, v[4 * k4 + 1][p], acc[p]);
          acc[p] = fmaf(m.z, v[4 * k4 + 2][p], acc[p]);
          acc[p] = fmaf(m.w, v[4 * k4 + 3][p], acc[p]);
        }
      }
      V out;
      float* of = reinterpret_cast<float*>(&out);
#pragma unroll
      for (int p = 0; p < PX; ++p) of[p] = acc[p];
      yp[c * ivn] = out;
    }
  }
}

// g_mat[c,k] = sum_q g_y[c,q] x[k,q],  g_bias[c] = sum_q g_y[c,q]: each CTA accumulates a slab of
// pixels into shared memory and writes its partial [C*C + C] block; a second kernel adds the partials in a
// fixed order (deterministic).  (One CTA summing all partials at the end capped the grid at a few dozen CTAs
// for 96 channels: 300 us for a 75 MFLOP problem.)
__global__ void __launch_bounds__(kMixThreads) chanmix_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                    float* __restrict__ g_mat, float* __restrict__ g_bias,
                                                                    float* __restrict__ partial,
                                                                    int64_t n_pix, int C, int64_t I, int64_t x_so, int64_t x_sc,
                                                                    int64_t x_si, int64_t y_so, int64_t y_sc, int64_t y_si) {
  extern __shared__ float smem[];
  const int pitch = kMixPix + 1;
  float* s_acc = smem;                  // [C*C + C]
  float* s_x = s_acc + C * C + C;       // [C][33]
  float* s_g = s_x + C * pitch;         // [C][33]
  const int n_out = C * C + C;
  for (int idx = threadIdx.x; idx < n_out; idx += kMixThreads) s_acc[idx] = 0.f;
  const bool chan_fast = (x_sc == 1);
  const int64_t n_tiles = (n_pix + kMixPix - 1) / kMixPix;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t q0 = tile * kMixPix;
    const int n_here = (int)min((int64_t)kMixPix, n_pix - q0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < C * kMixPix; idx += kMixThreads) {
      int k, pq;
      if (chan_fast) { pq = idx / C; k = idx - pq * C; } else { k = idx / kMixPix; pq = idx - k * kMixPix; }
      float vx = 0.f, vg = 0.f;
      if (pq < n_here) {
        const int64_t q = q0 + pq, o = q / I, i = q - o * I;
        vx = x[o * x_so + k * x_sc + i * x_si];
        vg = gy[o * y_so + k * y_sc + i * y_si];
      }
      s_x[k * pitch + pq] = vx;
      s_g[k * pitch + pq] = vg;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < n_out; idx += kMixThreads) {
      float acc = 0.f;
      if (idx < C * C) {
        const int c = idx / C, k = idx - c * C;
#pragma unroll 8
        for (int pq = 0; pq < kMixPix; ++pq) acc = fmaf(s_g[c * pitch + pq], s_x[k * pitch + pq], acc);
      } else {
        const int c = idx - C * C;
#pragma unroll 8
        for (int pq = 0; pq < kMixPix; ++pq) acc += s_g[c * pitch + pq];
      }
      s_acc[idx] += acc;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_out; idx += kMixThreads) partial[(int64_t)blockIdx.x * n_out + idx] = s_acc[idx];
}

// Image layout ([B, C, I] with the pixels of a channel contiguous, I a multiple of the tile): the same two-pass scheme
// with tiles of WP = 128 (64) pixels of ONE sample, 256 threads and 16-byte loads.  The 32-pixel tiles of the generic
// kernel above cost two barriers, a 64-bit division per element and an exposed load latency per 32 pixels: 33 us per
// launch for 25 MB at the 32x32 level, where the data move in 5.
template <int WP>
__global__ void __launch_bounds__(256) chanmix_wgrad_planes_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                   float* __restrict__ partial, int n_tiles, int tiles_per_sample,
                                                                   int C, int64_t x_so, int64_t x_sc, int64_t y_so, int64_t y_sc) {
  extern __shared__ float smem[];
  constexpr int pitch = WP + 1;
  const int n_out = C * C + C;
  float* s_acc = smem;                  // [C*C + C]
  float* s_x = s_acc + n_out;           // [C][pitch]
  float* s_g = s_x + C * pitch;         // [C][pitch]
  for (int idx = threadIdx.x; idx < n_out; idx += 256) s_acc[idx] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int o = tile / tiles_per_sample, i0 = (tile - o * tiles_per_sample) * WP;
    const float* xb = x + (int64_t)o * x_so + i0;
    const float* gb = gy + (int64_t)o * y_so + i0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < C * (WP / 4); idx += 256) {
      const int k = idx / (WP / 4), p4 = idx - k * (WP / 4);
      const float4 vx = ldg_stream4(reinterpret_cast<const float4*>(xb + k * x_sc) + p4);
      const float4 vg = ldg_stream4(reinterpret_cast<const float4*>(gb + k * y_sc) + p4);
      float* dx = s_x + k * pitch + 4 * p4;
      float* dg = s_g + k * pitch + 4 * p4;
      dx[0] = vx.x; dx[1] = vx.y; dx[2] = vx.z; dx[3] = vx.w;
      dg[0] = vg.x; dg[1] = vg.y; dg[2] = vg.z; dg[3] = vg.w;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < n_out; idx += 256) {
      float acc = 0.f;
      if (idx < C * C) {
        const int c = idx / C, k = idx - c * C;
        const float* pg = s_g + c * pitch;
        const float* px = s_x + k * pitch;
#pragma unroll 16
        for (int pq = 0; pq < WP; ++pq) acc = fmaf(pg[pq], px[pq], acc);
      } else {
        const float* pg = s_g + (idx - C * C) * pitch;
#pragma unroll 16
        for (int pq = 0; pq < WP; ++pq) acc += pg[pq];
      }
      s_acc[idx] += acc;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_out; idx += 256) partial[(int64_t)blockIdx.x * n_out + idx] = s_acc[idx];
}

// second pass: add the CTA partials (deterministic: fixed assignment, fixed tree).  A block owns 32 consecutive
// outputs; its 32 rows of threads sweep the partial blocks (coalesced 128-byte reads), then the rows are added
// in order through shared memory.
__global__ void __launch_bounds__(1024) chanmix_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ g_mat,
                                                                    float* __restrict__ g_bias, int n_parts, int C) {
  __shared__ float s_red[32][33];
  const int n_out = C * C + C;
  const int idx = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (idx < n_out)
    for (int b = threadIdx.y; b < n_parts; b += 32) acc += partial[(int64_t)b * n_out + idx];
  s_red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && idx < n_out) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += s_red[r][threadIdx.x];
    if (idx < C * C) g_mat[idx] = t; else if (g_bias) g_bias[idx - C * C] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// 2x2 space-to-depth + scalar affine      transforms/reshape.py:29-57, transforms/standard.py:41-57
// ---------------------------------------------------------------------------------------------
__global__ void squeeze_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int C, int H, int W,
                               int inverse, float scale, float shift) {
  // H, W are the LARGE (unsqueezed) spatial sizes; squeezed tensor is [B, 4C, H/2, W/2]
  const int H2 = H >> 1, W2 = W >> 1;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    if (!inverse) {  // idx enumerates the squeezed output (coalesced stores)
      int64_t r = idx;
      const int w2 = r % W2; r /= W2;
      const int h2 = r % H2; r /= H2;
      const int c4 = r % (4 * C); const int64_t b = r / (4 * C);
      const int c = c4 >> 2, fh = (c4 >> 1) & 1, fw = c4 & 1;
      const float v = x[((b * C + c) * H + (2 * h2 + fh)) * W + (2 * w2 + fw)];
      y[idx] = __fadd_rn(__fmul_rn(v, scale), shift);
    } else {  // idx enumerates the unsqueezed output
      int64_t r = idx;
      const int w = r % W; r /= W;
      const int h = r % H; r /= H;
      const int c = r % C; const int64_t b = r / C;
      const int c4 = c * 4 + (h & 1) * 2 + (w & 1);
      const float v = x[((b * 4 * C + c4) * H2 + (h >> 1)) * W2 + (w >> 1)];
      y[idx] = (scale == 1.f && shift == 0.f) ? v : __fdiv_rn(__fsub_rn(v, shift), scale);
    }
  }
}

// AffineScalarTransform alone (transforms/standard.py:41-57): mul-then-add / sub-then-div, like ATen
__global__ void affine_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float scale, float shift,
                                     int inverse) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[idx];
    y[idx] = inverse ? __fdiv_rn(__fsub_rn(v, shift), scale) : __fadd_rn(__fmul_rn(v, scale), shift);
  }
}

__global__ void permute_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ perm, float* __restrict__ y,
                                    int64_t rows, int64_t n) {
  const int64_t total = rows * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = idx / n, j = idx - o * n;
    y[idx] = x[o * n + perm[j]];
  }
}

// ---------------------------------------------------------------------------------------------
// LogTanh                                   transforms/nonlinearities.py:38-95
// ---------------------------------------------------------------------------------------------
struct LogTanhConsts {
  float cut, inv_cut, alpha, beta, log_ab;
};

__device__ __forceinline__ void logtanh_eval(float v, const LogTanhConsts& c, int inverse, float& out, float& lad) {
  if (!inverse) {
    if (v > c.cut) { out = c.alpha * logf(c.beta * v); lad = logf(c.alpha / v); }
    else if (v < -c.cut) { out = c.alpha * -logf(-c.beta * v); lad = logf(-c.alpha / v); }
    else { out = tanhf(v); lad = logf(1.f - out * out); }
  } else {
    if (v > c.inv_cut) { out = expf(v / c.alpha) / c.beta; lad = -c.log_ab + v / c.alpha; }
    else if (v < -c.inv_cut) { out = -expf(-v / c.alpha) / c.beta; lad = -c.log_ab - v / c.alpha; }
    else { out = 0.5f * logf((1.f + v) / (1.f - v)); lad = -logf(1.f - v * v); }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256) logtanh_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                      const float* __restrict__ gl, float* __restrict__ y,
                                                      float* __restrict__ lad_sample, int64_t n, int inverse,
                                                      LogTanhConsts c) {
  __shared__ float red[8];
  const int64_t o = blockIdx.x;
  const float* xo = x + o * n;
  float acc = 0.f;
  const float g_l = BWD ? (gl ? gl[o] : 0.f) : 0.f;
  for (int64_t j = threadIdx.x; j < n; j += 256) {
    const float v = xo[j];
    if (!BWD) {
      float out, lad;
      logtanh_eval(v, c, inverse, out, lad);
      y[o * n + j] = out;
      acc += lad;
    } else {
      // d out / d v and d lad / d v
      float dy, dl;
      if (!inverse) {
        if (v > c.cut || v < -c.cut) { dy = c.alpha / fabsf(v); dl = -1.f / v; }
        else { const float t = tanhf(v); dy = 1.f - t * t; dl = -2.f * t; }
      } else {
        if (v > c.inv_cut) { dy = expf(v / c.alpha) / (c.alpha * c.beta); dl = 1.f / c.alpha; }
        else if (v < -c.inv_cut) { dy = expf(-v / c.alpha) / (c.alpha * c.beta); dl = -1.f / c.alpha; }
        else { dy = 1.f / (1.f - v * v); dl = 2.f * v / (1.f - v * v); }
      }
      y[o * n + j] = (gy ? gy[o * n + j] : 0.f) * dy + g_l * dl;
    }
  }
  if (!BWD && lad_sample) {
    const float s = block_sum<256>(acc, red);
    if (threadIdx.x == 0) lad_sample[o] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Gaussian log-likelihood epilogue  distributions/normal.py:18-23,52-59; flows/manifold_flow.py:125-128
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gauss_logprob_kernel(const float* __restrict__ u, int64_t n_u, int64_t u_stride,
                                                            const float* __restrict__ v, int64_t n_v, int64_t v_stride,
                                                            float inv_eps2, float clip, float log_z_u, float log_z_v,
                                                            const float* __restrict__ add0, const float* __restrict__ add1,
                                                            float* __restrict__ out) {
  __shared__ float red[8];
  const int64_t o = blockIdx.x;
  float su = 0.f, sv = 0.f;
  for (int64_t j = threadIdx.x; j < n_u; j += 256) { const float t = u[o * u_stride + j]; su = fmaf(t, t, su); }
  for (int64_t j = threadIdx.x; j < n_v; j += 256) {
    float t = v[o * v_stride + j];
    if (clip > 0.f) t = fminf(fmaxf(t, -clip), clip);
    sv = fmaf(t, t, sv);
  }
  su = block_sum<256>(su, red);
  sv = block_sum<256>(sv, red);
  if (threadIdx.x == 0) {
    float lp = -0.5f * su - log_z_u;
    if (n_v > 0) lp = lp + (-0.5f * sv * inv_eps2 - log_z_v);
    if (add0) lp = lp + add0[o];
    if (add1) lp = lp + add1[o];
    out[o] = lp;
  }
}

__global__ void gauss_logprob_bwd_kernel(const float* __restrict__ u, int64_t n_u, int64_t u_stride,
                                         const float* __restrict__ v, int64_t n_v, int64_t v_stride, float inv_eps2,
                                         float clip, const float* __restrict__ g, float* __restrict__ g_u,
                                         float* __restrict__ g_v, int64_t rows) {
  const int64_t n = n_u + n_v, total = rows * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = idx / n, j = idx - o * n;
    const float go = g[o];
    if (j < n_u) {
      if (g_u) g_u[o * n_u + j] = -u[o * u_stride + j] * go;
    } else if (g_v) {
      const float t = v[o * v_stride + (j - n_u)];
      const bool clipped = clip > 0.f && (t < -clip || t > clip);
      g_v[o * n_v + (j - n_u)] = clipped ? 0.f : -t * inv_eps2 * go;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (d) log det(J^T J) per sample, J [D, n], n <= 32: one warp per sample.  The Gram matrix is
// accumulated and factorised (LDL^T, no pivoting: it is symmetric positive definite) in fp64 so that
// ill-conditioned Jacobians keep their digits.      flows/manifold_flow.py:140-147
// ---------------------------------------------------------------------------------------------
constexpr int kJtjWarps = 4;
__global__ void __launch_bounds__(kJtjWarps * 32) jtj_logdet_kernel(const float* __restrict__ jac, float* __restrict__ out,
                                                                   int64_t rows, int D, int n) {
  __shared__ double s_g[kJtjWarps][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t o = (int64_t)blockIdx.x * kJtjWarps + warp;
  if (o >= rows) return;
  double* g = s_g[warp];
  const float* J = jac + o * (int64_t)D * n;
  // lane owns column `lane` of the Gram matrix: g[a][lane] = sum_d J[d][a] J[d][lane]
  for (int a = 0; a < n; ++a) {
    double acc = 0.0;
    if (lane < n)
      for (int d = 0; d < D; ++d) acc += (double)J[d * n + a] * (double)J[d * n + lane];
    g[a * 33 + lane] = acc;
  }
  __syncwarp();
  double logdet = 0.0;
  for (int k = 0; k < n; ++k) {
    const double piv = g[k * 33 + k];
    logdet += log(fabs(piv));
    // rank-1 update of the trailing block: lane owns column `lane`
    if (lane > k && lane < n) {
      const double f = g[k * 33 + lane] / piv;
      for (int r = k + 1; r < n; ++r) g[r * 33 + lane] -= g[r * 33 + k] * f;
    }
    __syncwarp();
  }
  if (lane == 0) out[o] = (float)logdet;
}

static unsigned grid_for(int64_t total, int threads) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

static size_t mix_smem(int C) { const int Cp = mix_cp(C); return (size_t)(C * Cp + C * (kMixPix + 1) + Cp) * sizeof(float); }
static size_t wgrad_smem(int C) { return (size_t)(C * C + C + 2 * C * (kMixPix + 1)) * sizeof(float); }
static int wgrad_grid(int64_t n_pix, int64_t n_chan) {
  // as many CTAs as pixel tiles, up to 8 per SM; the partial blocks stay below 8 MB
  int64_t tiles = (n_pix + kMixPix - 1) / kMixPix;
  int64_t cap = (int64_t)kNumSMs * 8;
  const int64_t budget = (2 * 1024 * 1024) / (n_chan * n_chan + n_chan);
  if (cap > budget) cap = budget;
  if (cap < 8) cap = 8;
  return (int)(tiles < 1 ? 1 : (tiles > cap ? cap : tiles));
}

}  // namespace mflow

using namespace mflow;

extern "C" MFLOW_API int mflow_chanmix_apply(const mflow_mix_desc* d, const float* x, const float* mat, const float* bias, float* y,
                                   void* stream) {
  MFLOW_REQUIRE(d && x && mat && y, "null pointer");
  MFLOW_REQUIRE(d->n_chan >= 1 && d->n_chan <= 128, "chanmix supports 1..128 channels, got %lld", (long long)d->n_chan);
  const int64_t n_pix = d->n_outer * d->n_inner;
  if (n_pix == 0) return MFLOW_OK;
  const int C = (int)d->n_chan;
  // planes layout, 16-byte aligned rows, the image levels' channel counts: the register-resident kernel
  if ((C == 12 || C == 24 || C == 48) && d->x_si == 1 && d->y_si == 1 && d->x_sc == d->n_inner && d->y_sc == d->n_inner &&
      d->n_inner % 4 == 0 && d->x_so % 4 == 0 && d->y_so % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0) {
    const cudaStream_t st = (cudaStream_t)stream;
    auto grid_of = [&](int px, int cgroups) {
      int64_t g = (d->n_outer * (d->n_inner / px) * cgroups + 255) / 256;
      return (unsigned)(g > 8 * kNumSMs ? 8 * kNumSMs : g);
    };
    if (C == 12) chanmix_planes_kernel<12, 4, 1><<<grid_of(4, 1), 256, 0, st>>>(x, mat, bias, y, d->n_outer, d->n_inner, d->x_so, d->y_so);
    else if (C == 24) chanmix_planes_kernel<24, 4, 1><<<grid_of(4, 1), 256, 0, st>>>(x, mat, bias, y, d->n_outer, d->n_inner, d->x_so, d->y_so);
    else chanmix_planes_kernel<48, 2, 2><<<grid_of(2, 2), 256, 0, st>>>(x, mat, bias, y, d->n_outer, d->n_inner, d->x_so, d->y_so);
    // (96 channels, 4x4 maps: 16 pixels per image - the tiled kernel below is faster there: 24 vs 43 us)
    return check_launch("chanmix_planes_kernel");
  }
  // dense rows of a vector flow at large batch: the row-tile kernel
  if (d->n_inner == 1 && d->x_sc == 1 && d->y_sc == 1 && d->x_so == C && d->y_so == C && C <= 64 && n_pix >= 4096) {
    const size_t smem_rows = (size_t)(C * mix_cp(C) + mix_cp(C) + 2 * kMixRowTile * (C | 1)) * sizeof(float);
    if (smem_rows > 48 * 1024) cudaFuncSetAttribute(chanmix_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows);
    const int64_t t = (n_pix + kMixRowTile - 1) / kMixRowTile;
    const unsigned g = (unsigned)(t < 4 * kNumSMs ? t : 4 * kNumSMs);
    chanmix_rows_kernel<<<g, 256, smem_rows, (cudaStream_t)stream>>>(x, mat, bias, y, n_pix, C);
    return check_launch("chanmix_rows_kernel");
  }
  const size_t smem = mix_smem(C);
  if (smem > 48 * 1024) cudaFuncSetAttribute(chanmix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t tiles = (n_pix + kMixPix - 1) / kMixPix;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);   // grid-stride over the pixel tiles
  chanmix_kernel<<<grid, kMixFwdThreads, smem, (cudaStream_t)stream>>>(x, mat, bias, y, n_pix, C, d->n_inner, d->x_so, d->x_sc,
                                                                      d->x_si, d->y_so, d->y_sc, d->y_si);
  return check_launch("chanmix_kernel");
}

extern "C" MFLOW_API int64_t mflow_chanmix_wgrad_workspace_bytes(const mflow_mix_desc* d) {
  if (!d) return 0;
  const int64_t n_out = d->n_chan * d->n_chan + d->n_chan;
  return 16 + (int64_t)wgrad_grid(d->n_outer * d->n_inner, d->n_chan) * n_out * 4;
}

extern "C" MFLOW_API int mflow_chanmix_wgrad(const mflow_mix_desc* d, const float* x, const float* g_y, float* g_mat, float* g_bias,
                                   void* workspace, void* stream) {
  MFLOW_REQUIRE(d && x && g_y && g_mat && workspace, "null pointer");
  MFLOW_REQUIRE(d->n_chan >= 1 && d->n_chan <= 96, "chanmix_wgrad supports 1..96 channels, got %lld", (long long)d->n_chan);
  const int64_t n_pix = d->n_outer * d->n_inner;
  const int C = (int)d->n_chan;
  const int grid = wgrad_grid(n_pix, C);
  const int64_t n_out = (int64_t)C * C + C;
  float* partial = (float*)((char*)workspace + 16);
  // image layout with whole tiles of one sample: the 128- / 64-pixel-tile kernel
  const int64_t I = d->n_inner;
  const int wp = (I % 128 == 0) ? 128 : ((I % 64 == 0) ? 64 : 0);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g_y)) & 15) == 0 &&
                       ((d->x_so | d->x_sc | d->y_so | d->y_sc) & 3) == 0;
  const size_t smem_planes = wp ? (size_t)(n_out + 2 * C * (wp + 1)) * sizeof(float) : 0;
  if (wp && d->x_si == 1 && d->y_si == 1 && aligned && smem_planes <= 48 * 1024 && n_pix / wp < (1ll << 31)) {
    const int n_tiles = (int)(n_pix / wp), tps = (int)(I / wp);
    const int grid_p = n_tiles < grid ? n_tiles : grid;   // never more partial blocks than the workspace was sized for
    if (wp == 128) chanmix_wgrad_planes_kernel<128><<<grid_p, 256, smem_planes, (cudaStream_t)stream>>>(x, g_y, partial, n_tiles, tps, C, d->x_so,
                                                                                                        d->x_sc, d->y_so, d->y_sc);
    else chanmix_wgrad_planes_kernel<64><<<grid_p, 256, smem_planes, (cudaStream_t)stream>>>(x, g_y, partial, n_tiles, tps, C, d->x_so, d->x_sc,
                                                                                             d->y_so, d->y_sc);
    chanmix_wgrad_reduce_kernel<<<(unsigned)((n_out + 31) / 32), dim3(32, 32), 0, (cudaStream_t)stream>>>(partial, g_mat, g_bias, grid_p, C);
    return check_launch("chanmix_wgrad_planes_kernel");
  }
  const size_t smem = wgrad_smem(C);
  if (smem > 48 * 1024) cudaFuncSetAttribute(chanmix_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  chanmix_wgrad_kernel<<<grid, kMixThreads, smem, (cudaStream_t)stream>>>(x, g_y, g_mat, g_bias, partial, n_pix, C, d->n_inner,
                                                                            d->x_so, d->x_sc, d->x_si, d->y_so, d->y_sc, d->y_si);
  chanmix_wgrad_reduce_kernel<<<(unsigned)((n_out + 31) / 32), dim3(32, 32), 0, (cudaStream_t)stream>>>(partial, g_mat, g_bias, grid, C);
  return check_launch("chanmix_wgrad_kernel");
}

// ---- per-channel sum over [B, C, I] (bias gradient of a convolution) ---------------------------
// Pass 1: CTA (c, s) adds the images b = s, s + S, ... of channel c (each a contiguous run of I floats) and
// writes partial[c][s]; pass 2 adds the S partials of a channel in index order.  Deterministic.
constexpr int kSumThreads = 256;
__global__ void __launch_bounds__(kSumThreads) channel_sum_partial_kernel(const float* __restrict__ x, float* __restrict__ partial,
                                                                          int B, int C, int I, int S) {
  __shared__ float red[kSumThreads / 32];
  const int c = blockIdx.x, s = blockIdx.y;
  float acc = 0.f;
  if ((I & 3) == 0) {
    const int i4n = I >> 2;
    // threads sweep (image, float4) pairs: consecutive threads read consecutive float4s of one image
    const int n_img = (B - s + S - 1) / S;
    for (int j = threadIdx.x; j < n_img * i4n; j += kSumThreads) {
      const int bi = j / i4n, i4 = j - bi * i4n;
      const int b = s + bi * S;
      const float4 v = ldg_stream4(reinterpret_cast<const float4*>(x + ((int64_t)b * C + c) * I) + i4);
      acc += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    const int n_img = (B - s + S - 1) / S;
    for (int j = threadIdx.x; j < n_img * I; j += kSumThreads) {
      const int bi = j / I, i = j - bi * I;
      acc += x[((int64_t)(s + bi * S) * C + c) * I + i];
    }
  }
  acc = block_sum<kSumThreads>(acc, red);
  if (threadIdx.x == 0) partial[c * S + s] = acc;
}
__global__ void channel_sum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int C, int S) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += partial[c * S + s];
  out[c] = acc;
}
static int channel_sum_splits(int64_t batch, int64_t chan) {
  int64_t s = (4 * 148 + chan - 1) / chan;   // about four waves of CTAs
  if (s > batch) s = batch;
  if (s < 1) s = 1;
  return (int)s;
}
extern "C" MFLOW_API int64_t mflow_channel_sum_workspace_bytes(int64_t batch, int64_t chan, int64_t inner) {
  (void)inner;
  return chan * channel_sum_splits(batch, chan) * (int64_t)sizeof(float);
}
extern "C" MFLOW_API int mflow_channel_sum(const float* x, float* out, int64_t batch, int64_t chan, int64_t inner, void* workspace,
                                           void* stream) {
  MFLOW_REQUIRE(out && chan >= 1 && batch >= 0 && inner >= 0, "bad arguments");
  MFLOW_REQUIRE(chan <= 65535 * 32 && batch * chan * inner < ((int64_t)1 << 40), "tensor too large");
  if (batch == 0 || inner == 0) {
    return cudaMemsetAsync(out, 0, chan * sizeof(float), (cudaStream_t)stream) == cudaSuccess ? MFLOW_OK : MFLOW_E_LAUNCH;
  }
  MFLOW_REQUIRE(x && workspace, "null pointer");
  MFLOW_REQUIRE(batch <= INT32_MAX && inner <= (1 << 24), "extent too large");
  const int S = channel_sum_splits(batch, chan);
  dim3 grid((unsigned)chan, (unsigned)S);
  channel_sum_partial_kernel<<<grid, kSumThreads, 0, (cudaStream_t)stream>>>(x, (float*)workspace, (int)batch, (int)chan, (int)inner, S);
  channel_sum_final_kernel<<<(unsigned)((chan + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const float*)workspace, out, (int)chan, S);
  return check_launch("channel_sum_kernel");
}

extern "C" MFLOW_API int mflow_squeeze2x2(const float* x, float* y, int64_t batch, int64_t chan, int64_t height, int64_t width,
                                int32_t inverse, float scale, float shift, void* stream) {
  MFLOW_REQUIRE(x && y, "null pointer");
  // forward: x is [B, C, H, W]; inverse: x is [B, 4C', H', W'] and chan/height/width describe x as given
  int64_t C, H, W;
  if (!inverse) {
    MFLOW_REQUIRE(height % 2 == 0 && width % 2 == 0, "Input image size not compatible with the factor.");
    C = chan; H = height; W = width;
  } else {
    MFLOW_REQUIRE(chan >= 4 && chan % 4 == 0, "Invalid number of channel dimensions.");
    C = chan / 4; H = height * 2; W = width * 2;
  }
  MFLOW_REQUIRE(scale != 0.f, "Scale cannot be zero.");
  const int64_t total = batch * C * H * W;
  if (total == 0) return MFLOW_OK;
  squeeze_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, total, (int)C, (int)H, (int)W, inverse, scale,
                                                                          shift);
  return check_launch("squeeze_kernel");
}

extern "C" MFLOW_API int mflow_affine_scalar(const float* x, float* y, int64_t n, float scale, float shift, int32_t inverse,
                                             void* stream) {
  MFLOW_REQUIRE(x && y, "null pointer");
  MFLOW_REQUIRE(scale != 0.f, "Scale cannot be zero.");
  if (n == 0) return MFLOW_OK;
  affine_scalar_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, scale, shift, inverse);
  return check_launch("affine_scalar_kernel");
}

extern "C" MFLOW_API int mflow_permute_rows(const float* x, const int64_t* perm, float* y, int64_t rows, int64_t n, void* stream) {
  MFLOW_REQUIRE(x && perm && y, "null pointer");
  if (rows * n == 0) return MFLOW_OK;
  permute_rows_kernel<<<grid_for(rows * n, 256), 256, 0, (cudaStream_t)stream>>>(x, perm, y, rows, n);
  return check_launch("permute_rows_kernel");
}

static LogTanhConsts logtanh_consts(float cut_point) {
  // transforms/nonlinearities.py:49-53 (float64 on the host)
  const double cut = cut_point;
  const double inv_cut = tanh(cut);
  const double alpha = (1.0 - tanh(tanh(cut))) / cut;
  const double beta = exp((tanh(cut) - alpha * log(cut)) / alpha);
  LogTanhConsts c;
  c.cut = (float)cut; c.inv_cut = (float)inv_cut; c.alpha = (float)alpha; c.beta = (float)beta;
  c.log_ab = (float)log(alpha * beta);
  return c;
}

extern "C" MFLOW_API int mflow_logtanh_apply(const float* x, float* y, float* lad_sample, int64_t rows, int64_t n, int32_t inverse,
                                   float cut_point, void* stream) {
  MFLOW_REQUIRE(x && y, "null pointer");
  MFLOW_REQUIRE(cut_point > 0.f, "Cut point must be positive.");
  if (rows == 0) return MFLOW_OK;
  logtanh_kernel<false><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, y, lad_sample, n, inverse,
                                                                           logtanh_consts(cut_point));
  return check_launch("logtanh_kernel");
}

extern "C" MFLOW_API int mflow_logtanh_backward(const float* x, const float* g_y, const float* g_lad_sample, float* g_x, int64_t rows,
                                      int64_t n, int32_t inverse, float cut_point, void* stream) {
  MFLOW_REQUIRE(x && g_x, "null pointer");
  if (rows == 0) return MFLOW_OK;
  logtanh_kernel<true><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, g_y, g_lad_sample, g_x, nullptr, n, inverse,
                                                                          logtanh_consts(cut_point));
  return check_launch("logtanh_kernel(bwd)");
}

namespace mflow {
// Vector models: a few latent features per row (2 + 1 in cfg1) and up to 10^6 rows - one THREAD per row instead of one
// 256-thread CTA with two block reductions per row (2 ms for 2^20 rows of three floats).
__global__ void __launch_bounds__(256) gauss_logprob_rows_kernel(const float* __restrict__ u, int n_u, int64_t u_stride,
                                                                 const float* __restrict__ v, int n_v, int64_t v_stride, float inv_eps2,
                                                                 float clip, float log_z_u, float log_z_v, const float* __restrict__ add0,
                                                                 const float* __restrict__ add1, float* __restrict__ out, int64_t rows) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= rows) return;
  float su = 0.f, sv = 0.f;
  for (int j = 0; j < n_u; ++j) { const float t = u[o * u_stride + j]; su = fmaf(t, t, su); }
  for (int j = 0; j < n_v; ++j) {
    float t = v[o * v_stride + j];
    if (clip > 0.f) t = fminf(fmaxf(t, -clip), clip);
    sv = fmaf(t, t, sv);
  }
  float lp = -0.5f * su - log_z_u;
  if (n_v > 0) lp = lp + (-0.5f * sv * inv_eps2 - log_z_v);
  if (add0) lp = lp + add0[o];
  if (add1) lp = lp + add1[o];
  out[o] = lp;
}
}  // namespace mflow

extern "C" MFLOW_API int mflow_gauss_logprob(const float* u, int64_t n_u, int64_t u_stride, const float* v, int64_t n_v,
                                   int64_t v_stride, float eps, float clip, const float* add0, const float* add1,
                                   float* log_prob, int64_t rows, void* stream) {
  MFLOW_REQUIRE((u || n_u == 0) && log_prob, "null pointer");
  MFLOW_REQUIRE(n_v == 0 || (v != nullptr && eps > 0.f), "orthogonal part needs a pointer and eps > 0");
  if (rows == 0) return MFLOW_OK;
  const double two_pi = 6.283185307179586;
  const float log_z_u = (float)(0.5 * (double)n_u * log(two_pi));
  const float log_z_v = n_v ? (float)(0.5 * (double)n_v * log(two_pi) + (double)n_v * log((double)eps)) : 0.f;
  const float inv_eps2 = n_v ? (float)(1.0 / ((double)eps * (double)eps)) : 0.f;
  if (n_u + n_v <= 64 && rows >= 4096) {
    gauss_logprob_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, (int)n_u, u_stride, v, (int)n_v, v_stride,
                                                                                                  inv_eps2, clip, log_z_u, log_z_v, add0, add1,
                                                                                                  log_prob, rows);
    return check_launch("gauss_logprob_rows_kernel");
  }
  gauss_logprob_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(u, n_u, u_stride, v, n_v, v_stride, inv_eps2, clip,
                                                                          log_z_u, log_z_v, add0, add1, log_prob);
  return check_launch("gauss_logprob_kernel");
}

extern "C" MFLOW_API int mflow_gauss_logprob_backward(const float* u, int64_t n_u, int64_t u_stride, const float* v, int64_t n_v,
                                            int64_t v_stride, float eps, float clip, const float* g_log_prob, float* g_u,
                                            float* g_v, int64_t rows, void* stream) {
  MFLOW_REQUIRE((u || n_u == 0) && g_log_prob, "null pointer");
  if (rows == 0) return MFLOW_OK;
  const float inv_eps2 = n_v ? (float)(1.0 / ((double)eps * (double)eps)) : 0.f;
  gauss_logprob_bwd_kernel<<<grid_for(rows * (n_u + n_v), 256), 256, 0, (cudaStream_t)stream>>>(
      u, n_u, u_stride, v, n_v, v_stride, inv_eps2, clip, g_log_prob, g_u, g_v, rows);
  return check_launch("gauss_logprob_bwd_kernel");
}

extern "C" MFLOW_API int mflow_jtj_logdet(const float* jac, float* logdet, int64_t rows, int32_t d, int32_t n, void* stream) {
  MFLOW_REQUIRE(jac && logdet, "null pointer");
  MFLOW_REQUIRE(n >= 1 && n <= 32, "jtj_logdet supports latent dimension 1..32, got %d", (int)n);
  MFLOW_REQUIRE(d >= n, "J must have at least as many rows as columns");
  if (rows == 0) return MFLOW_OK;
  jtj_logdet_kernel<<<(unsigned)((rows + kJtjWarps - 1) / kJtjWarps), kJtjWarps * 32, 0, (cudaStream_t)stream>>>(jac, logdet,
                                                                                                                rows, d, n);
  return check_launch("jtj_logdet_kernel");
}

// ---------------------------------------------------------------------------------------------
// Gated residual of a ResidualBlock with context (nn/resnet.py:36-40): glu(cat(t, g)) = t * sigmoid(g), out = h + that.
// The large-batch vector conditioners (ResidualNet on the convolution kernels) ran it as sigmoid / mul / add launches
// forward and five more backward, each a round trip of a [rows, 100] activation, plus an absmax pass for the next
// convolution's operand scale: one launch each way here, with the absolute maxima published on the side.
// ---------------------------------------------------------------------------------------------
namespace mflow {
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float absmax4(float m, const float4& v) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
__device__ __forceinline__ void publish_max(float* slot, float m) {   // all lanes call; slot zero-initialised by the caller
  const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  if ((threadIdx.x & 31) == 0 && w > *reinterpret_cast<volatile uint32_t*>(slot)) atomicMax(reinterpret_cast<unsigned int*>(slot), w);
}

__global__ void __launch_bounds__(256) glu_residual_kernel(const float4* __restrict__ h, const float4* __restrict__ t,
                                                           const float4* __restrict__ g, float4* __restrict__ out, float* out_amax,
                                                           int64_t n4) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 hv = ldg_stream4(h + i), tv = ldg_stream4(t + i), gv = ldg_stream4(g + i);
    float4 o;
    o.x = hv.x + tv.x * sigmoidf_(gv.x); o.y = hv.y + tv.y * sigmoidf_(gv.y);
    o.z = hv.z + tv.z * sigmoidf_(gv.z); o.w = hv.w + tv.w * sigmoidf_(gv.w);
    out[i] = o;
    m = absmax4(m, o);
  }
  if (out_amax != nullptr) publish_max(out_amax, m);
}

__global__ void __launch_bounds__(256) glu_residual_backward_kernel(const float4* __restrict__ t, const float4* __restrict__ g,
                                                                    const float4* __restrict__ go, float4* __restrict__ g_t,
                                                                    float4* __restrict__ g_g, float* g_t_amax, float* g_g_amax,
                                                                    int64_t n4) {
  float mt = 0.f, mg = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 tv = ldg_stream4(t + i), gv = ldg_stream4(g + i), ov = ldg_stream4(go + i);
    const float sx = sigmoidf_(gv.x), sy = sigmoidf_(gv.y), sz = sigmoidf_(gv.z), sw = sigmoidf_(gv.w);
    const float4 dt = make_float4(ov.x * sx, ov.y * sy, ov.z * sz, ov.w * sw);
    // d sigmoid = s (1 - s), in the order torch's sigmoid_backward evaluates it: grad * ((1 - s) * s)
    const float4 dg = make_float4((ov.x * tv.x) * ((1.f - sx) * sx), (ov.y * tv.y) * ((1.f - sy) * sy),
                                  (ov.z * tv.z) * ((1.f - sz) * sz), (ov.w * tv.w) * ((1.f - sw) * sw));
    g_t[i] = dt;
    g_g[i] = dg;
    mt = absmax4(mt, dt);
    mg = absmax4(mg, dg);
  }
  if (g_t_amax != nullptr) publish_max(g_t_amax, mt);
  if (g_g_amax != nullptr) publish_max(g_g_amax, mg);
}
}  // namespace mflow

static unsigned glu_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  if (b > 8 * kNumSMs) b = 8 * kNumSMs;
  return (unsigned)(b < 1 ? 1 : b);
}

extern "C" MFLOW_API int mflow_glu_residual(const float* h, const float* t, const float* g, float* out, float* out_amax, int64_t n,
                                            void* stream) {
  MFLOW_REQUIRE(h && t && g && out && n >= 0, "null pointer");
  MFLOW_REQUIRE((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(g) |
                                  reinterpret_cast<uintptr_t>(out)) & 15) == 0, "n must be a multiple of 4 and the tensors 16-byte aligned");
  if (n == 0) return MFLOW_OK;
  glu_residual_kernel<<<glu_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)h, (const float4*)t, (const float4*)g, (float4*)out,
                                                                          out_amax, n / 4);
  return check_launch("glu_residual_kernel");
}

extern "C" MFLOW_API int mflow_glu_residual_backward(const float* t, const float* g, const float* g_out, float* g_t, float* g_g,
                                                     float* g_t_amax, float* g_g_amax, int64_t n, void* stream) {
  MFLOW_REQUIRE(t && g && g_out && g_t && g_g && n >= 0, "null pointer");
  MFLOW_REQUIRE((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(g_out) |
                                  reinterpret_cast<uintptr_t>(g_t) | reinterpret_cast<uintptr_t>(g_g)) & 15) == 0,
                "n must be a multiple of 4 and the tensors 16-byte aligned");
  if (n == 0) return MFLOW_OK;
  glu_residual_backward_kernel<<<glu_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)t, (const float4*)g, (const float4*)g_out,
                                                                                   (float4*)g_t, (float4*)g_g, g_t_amax, g_g_amax, n / 4);
  return check_launch("glu_residual_backward_kernel");
}

// out = a + b with max |out| published: the gradient accumulation of a residual block's skip connection.  autograd would
// add the two gradients with an ATen kernel and the consuming convolution would then need an absmax pass over the sum.
namespace mflow {
__global__ void __launch_bounds__(256) add_amax_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                                                       float* out_amax, int64_t n4) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = ldg_stream4(a + i), y = ldg_stream4(b + i);
    const float4 o = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    out[i] = o;
    m = absmax4(m, o);
  }
  if (out_amax != nullptr) publish_max(out_amax, m);
}
}  // namespace mflow

extern "C" MFLOW_API int mflow_add_amax(const float* a, const float* b, float* out, float* out_amax, int64_t n, void* stream) {
  MFLOW_REQUIRE(a && b && out && n >= 0, "null pointer");
  MFLOW_REQUIRE((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "n must be a multiple of 4 and the tensors 16-byte aligned");
  if (n == 0) return MFLOW_OK;
  add_amax_kernel<<<glu_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)a, (const float4*)b, (float4*)out, out_amax, n / 4);
  return check_launch("add_amax_kernel");
}
