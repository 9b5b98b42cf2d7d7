// Weight gradient of the conditioner convolutions on the 5th-generation tensor cores.
//
//   g_w[co, ci, ky, kx] = sum_{b, h, w} g_y[b, co, h, w] * relu?(x)[b, ci, h + ky - 1, w + kx - 1]
//
// (autograd of nn.Conv2d 3x3 pad 1 / 1x1 in ConvResidualNet, reference manifold_flow/nn/resnet.py:85-142), as a
// GEMM whose reduction dimension K is the PIXELS of the batch:
//
//   D_tap[ci, co] += P_tap[ci, 32 px] * Q[co, 32 px]^T        per 32-pixel block, per tap
//
// * Both tensors are NCHW, i.e. already "K-major" (the pixels of one channel of one image are contiguous): no
//   im2col, no transposition pass.  Q (the output gradient) lands in shared memory as a 128-byte-swizzled
//   K-major tile straight from one TMA load and is split in place into its TF32 hi / lo parts.
// * P (the activations) is the operand that needs the 3x3 shifts.  It goes to TENSOR MEMORY with lane = input
//   channel and columns = the 32 pixels of the block: the transform thread of a channel holds the block's
//   image rows INCLUDING the left / right halo in registers and stores, per horizontal tap, the window
//   shifted by that tap (a TMEM operand address must be a multiple of 4 columns, so the shift cannot be a
//   column offset of the MMA - measured: "misaligned address").  The vertical tap is the row the TMA box is
//   taken from; out-of-bounds rows and columns are TMA zero fill = the convolution's padding.  ReLU and the
//   hi / lo split happen on the way.
// * Error-compensated split, D += P_hi Q_hi + P_lo Q_hi + P_hi Q_lo with fp32 accumulation in TMEM (fp32-level
//   accuracy), in TF32 containers (kind::tf32, K = 8) or - template parameter F16, the default path - in fp16 halves
//   (kind::f16, K = 16: the same 11 significant bits at twice the tensor-core rate).  fp16's 5-bit exponent is dealt with
//   by scaling each operand with a power of two derived from its absolute maximum (x_amax / g_amax, published by the
//   kernels that produced the tensors); the scales are undone when the partials are written.
// * Split-K: CTA (ks, dy) reduces its share of the pixel blocks for the three taps of kernel row dy into
//   three TMEM accumulators and writes a partial [tap][ci][co]; a second kernel adds the partials in index
//   order (deterministic) and writes g_w in the reference's [co][ci][ky][kx] layout.
// * Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 P transform (two sets of four warps) and epilogue,
//   warps 10-13 Q split.  fp16 path: a set owns every other pixel BLOCK, converts the block's pixels (halo included)
//   ONCE - each becomes one register {fp16 hi, fp16 lo} - and assembles the three horizontally shifted windows with one
//   byte-permute per pixel pair into a ring of five tensor-memory stages (one per (block, tap) work item).  Splitting
//   inside every (block, tap) item made the kernel instruction-issue bound: 2250 cycles per block against 1008 of MMAs
//   (per-block trace, tools/wgrad_trace.py).  TF32 path: the sets alternate work items, one stage each.
#include <stdlib.h>

#include "mflow_common.cuh"
#include "tc_ptx.cuh"
#include "tma_encode.cuh"

namespace mflow {

using namespace ptx;

constexpr int kWgThreads = 32 * 14;
constexpr int kWgMaxPStages = 5;                // P stages in tensor memory: 5 (fp16: 32 columns each) or 2 (TF32: 64 each)
constexpr int kWgBlockPx = 32;                 // pixels per K block
constexpr int kWgQBytes = 128 * kWgBlockPx * 4;  // 16 KB: one Q operand tile (hi or lo), up to 128 rows
constexpr int kWgPCols = 32;                   // TMEM columns of one P part (hi or lo): the block's 32 pixels
constexpr int kWgRawSlots = 4, kWgQSlots = 6;   // at most: a.n_raw (see launch_wgrad) and a.n_q (what shared memory allows) are used

struct WgradArgs {
  float* partial;   // [ks][taps][Cin][Cout]
  float* bias_partial;   // [ks][Cout] or null: per-split sums of g_y over the pixels (the bias gradient comes for free)
  float* direct_w;       // 1x1 kernel with a single K split: the epilogue writes g_w[co][ci] itself (no reduction pass)
  int B, Cin, Cout, H, W;
  int taps;         // 9 or 1
  int relu_in;
  int rows;         // image rows per 32-pixel block (32 / W)
  int box_w;        // raw activation box width: W + 8 (3x3, starts 4 pixels left: TMA needs 16-byte aligned starts) or W
  int raw_bytes;    // one raw activation box [images][128 ci][rows][box_w] fp32
  int n_raw;        // raw-box ring depth (2, 3 or 4: see launch_wgrad)
  int n_q;          // Q-tile ring depth (3 .. kWgQSlots): the chain TMA -> split -> MMA -> slot free is ~3000 cycles long
  int n_tile;       // MMA N extent (output channels of this launch's q tile, multiple of 16, <= 128)
  int blocks_per_img;  // H * W / 32
  int n_blocks;     // B * blocks_per_img
  int n_dy;         // 3 (3x3) or 1
  int tmem_cols;
  const float* x_amax;   // fp16 path: device scalars with max |x| / max |g_y| (null = no scaling)
  const float* g_amax;
};

__device__ __forceinline__ float wg_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float wg_round(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

#ifdef MFLOW_CONV_ABLATE
// dev-only per-block stamps of CTA (ks = 20, dy = 1, tile 0): [0] producer: raw slot free, [1] producer: Q slot free,
// [2] Q split: raw tile arrived, [3] Q split: done, [4] P set 0: raw box arrived, [5] MMA: Q ready, [6] MMA: block issued,
// [7] P transform of tap 0 done
__device__ long long g_wg_steps[8][64];
#define MFLOW_WST(row, idx, cond) do { if (blockIdx.x == 20 && blockIdx.y == (gridDim.y > 1 ? 1 : 0) && blockIdx.z == 0 && (cond) && (idx) < 64) g_wg_steps[row][idx] = clock64(); } while (0)
#else
#define MFLOW_WST(row, idx, cond) do { } while (0)
#endif

struct WgRing {
  int slot, phase, n;
  __device__ __forceinline__ WgRing(int depth) : slot(0), phase(0), n(depth) {}
  __device__ __forceinline__ void next() { if (++slot == n) { slot = 0; phase ^= 1; } }
};

// fp16 path: where the split Q tiles live inside a Q slot (the TMA-loaded fp32 tile occupies the first 16 KB)
constexpr int kWgQHiOff = kWgQBytes, kWgQLoOff = kWgQBytes + kWgQBytes / 2;

template <int kW, int kTaps, bool F16>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_split3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_q = smem;                                   // kWgQSlots x {Q_hi 16 KB, Q_lo 16 KB}
  uint8_t* s_p = s_q + a.n_q * 2 * kWgQBytes;            // a.n_raw x raw activation box
  __shared__ __align__(8) uint64_t rawp_full[kWgRawSlots], rawp_empty[kWgRawSlots], qraw_full[kWgQSlots], q_full[kWgQSlots],
      q_empty[kWgQSlots], p_full[kWgMaxPStages], p_empty[kWgMaxPStages], acc_full;
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = K split, blockIdx.y = kernel row dy, blockIdx.z = (q tile, p tile)
  const int ks = blockIdx.x, n_ks = gridDim.x;
  const int dy = blockIdx.y;
  const int p_tiles = (a.Cin + 127) / 128;
  const int p0 = (blockIdx.z % p_tiles) * 128, q0 = (blockIdx.z / p_tiles) * 128;
  const int per = a.n_blocks / n_ks, extra = a.n_blocks % n_ks;
  const int blk0 = ks * per + min(ks, extra), n_my = per + (ks < extra ? 1 : 0);
  constexpr int taps_here = kTaps == 9 ? 3 : 1;
  constexpr int pad = kTaps == 9 ? 1 : 0;
  constexpr int kPPart = F16 ? kWgPCols / 2 : kWgPCols;   // TMEM columns of one part (hi or lo) of a P stage
  constexpr int kPStages = F16 ? 5 : 2;                   // all that fits beside three accumulators of <= 112 columns

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_g);
    // a raw box is released by the warps that read it: one set (fp16: it owns the block; TF32 1x1: it owns the block's
    // only work item) or both (TF32 3x3: three items over two sets).  A set that skips a block must not arrive: skipping
    // is not gated by the box's arrival, so it could arrive for a later round of the slot inside the current phase.
    for (int i = 0; i < kWgRawSlots; ++i) { mbar_init(&rawp_full[i], 1); mbar_init(&rawp_empty[i], (!F16 && kTaps == 9) ? 8 : 4); }
    for (int i = 0; i < kWgQSlots; ++i) { mbar_init(&qraw_full[i], 1); mbar_init(&q_full[i], 4); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < kWgMaxPStages; ++i) { mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1); }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&s_tmem, a.tmem_cols);  // taps_here x n_tile accumulator columns, then 2 x {P_hi, P_lo} x 32
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t tmem_p = tmem + taps_here * a.n_tile;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      WgRing rp(a.n_raw), rq(a.n_q);
      for (int i = 0; i < n_my; ++i, rp.next(), rq.next()) {
        const int blk = blk0 + i;
        // image and 32-pixel block within it; 4x4 maps: one block = two whole images
        const int b = kW == 4 ? 2 * blk : blk / a.blocks_per_img, r = kW == 4 ? 0 : blk - b * a.blocks_per_img;
        const int h0 = r * a.rows;
        mbar_wait(&rawp_empty[rp.slot], rp.phase ^ 1);
        MFLOW_WST(0, i, true);
        mbar_expect_tx(&rawp_full[rp.slot], a.raw_bytes);
        // 1x1: the block's 32 pixels of a channel are contiguous in NCHW (two whole images for 4x4 maps), and the tensor is
        // described as [B, C, 1, H*W] - one 128-byte (64-byte) TMA row per channel instead of rows x W-float pieces
        if (kTaps == 1) tma_load_4d(s_p + rp.slot * a.raw_bytes, &map_x, &rawp_full[rp.slot], r * kWgBlockPx, 0, p0, b);
        else tma_load_4d(s_p + rp.slot * a.raw_bytes, &map_x, &rawp_full[rp.slot], -4, h0 + dy - 1, p0, b);
        mbar_wait(&q_empty[rq.slot], rq.phase ^ 1);
        MFLOW_WST(1, i, true);
        mbar_expect_tx(&qraw_full[rq.slot], a.n_tile * kWgBlockPx * 4);
        tma_load_3d(s_q + rq.slot * 2 * kWgQBytes, &map_g, &qraw_full[rq.slot], r * kWgBlockPx, q0, b);   // 4x4: box {16 px, n_tile, 2 images}
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = F16 ? idesc_f16(128, a.n_tile, 0, 0) : idesc_tf32(128, a.n_tile, /*a K-major*/ 0, /*b K-major*/ 0);
      WgRing rq(a.n_q);
      int sp = 0, pp = 0;   // P stage and its phase; one stage per (block, tap) work item
      for (int i = 0; i < n_my; ++i, rq.next()) {
        mbar_wait(&q_full[rq.slot], rq.phase);
        MFLOW_WST(5, i, true);
        const uint32_t qhi = smem_u32(s_q + rq.slot * 2 * kWgQBytes);
        // 8..32-wide maps: one K-major tile [n_tile][128 B], 128B swizzle.  4x4 maps: the TMA box {16 px, n_tile, 2 images}
        // lands as two tiles [n_tile][64 B] (64B swizzle), k-steps 0,1 in the first image's tile and 2,3 in the second's.
        // fp16: the Q-split warps wrote ONE K-major tile [n_tile][32 px x 2 B = 64 B] (64B swizzle) per part, whatever kW
        auto q_desc = [&](uint32_t base, int kk) -> uint64_t {
          if (F16) return smem_desc(base + kk * 32, 16, 512, 4);
          if (kW == 4) return smem_desc(base + (kk >> 1) * (a.n_tile * 64) + (kk & 1) * 32, 16, 512, 4);
          return smem_desc(base + kk * 32, 16, 1024, 2);
        };
        const uint32_t q_hi_base = F16 ? qhi + kWgQHiOff : qhi, q_lo_base = F16 ? qhi + kWgQLoOff : qhi + kWgQBytes;
#pragma unroll
        for (int t = 0; t < taps_here; ++t) {
          mbar_wait(&p_full[sp], pp);
          tc_fence_after();
          const uint32_t acc = tmem + t * a.n_tile;
          const uint32_t phi = tmem_p + sp * 2 * kPPart, plo = phi + kPPart;
#pragma unroll
          for (int kk = 0; kk < (F16 ? 2 : 4); ++kk) {   // 32 pixels = 2 x K16 (fp16) or 4 x K8 (TF32)
            const uint64_t d_qhi = q_desc(q_hi_base, kk), d_qlo = q_desc(q_lo_base, kk);
            if (F16) {
              umma_f16_ts(acc, phi + kk * 8, d_qhi, idesc, (i == 0 && kk == 0) ? 0u : 1u);
              umma_f16_ts(acc, plo + kk * 8, d_qhi, idesc, 1);
              umma_f16_ts(acc, phi + kk * 8, d_qlo, idesc, 1);
            } else {
              umma_tf32_ts(acc, phi + kk * 8, d_qhi, idesc, (i == 0 && kk == 0) ? 0u : 1u);
              umma_tf32_ts(acc, plo + kk * 8, d_qhi, idesc, 1);
              umma_tf32_ts(acc, phi + kk * 8, d_qlo, idesc, 1);
            }
          }
          umma_commit(&p_empty[sp]);
          if (++sp == kPStages) { sp = 0; pp ^= 1; }
        }
        umma_commit(&q_empty[rq.slot]);
        MFLOW_WST(6, i, true);
      }
      umma_commit(&acc_full);
    }
  } else if (warp < 10) {
    // ===== P transform: TMEM lane = input channel; two sets of four warps (warps 2-5 / 6-9) =====
    const int quad = warp & 3, set = (warp - 2) >> 2;
    const int ch = quad * 32 + lane;                          // channel within the p tile = TMEM lane
    const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
    float sx = 1.f, inv_sx = 1.f, sg = 1.f, inv_sg = 1.f;   // fp16 path: power-of-two operand scales
    if (F16) { pow2_scale_from_amax(a.x_amax, sx, inv_sx); pow2_scale_from_amax(a.g_amax, sg, inv_sg); }

    WgRing rp(a.n_raw);
    constexpr bool kPad = kTaps == 9;
    constexpr int kImgs = kW == 4 ? 2 : 1;                       // images per 32-pixel block
    constexpr int kRows = kWgBlockPx / kW / kImgs, kBoxW = kPad ? kW + 8 : kW;   // rows per image in the block
    constexpr int kImgVec = kRows * kBoxW / 4;                   // float4s per (image, channel) in the raw box
    constexpr int kRawVec = kImgs * kImgVec;
    auto load_raw = [&](float (&raw)[kRawVec * 4]) {
      // the channel's rows of the block (halo included) -> registers with aligned 16-byte loads
      mbar_wait(&rawp_full[rp.slot], rp.phase);
      const float4* src = reinterpret_cast<const float4*>(s_p + rp.slot * a.raw_bytes) + ch * kImgVec;   // box [image][channel][row][w]
#pragma unroll
      for (int j = 0; j < kRawVec; ++j) {
        const float4 v = src[(j / kImgVec) * 128 * kImgVec + (j % kImgVec)];
        raw[4 * j] = v.x; raw[4 * j + 1] = v.y; raw[4 * j + 2] = v.z; raw[4 * j + 3] = v.w;
      }
    };
    if (F16) {
      // set s owns the blocks of parity s; the (block, tap) items of ALL blocks walk the stage ring in order
      int stage = 0, sphase = 0;
      auto advance = [&](int n) { stage += n; if (stage >= kPStages) { stage -= kPStages; sphase ^= 1; } };
      constexpr int kRowsTot = kRows * kImgs, kCvW = kPad ? kW + 2 : kW;   // converted pixels per row: the row and its halo
      for (int i = 0; i < n_my; ++i, rp.next()) {
        if ((i & 1) != set) { advance(taps_here); continue; }
        float raw[kRawVec * 4];
        load_raw(raw);
        MFLOW_WST(4, i, set == 0 && quad == 0 && lane == 0);
        uint32_t cv[kRowsTot * kCvW];                        // per pixel: {fp16 hi (low half), fp16 lo (high half)}
#pragma unroll
        for (int r = 0; r < kRowsTot; ++r)
#pragma unroll
          for (int j = 0; j < kCvW; ++j) {
            float v = raw[r * kBoxW + (kPad ? 3 : 0) + j];
            if (a.relu_in) v = fmaxf(v, 0.f);
            v *= sx;
            const float hf = __half2float(__float2half_rn(v));
            const __half2 pk = __floats2half2_rn(v, v - hf);   // low half = fp16(v) again, high half = fp16(v - hi)
            cv[r * kCvW + j] = *reinterpret_cast<const uint32_t*>(&pk);
          }
        // The box can go back to the producer once its values have been CONSUMED (not merely requested: an arrive right
        // after the loads let the next TMA overwrite the slot under loads still in flight - wrong gradients at batch 256).
        // Rows beyond the converted window are only halo padding, so every loaded register that matters is used above.
        __syncwarp();
        if (lane == 0) mbar_arrive(&rawp_empty[rp.slot]);
#pragma unroll
        for (int t = 0; t < taps_here; ++t) {
          mbar_wait(&p_empty[stage], sphase ^ 1);
          tc_fence_after();
          const uint32_t t_p = tmem_p + t_lane + stage * 2 * kPPart;
          // the 32 pixels shifted by this tap (x offset t - 1): pixel pair (px, px + 1) -> one column of packed hi halves
          // and one of packed lo halves, each a byte permute of the two converted registers
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t h8[8], l8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int px = 16 * j + 2 * c, idx = (px / kW) * kCvW + (px % kW) + (kPad ? t : 0);
              h8[c] = __byte_perm(cv[idx], cv[idx + 1], 0x5410);
              l8[c] = __byte_perm(cv[idx], cv[idx + 1], 0x7632);
            }
            tmem_st_32x8u(t_p + 8 * j, h8);
            tmem_st_32x8u(t_p + kPPart + 8 * j, l8);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[stage]);
          MFLOW_WST(7, i, t == 0 && quad == 0 && lane == 0);
          advance(1);
        }
      }
    } else {
      // TF32: set s takes the (block, tap) work items of parity s and owns stage s
      const uint32_t t_p = tmem_p + t_lane + set * 2 * kPPart;
      int pp = 0, item = 0;   // phase of this set's stage; running work item
      for (int i = 0; i < n_my; ++i, rp.next()) {
        bool loaded = false;
        float raw[kRawVec * 4];
#pragma unroll
        for (int t = 0; t < taps_here; ++t, ++item) {
          if ((item & 1) != set) continue;
          if (!loaded) {
            load_raw(raw);
            loaded = true;
          }
          mbar_wait(&p_empty[set], pp ^ 1);
          pp ^= 1;
          tc_fence_after();
          // the 32 pixels shifted by this tap (x offset t - 1; box column = w + 4), ReLU, hi / lo split -> TMEM
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float h8[8], l8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int px = 8 * j + c;
              float v = raw[(px / kW) * kBoxW + (px % kW) + (kPad ? t + 3 : 0)];   // rows of both images are consecutive in raw[]
              if (a.relu_in) v = fmaxf(v, 0.f);
              h8[c] = wg_trunc(v);
              l8[c] = wg_round(v - h8[c]);
            }
            tmem_st_32x8(t_p + 8 * j, h8);
            tmem_st_32x8(t_p + kPPart + 8 * j, l8);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[set]);
        }
        __syncwarp();
        if (lane == 0 && loaded) mbar_arrive(&rawp_empty[rp.slot]);   // this warp is done with the block's raw box
      }
    }
    // ===== epilogue: accumulators -> partial[ks][tap][ci][co]; the sets take alternate 32-column chunks =====
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int ci = p0 + ch;
    const int n_here = min(128, a.Cout - q0);
    const int n_chunks = (n_here + 31) / 32;
    const float unscale = inv_sx * inv_sg;   // exact: both are powers of two (1 on the TF32 path)
    const bool vec4 = (a.Cout & 3) == 0 && (reinterpret_cast<uintptr_t>(a.partial) & 15) == 0;   // rows of the partials are 16-byte aligned
    for (int t = 0; t < taps_here; ++t) {
      const int tap = kTaps == 9 ? dy * 3 + t : 0;
      float* dst = a.partial + (((int64_t)ks * a.taps + tap) * a.Cin + ci) * a.Cout + q0;
      for (int c = set; c < n_chunks; c += 2) {
        float v[32];
        tmem_ld_32x32(tmem + t_lane + t * a.n_tile + c * 32, v);
        if (ci < a.Cin) {
          if (kTaps == 1 && a.direct_w != nullptr) {
            // final layout [co][ci]: for a fixed output channel the warp's lanes are consecutive input channels
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < n_here) a.direct_w[(int64_t)(q0 + c * 32 + j) * a.Cin + ci] = F16 ? v[j] * unscale : v[j];
          } else if (vec4) {
            // handled below (all lanes of the warp take part)
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < n_here) dst[c * 32 + j] = F16 ? v[j] * unscale : v[j];
          }
        }
        if (vec4 && !(kTaps == 1 && a.direct_w != nullptr)) {
          // The warp holds a [32 input channels (lanes)] x [32 output channels] block whose rows are 4 * Cout bytes apart
          // in the partials: stored from the lanes directly, every store instruction touches 32 different lines (the
          // epilogue was ~50 us of address-divergent stores per CTA with 4-byte stores, ~13 us with 16-byte ones).
          // Through a padded shared-memory tile (the drained raw-box ring) eight lanes write one whole 128-byte row
          // segment, four rows per instruction.
          float* tile = reinterpret_cast<float*>(s_p) + (warp - 2) * (32 * 33);
#pragma unroll
          for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = F16 ? v[j] * unscale : v[j];
          __syncwarp();
          const int rr = lane >> 3, q4 = lane & 7;
          const int64_t row0 = ((int64_t)ks * a.taps + tap) * a.Cin + p0 + quad * 32;
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += 4) {
            const int row = r0 + rr;
            const float* src = tile + row * 33 + 4 * q4;
            const float4 val = make_float4(src[0], src[1], src[2], src[3]);
            if (p0 + quad * 32 + row < a.Cin && c * 32 + 4 * q4 < n_here)
              *reinterpret_cast<float4*>(a.partial + (row0 + row) * a.Cout + q0 + c * 32 + 4 * q4) = val;
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== Q split: the TMA-loaded fp32 tile becomes its TF32 hi part in place, the lo part goes next to it =====
    // These threads see every element of g_y exactly once, so the kernel-row-1 CTA of input-channel tile 0 also
    // accumulates the bias gradient sum_pixels g_y[co] (a thread's float4 positions are the same in every block).
    const int tid = threadIdx.x - 320;   // 0..127
    WgRing rq(a.n_q);
    const int n_vec = a.n_tile * (kWgBlockPx / 4);   // float4s of the tile
    const bool do_bias = a.bias_partial != nullptr && dy == (kTaps == 9 ? 1 : 0) && p0 == 0;
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // n_vec <= 1024 = 8 x 128
    int last_slot = 0;
    float sg = 1.f, inv_sg = 1.f;
    if (F16) pow2_scale_from_amax(a.g_amax, sg, inv_sg);
    for (int i = 0; i < n_my; ++i, rq.next()) {
      mbar_wait(&qraw_full[rq.slot], rq.phase);
      MFLOW_WST(2, i, tid == 0);
      float4* hi = reinterpret_cast<float4*>(s_q + rq.slot * 2 * kWgQBytes);
      float4* lo = reinterpret_cast<float4*>(s_q + rq.slot * 2 * kWgQBytes + kWgQBytes);
      uint8_t* f16_hi = s_q + rq.slot * 2 * kWgQBytes + kWgQHiOff;
      uint8_t* f16_lo = s_q + rq.slot * 2 * kWgQBytes + kWgQLoOff;
      // all loads first: the stores below go to the same shared memory, so the compiler keeps every later load behind
      // every earlier store - interleaved, the seven 16-byte chunks of a thread formed one serial latency chain
      float4 vq[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = tid + 128 * k;
        vq[k] = j < n_vec ? hi[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = tid + 128 * k;
        if (j < n_vec) {
          const float4 v = vq[k];
          bsum[k] += (v.x + v.y) + (v.z + v.w);
          if (F16) {
            // float4 j of the TMA-swizzled fp32 tile -> (row r, 4-pixel group c8 of the 32-pixel block) -> 8 bytes of the
            // fp16 K-major tile [n_tile][64 B] with the 64B swizzle (16-byte chunk index ^= (row >> 1) & 3)
            int r, c8;
            if (kW == 4) {   // two fp32 tiles [n_tile][16 px] (64B swizzle), one per image: block pixel = 16 * image + px
              const int img = j / (a.n_tile * 4), jj = j - img * (a.n_tile * 4);
              r = jj >> 2;
              c8 = 4 * img + ((jj & 3) ^ ((r >> 1) & 3));
            } else {         // one fp32 tile [n_tile][32 px] (128B swizzle)
              r = j >> 3;
              c8 = (j & 7) ^ (r & 7);
            }
            uint2 h, l;
            split_f16x2(v.x * sg, v.y * sg, h.x, l.x);
            split_f16x2(v.z * sg, v.w * sg, h.y, l.y);
            const int off = r * 64 + (((c8 >> 1) ^ ((r >> 1) & 3)) << 4) + ((c8 & 1) << 3);
            *reinterpret_cast<uint2*>(f16_hi + off) = h;
            *reinterpret_cast<uint2*>(f16_lo + off) = l;
          } else {
            float4 h, l;
            h.x = wg_trunc(v.x); h.y = wg_trunc(v.y); h.z = wg_trunc(v.z); h.w = wg_trunc(v.w);
            l.x = wg_round(v.x - h.x); l.y = wg_round(v.y - h.y); l.z = wg_round(v.z - h.z); l.w = wg_round(v.w - h.w);
            hi[j] = h;
            lo[j] = l;
          }
        }
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's shared-memory reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full[rq.slot]);
      MFLOW_WST(3, i, tid == 0);
      last_slot = rq.slot;
    }
    if (do_bias) {
      // per-thread sums -> per-channel sums in a fixed order, through the (drained) Q ring
      mbar_wait(&acc_full, 0);                       // all MMAs have read the ring
      float* scratch = reinterpret_cast<float*>(s_q);
      (void)last_slot;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = tid + 128 * k;
        if (j < n_vec) scratch[j] = bsum[k];
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");   // the four Q-split warps
      if (tid < a.n_tile && q0 + tid < a.Cout) {
        float t = 0.f;
        if (kW == 4) {   // two tiles [n_tile][4 float4]: image 0 then image 1
          for (int e = 0; e < 4; ++e) t += scratch[tid * 4 + e];
          for (int e = 0; e < 4; ++e) t += scratch[a.n_tile * 4 + tid * 4 + e];
        } else {         // one tile [n_tile][8 float4] (the 128B swizzle permutes within a row only)
          for (int e = 0; e < 8; ++e) t += scratch[tid * 8 + e];
        }
        a.bias_partial[(int64_t)ks * a.Cout + q0 + tid] = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, a.tmem_cols);
}

// out[co][ci][tap] = sum over ks of partial[ks][tap][ci][co].
// block = 32 outputs x 8 k-lanes: lane q of an output adds the splits q, q + 8, ... (short dependent chains, 128-byte
// coalesced reads for a fixed split), then the 8 lane sums are combined in a fixed shuffle tree: deterministic.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n_ks, int taps,
                                                           int Cin, int Cout, const float* __restrict__ bias_partial,
                                                           float* __restrict__ g_bias) {
  const int64_t n = (int64_t)taps * Cin * Cout;
  const int o_in_blk = threadIdx.x & 31, q = threadIdx.x >> 5;   // warp w = k-lane w: its 32 lanes are 32 consecutive outputs
  __shared__ float s_part[8][33];
  const int64_t n_items = n + (g_bias != nullptr ? Cout : 0);    // the bias gradient's Cout sums are appended to the work list
  for (int64_t base = (int64_t)blockIdx.x * 32; base < n_items; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + o_in_blk;
    float acc = 0.f;
    if (i < n) {
      for (int k = q; k < n_ks; k += 8) acc += partial[(int64_t)k * n + i];
    } else if (i < n_items) {
      const int co = (int)(i - n);
      for (int k = q; k < n_ks; k += 8) acc += bias_partial[(int64_t)k * Cout + co];
    }
    s_part[q][o_in_blk] = acc;
    __syncthreads();
    if (q == 0 && i < n_items) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += s_part[r][o_in_blk];
      if (i < n) {
        // i indexes [tap][ci][co] (coalesced reads); the write is the permuted position
        const int co = (int)(i % Cout);
        const int ci = (int)((i / Cout) % Cin);
        const int tap = (int)(i / ((int64_t)Cout * Cin));
        out[((int64_t)co * Cin + ci) * taps + tap] = t;
      } else {
        g_bias[i - n] = t;
      }
    }
    __syncthreads();
  }
}

static int wgrad_splits(const mflow_conv_desc* d) {
  const int n_dy = d->kernel == 3 ? 3 : 1;
  const int tiles = (int)((d->c_in + 127) / 128) * (int)((d->c_out + 127) / 128);
  const int64_t n_blocks = d->width == 4 ? (d->batch + 1) / 2 : d->batch * d->height * d->width / kWgBlockPx;
  int64_t ks = kNumSMs / (n_dy * tiles);
  if (ks > n_blocks) ks = n_blocks;
  if (ks < 1) ks = 1;
  return (int)ks;
}

}  // namespace mflow

using namespace mflow;

extern "C" MFLOW_API int64_t mflow_conv2d_wgrad_workspace_bytes(const mflow_conv_desc* d) {
  if (!d) return 0;
  const int64_t taps = d->kernel == 3 ? 9 : 1;
  return (int64_t)wgrad_splits(d) * (taps * d->c_in * d->c_out + d->c_out) * (int64_t)sizeof(float);
}

#ifdef MFLOW_CONV_ABLATE
extern "C" MFLOW_API int mflow_debug_wgrad_steps(long long* out) {
  return cudaMemcpyFromSymbol(out, mflow::g_wg_steps, sizeof(long long) * 8 * 64) == cudaSuccess ? MFLOW_OK : MFLOW_E_LAUNCH;
}
#endif

template <bool F16>
static int launch_wgrad(const mflow_conv_desc* d, const float* x, const float* g_y, const float* x_amax, const float* g_amax,
                        float* g_w, float* g_bias, void* workspace, void* stream) {
  MFLOW_REQUIRE(d && x && g_y && g_w && workspace, "null pointer");
  MFLOW_REQUIRE(d->kernel == 3 || d->kernel == 1, "kernel size must be 1 or 3");
  MFLOW_REQUIRE(d->width == 32 || d->width == 16 || d->width == 8 || d->width == 4, "width must be 4, 8, 16 or 32");
  MFLOW_REQUIRE(d->height == d->width, "square feature maps only");
  MFLOW_REQUIRE(d->c_in >= 1 && d->c_out >= 1 && d->batch >= 1, "bad extents");
  MFLOW_REQUIRE(d->kernel == 1 || d->c_out <= 112, "3x3 weight gradients support at most 112 output channels");
  const int W = (int)d->width, H = (int)d->height, B = (int)d->batch;
  WgradArgs a{};
  a.partial = (float*)workspace;
  a.B = B; a.Cin = (int)d->c_in; a.Cout = (int)d->c_out; a.H = H; a.W = W;
  a.taps = d->kernel == 3 ? 9 : 1;
  a.relu_in = d->relu_in;
  const int imgs = W == 4 ? 2 : 1;            // 4x4 maps: a 32-pixel block is two whole images
  a.rows = kWgBlockPx / W / imgs;             // rows per image in a block
  a.box_w = a.taps == 9 ? W + 8 : W;
  a.raw_bytes = imgs * 128 * a.rows * a.box_w * 4;
  // Raw-box ring depth.  A transform set waits on rawp_full only for the blocks it owns, i.e. (fp16 path, and 1x1 on
  // either path) for every other block.  mbarrier waits are by phase PARITY: a waiter that sits out one phase of a slot
  // cannot tell "two completions ago" from "not yet".  3x3 / fp16: before a set reaches block i it has written block
  // i - 2's items into the five-stage P ring, which needs the MMAs of block i - 3's first item, hence block i - 3's box
  // (the previous user of the slot, owned by the OTHER set) has landed - three slots are safe.  1x1: one item per block,
  // the stage ring only reaches back to block i - 5, so with three slots a set could pass the wait for block i on the
  // stale parity while block i - 3's box was still in flight, release the slot early and let the producer arm the barrier
  // twice in one phase (seen as rare hangs / launch failures of the 1x1 weight gradients at 10^5 .. 10^6 rows).  With an
  // EVEN depth a slot always belongs to the same set, which then sees every one of its phases: 1x1 uses four slots
  // (16 KB boxes: the Q ring keeps its five slots).
  a.n_raw = a.raw_bytes > 32 * 1024 ? 2 : (a.taps == 1 ? 4 : 3);
#ifdef MFLOW_WGRAD_FORCE_RAW3   // dev-only: the pre-fix ring depth, to reproduce the hazard (tools/wgrad_repeat.py)
  a.n_raw = a.raw_bytes > 32 * 1024 ? 2 : 3;
#endif
  a.blocks_per_img = W == 4 ? 1 : H * W / kWgBlockPx;
  a.n_blocks = W == 4 ? (B + 1) / 2 : B * a.blocks_per_img;
  a.n_dy = a.taps == 9 ? 3 : 1;
  const int q_tiles = (a.Cout + 127) / 128, p_tiles = (a.Cin + 127) / 128;
  // one n_tile for the whole launch: a multiple of 16 covering the widest q tile
  { const int widest = a.Cout < 128 ? a.Cout : 128; a.n_tile = (widest + 15) / 16 * 16; }
  a.tmem_cols = 512;
  a.x_amax = x_amax; a.g_amax = g_amax;
  MFLOW_REQUIRE((a.taps == 9 ? 3 : 1) * a.n_tile + (F16 ? 5 * kWgPCols : 4 * kWgPCols) <= 512, "accumulators do not fit in tensor memory");
  const int n_ks = wgrad_splits(d);
  a.bias_partial = g_bias ? a.partial + (int64_t)n_ks * a.taps * a.Cin * a.Cout : nullptr;
  const bool direct = a.taps == 1 && n_ks == 1;   // nothing to reduce: the kernel writes g_w / g_bias in their final layout
  if (direct) { a.direct_w = g_w; a.bias_partial = g_bias; }

  CUtensorMap mx, mg;
  if (a.taps == 1) {  // activations of a 1x1 convolution as [B, C, 1, H*W]: box {32 / imgs contiguous pixels, 1, 128 channels, imgs}
    cuuint64_t dims[4] = {(cuuint64_t)H * W, 1, (cuuint64_t)a.Cin, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)a.Cin * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)(kWgBlockPx / imgs), 1u, 128u, (cuuint32_t)imgs};
    if (!encode(&mx, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return MFLOW_E_BADARG;
  } else {  // activations, NCHW: box {box_w, rows, 128 channels, imgs}; lands as [image][channel][row][w]
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)a.Cin, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)a.Cin * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)a.box_w, (cuuint32_t)a.rows, 128u, (cuuint32_t)imgs};
    if (!encode(&mx, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return MFLOW_E_BADARG;
  }
  {  // output gradient as [B][Cout][H*W]: box {32 pixels, n_tile channels, 1 image}, 128-byte swizzled K-major rows
    cuuint64_t dims[3] = {(cuuint64_t)H * W, (cuuint64_t)a.Cout, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)H * W * 4, (cuuint64_t)a.Cout * H * W * 4};
    cuuint32_t box[3] = {(cuuint32_t)(kWgBlockPx / imgs), (cuuint32_t)a.n_tile, (cuuint32_t)imgs};
    if (!encode(&mg, g_y, 3, dims, str, box, imgs == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)) return MFLOW_E_BADARG;
  }
  dim3 grid(n_ks, a.n_dy, q_tiles * p_tiles);
  a.n_q = (226 * 1024 - 1024 - a.n_raw * a.raw_bytes) / (2 * kWgQBytes);
  a.n_q = a.n_q > kWgQSlots ? kWgQSlots : (a.n_q < 3 ? 3 : a.n_q);
  const size_t smem = (size_t)a.n_q * 2 * kWgQBytes + (size_t)a.n_raw * a.raw_bytes + 1024;
  const cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  cudaGetDevice(&dev);
#define MFLOW_WGRAD_LAUNCH(WW, TT)                                                                                   \
  do {                                                                                                               \
    static bool attr_set[64] = {};   /* the opt-in shared-memory limit is a per-device function attribute */         \
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {                                                                    \
      cudaFuncSetAttribute(wgrad_split3_kernel<WW, TT, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); \
      attr_set[dev] = true;                                                                                          \
    }                                                                                                                \
    wgrad_split3_kernel<WW, TT, F16><<<grid, kWgThreads, smem, st>>>(mx, mg, a);                                     \
  } while (0)
  if (a.taps == 9) {
    if (W == 32) MFLOW_WGRAD_LAUNCH(32, 9); else if (W == 16) MFLOW_WGRAD_LAUNCH(16, 9); else if (W == 8) MFLOW_WGRAD_LAUNCH(8, 9);
    else MFLOW_WGRAD_LAUNCH(4, 9);
  } else {
    if (W == 32) MFLOW_WGRAD_LAUNCH(32, 1); else if (W == 16) MFLOW_WGRAD_LAUNCH(16, 1); else if (W == 8) MFLOW_WGRAD_LAUNCH(8, 1);
    else MFLOW_WGRAD_LAUNCH(4, 1);
  }
#undef MFLOW_WGRAD_LAUNCH
  if (direct) return check_launch(F16 ? "wgrad_split3_kernel<fp16>" : "wgrad_split3_kernel<tf32>");
  const int64_t n_out = (int64_t)a.taps * a.Cin * a.Cout;
  int64_t rblocks = (n_out + a.Cout + 31) / 32;
  const int rgrid = (int)(rblocks > 16 * kNumSMs ? 16 * kNumSMs : rblocks);
  wgrad_reduce_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(a.partial, g_w, n_ks, a.taps, a.Cin, a.Cout, a.bias_partial, g_bias);
  return check_launch(F16 ? "wgrad_split3_kernel<fp16>" : "wgrad_split3_kernel<tf32>");
}

extern "C" MFLOW_API int mflow_conv2d_wgrad_tf32x3(const mflow_conv_desc* d, const float* x, const float* g_y, float* g_w,
                                                   float* g_bias, void* workspace, void* stream) {
  return launch_wgrad<false>(d, x, g_y, nullptr, nullptr, g_w, g_bias, workspace, stream);
}

extern "C" MFLOW_API int mflow_conv2d_wgrad_f16x3(const mflow_conv_desc* d, const float* x, const float* g_y, const float* x_amax,
                                                  const float* g_amax, float* g_w, float* g_bias, void* workspace, void* stream) {
  return launch_wgrad<true>(d, x, g_y, x_amax, g_amax, g_w, g_bias, workspace, stream);
}
