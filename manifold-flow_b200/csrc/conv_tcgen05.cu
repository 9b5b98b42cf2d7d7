// Conditioner convolutions on the 5th-generation tensor cores (kernel (b) of the north star).
//
//   y[b, co, h, w] = bias[co] + sum_{tap, ci} relu?(x)[b, ci, h + dy, w + dx] * W[co, ci, tap]   (+ residual)
//
// for the 3x3 (pad 1) and 1x1 convolutions of ConvResidualNet (reference: manifold_flow/nn/resnet.py:75-142),
// as an implicit GEMM per 128-pixel tile:  D[128 pixels, N out channels] += A_tap[128, 32 ci] * B_tap[N, 32 ci]^T.
//
// * Activations stay in the reference's NCHW fp32 layout: no im2col, no layout conversion and no channel
//   padding in HBM.  ONE TMA box {W + 8, rows + 2, 32 channels, images} per channel block lands in shared
//   memory as [channel][pixel + halo]; all nine taps of the stencil are cut out of it.  (The innermost TMA
//   coordinate must be a multiple of 16 bytes, so the halo box starts 4 pixels to the left.)  The zero
//   padding of the convolution and of channels 100..127 is TMA's out-of-bounds fill.
// * fp32 parity on the tensor cores by an error-compensated operand split, a = a_hi + a_lo, w = w_hi + w_lo,
//   D += A_hi B_hi + A_lo B_hi + A_hi B_lo (fp32 accumulation in tensor memory), in one of two operand formats
//   (template parameter F16):
//     F16 = true  (default path): fp16 halves, kind::f16, K = 16 per instruction - fp16 carries the same 11
//                 significant bits as TF32 at TWICE the tensor-core rate.  fp16 has a 5-bit exponent, so both operands
//                 are scaled by powers of two first: the weights per output channel on the host (w_unscale undoes it
//                 in the epilogue), the activations by a scale derived from their absolute maximum, which the
//                 producing kernel publishes (out_amax -> the next launch's in_amax) - gradients of 1e-7 and
//                 activations of 1e4 both land in [2^13, 2^14) at the top and keep 22 significant bits.
//     F16 = false: TF32 containers (fp32 range, no scaling), kind::tf32, K = 8 per instruction.
//   Weights are split once per optimizer step on the host side of the ABI.  Activations are handled by the CTA's
//   eight transform warps between TMA arrival and the MMA:
//   two threads per pixel (16 channels each) read their channels (conflict-free), apply the
//   pre-activation ReLU, split, and store A_hi / A_lo straight into TENSOR MEMORY (tcgen05.st, TMEM lane = pixel), from
//   where the MMA reads its A operand: the NCHW -> [pixel][channel] transposition costs no extra pass
//   and the activation operand never touches shared memory again.
// * Accumulator: 128 lanes x 128 columns of TMEM; epilogue tcgen05.ld 32x32b gives every thread one pixel
//   row, so for a fixed output channel a warp stores 32 consecutive pixels (coalesced NCHW writes) with
//   bias / residual / ReLU / ReLU-mask fused.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = transform, then
//   epilogue (a warp may only touch the TMEM lane quadrant warp_id % 4, so every quadrant has two warps:
//   they split the channels of a step and the column chunks of the epilogue).  Hand-offs through mbarrier rings; tcgen05.commit releases slots.  Two CTAs share an SM so
//   that one CTA's prologue / epilogue overlaps the other's main loop.
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>

#include "mflow_common.cuh"
#include "rqs_math.cuh"
#include "tc_ptx.cuh"
#include "tma_encode.cuh"

namespace mflow {

using namespace ptx;

constexpr int kTileM = 128;   // pixels per CTA tile
constexpr int kTileN = 128;   // accumulator columns (output channels per CTA tile, N extent <= 128)
constexpr int kBlockK = 32;   // input channels per pipeline step
constexpr int kStageBytes = kTileM * kBlockK * 4;  // 16 KB: one 32-channel chunk of the output / aux tile (fp32) in the staging area
constexpr int kConvThreads = 320;   // 2 control warps + 8 transform / epilogue warps
constexpr int kXfWarps = 8;
constexpr int kMaxStagesA = 4;    // activation operand stages in TMEM: 2 (TF32: 64 columns each) or 4 (fp16: 32 columns each)
constexpr int kTmemCols = 256;    // 128 accumulator columns + stages x (A_hi + A_lo): 2 x (32 + 32) columns (TF32) or 4 x (16 + 16) (fp16 pairs)
constexpr int kMaxStagesB = 8;
constexpr int kMaxRaw = 2;

struct ConvArgs {
  float* y;
  const float* bias;      // [Cout] or null
  const float* residual;  // same shape as y or null
  const float* gate;      // same shape as y or null: result is zeroed where gate <= 0 (ReLU backward mask)
  int B, Cin, Cout, H, W;
  int taps;               // 9 (3x3, pad 1) or 1
  int relu_in;            // apply relu to the input on the fly
  int relu_out;           // apply relu to the output
  int rows_per_tile;      // image rows per 128-pixel tile (box H extent without halo)
  int imgs_per_tile;      // images per tile (box B extent)
  int k_blocks;           // ceil(Cin / 32)
  int box_w, box_h;       // raw activation box extents (W + 8, rows + 2 with a halo for 3x3; W, rows for 1x1)
  int raw_bytes;          // ring pitch of the raw boxes (box bytes rounded up to 1024)
  int raw_tx;             // true bytes of one raw box (all 32 channels, all images of the tile)
  int n_tile;             // MMA N extent: 112 when Cout <= 112 (13% less weight traffic and MMA time), else 128
  int n_raw, n_b;         // ring depths: raw boxes, weight tiles
  int w_tile_bytes;       // one weight operand tile (hi or lo): w_rows x 128 bytes, w_rows = C_out of the tile rounded up to 8
  int tail_k8;            // the last channel block has <= 8 channels: its weight tiles travel as 32-byte rows
  int flat;               // whole-image tiles of 8x8 / 4x4 maps: raw boxes without halo, [image][channel][H*W] (see launch_conv)
  int ring_bytes;         // weight ring (also the epilogue's staging area): max(n_b x 2 x w_tile_bytes, chunks x 16 KB)
  int aux_mode;           // 0 none, 1 residual add, 2 ReLU mask (gate)
  int aux_chunks;         // 32-channel boxes of the aux tile to fetch (0 when aux_mode == 0)
  int w_tail_bytes;       // one weight operand tile of the short last channel block (32-byte rows)
  // fp16 path: operand scaling (see the header comment); all null / unused on the TF32 path
  const float* in_amax;   // device scalar: absolute maximum of x (null = no scaling: the caller vouches for the range)
  const float* w_unscale; // [n_pad] 2^-s_n undoing the per-output-channel weight scale
  float* out_amax;        // device scalar (zero-initialised by the caller) receiving max |y| of this launch, or null
  // Spline epilogue (template parameter SPL != 0): this launch is the conditioner's FINAL 1x1 convolution of an image
  // coupling layer (C_out = C_t * 32: the P = 32 spline parameters of transformed channel c are the accumulator columns
  // 32 c .. 32 c + 31 of the pixel's TMEM lane) and the coupling layer's spline is evaluated right there: the
  // [B, C_t * 32, H, W] parameter tensor never exists (coupling.py:108-110; evaluation only).
  const float* sp_x;      // activation tensor of the coupling layer, pointing at its first transformed channel
  float* sp_y;            // output tensor, same indexing
  float* sp_lad;          // [B][sp_slots] partial log-dets: one per 16 pixels and transformed channel
  int64_t sp_so, sp_sc;   // sample / transformed-channel strides of x and y (floats)
  int64_t sp_id_off, sp_id_sc;   // identity channels (copied through), relative to sp_x
  int sp_n_id, sp_slots;
  RqsConsts sp_c;
#ifdef MFLOW_CONV_ABLATE
  int ablate;             // dev-only timing experiments (wrong results): 1 no B_lo product, 2 no gather loads, 4 MMA N = 64, 16 no TMEM stores
#endif
};
#ifdef MFLOW_CONV_ABLATE
#define MFLOW_ABL(bit) (a.ablate & (bit))
// dev-only per-CTA timeline: [cta][0] clock at entry, [1] after the set-up barrier, [2] first MMA issue, [3] accumulator
// complete (seen by an epilogue warp), [4] epilogue done, [5] exit, [6] globaltimer at entry, [7] globaltimer at exit, [8] SM id
__device__ long long g_conv_tl[4096][9];
// ... and per-step stamps of CTA 600: [0] transform: before the stage wait, [1] stage free, [2] stage filled + arrived,
// [3] MMA warp: activation stage full, [4] weights full, [5] MMAs issued + committed, [6] producer: weight slot free (TMA issue)
__device__ long long g_conv_steps[7][64];
#define MFLOW_ST(row, idx, cond) do { if (blockIdx.x == 600 && (cond) && (idx) < 64) g_conv_steps[row][idx] = clock64(); } while (0)
#define MFLOW_TL(slot, cond) do { if ((cond) && blockIdx.x < 4096 && blockIdx.y == 0) g_conv_tl[blockIdx.x][slot] = clock64(); } while (0)
#else
#define MFLOW_ST(row, idx, cond) do { } while (0)
#define MFLOW_TL(slot, cond) do { } while (0)
#define MFLOW_ABL(bit) false
#endif

__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// round to nearest TF32 (ties away): the tensor core truncates its fp32 inputs, so the low part of the
// split is pre-rounded to keep the truncation from biasing the sum
__device__ __forceinline__ float tf32_round(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

// ring cursor without divisions: slot index + phase parity
struct Ring {
  int slot, phase, n;
  __device__ __forceinline__ Ring(int depth) : slot(0), phase(0), n(depth) {}
  __device__ __forceinline__ void next() { if (++slot == n) { slot = 0; phase ^= 1; } }
};

// CS: compile-time number of floats between two channels of the raw box (box_w * box_h; 0 = read it from the
// arguments): the gather's 32 shared-memory loads then use immediate offsets instead of one address multiply each.
template <bool F16, int CS, int SPL = 0>   // SPL: 0 plain convolution, 1 / 2 spline epilogue forward / inverse
__global__ void __launch_bounds__(kConvThreads, 2)
conv_split3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                   const __grid_constant__ CUtensorMap map_wlo, const __grid_constant__ CUtensorMap map_whi8,
                   const __grid_constant__ CUtensorMap map_wlo8, const __grid_constant__ CUtensorMap map_aux,
                   const __grid_constant__ CUtensorMap map_y, const ConvArgs a) {
  // The kernel has NO static shared memory, so the dynamic array starts at the CTA's (1 KB aligned) base and the
  // 128-byte-swizzled weight tiles need no alignment slack: two raw boxes + two weight stages + the barriers
  // are 114.9 KB, which still lets two CTAs share an SM.  (Pointers are derived from the __shared__ array by
  // pointer arithmetic: a cast through uintptr_t makes the compiler forget the address space and emit generic
  // loads for the activation gather.)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_b = smem;                               // n_b x {B_hi, B_lo} weight tiles
  uint8_t* s_raw = s_b + a.ring_bytes;               // n_raw x raw box
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_raw + a.n_raw * a.raw_bytes);
  uint64_t* raw_full = bars;                         // [kMaxRaw]
  uint64_t* raw_empty = raw_full + kMaxRaw;          // [kMaxRaw]
  uint64_t* b_full = raw_empty + kMaxRaw;            // [kMaxStagesB]
  uint64_t* b_empty = b_full + kMaxStagesB;          // [kMaxStagesB]
  uint64_t* a_full = b_empty + kMaxStagesB;          // [kMaxStagesA]
  uint64_t* a_empty = a_full + kMaxStagesA;          // [kMaxStagesA]
  uint64_t* acc_full_p = a_empty + kMaxStagesA;
  uint64_t* aux_full_p = acc_full_p + 1;
  uint32_t* s_tmem_p = reinterpret_cast<uint32_t*>(aux_full_p + 1);
#define acc_full (*acc_full_p)
#define aux_full (*aux_full_p)
#define s_tmem (*s_tmem_p)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  MFLOW_TL(0, threadIdx.x == 0);
#ifdef MFLOW_CONV_ABLATE
  if (threadIdx.x == 0 && blockIdx.x < 4096 && blockIdx.y == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_conv_tl[blockIdx.x][6] = (long long)t;
    unsigned int sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_conv_tl[blockIdx.x][8] = sm;
  }
#endif
  constexpr int kAPart = F16 ? kBlockK / 2 : kBlockK;   // TMEM columns of one part (hi or lo) of an activation stage
  // fp16 stages are half as wide, so four fit: each half of the transform warps owns two and can run one step ahead of
  // the MMAs that read its previous stage (with one stage per half it idled for the whole MMA + commit latency: 20% of
  // all warp samples in ncu were that wait)
  constexpr int kStagesA = F16 ? 4 : 2;
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * kTileN;
  const int tiles_per_img = a.imgs_per_tile > 1 ? 1 : a.H / a.rows_per_tile;
  const int b0 = a.imgs_per_tile > 1 ? tile * a.imgs_per_tile : tile / tiles_per_img;
  const int h0 = a.imgs_per_tile > 1 ? 0 : (tile % tiles_per_img) * a.rows_per_tile;
  const int pad = a.taps == 9 ? 1 : 0;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_whi);
    prefetch_tensormap(&map_wlo);
    for (int i = 0; i < kMaxRaw; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], kXfWarps); }
    for (int i = 0; i < kMaxStagesB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < kMaxStagesA; ++i) { mbar_init(&a_full[i], kXfWarps / 2); mbar_init(&a_empty[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&aux_full, 1);
    fence_barrier_init();
    // First loads before the TMEM allocation and the CTA-wide barrier (the barriers they signal were initialised by
    // this very thread): the first raw box(es) and as many weight tiles of channel block 0 as the ring holds.
    // The producer loop below skips exactly these.
    {
      const int n_raw_pre = (a.n_raw >= 2 && a.k_blocks >= 2) ? 2 : 1;
      for (int kb = 0; kb < n_raw_pre; ++kb) {
        mbar_expect_tx(&raw_full[kb], a.raw_tx);
        tma_load_4d(s_raw + kb * a.raw_bytes, &map_x, &raw_full[kb], (pad && !a.flat) ? -4 : 0, a.flat ? 0 : h0 - pad, kb * kBlockK, b0);
      }
      const int n_w_pre = min(a.n_b, a.taps);
      const bool t8 = a.tail_k8 && a.k_blocks == 1;
      for (int tap = 0; tap < n_w_pre; ++tap) {
        uint8_t* dst = s_b + tap * 2 * a.w_tile_bytes;
        mbar_expect_tx(&b_full[tap], t8 ? 2 * a.w_tail_bytes : 2 * a.w_tile_bytes);
        tma_load_3d(dst, t8 ? &map_whi8 : &map_whi, &b_full[tap], 0, n0, tap);
        tma_load_3d(dst + a.w_tile_bytes, t8 ? &map_wlo8 : &map_wlo, &b_full[tap], 0, n0, tap);
      }
    }
  }
  if (warp == 1) tmem_alloc(&s_tmem, kTmemCols);  // [0,128) accumulator, then kStagesA x {A_hi, A_lo} x 32 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  MFLOW_TL(1, threadIdx.x == 0);

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      Ring rb(a.n_raw), sb(a.n_b);
      // With two raw slots the boxes run one channel block ahead of the weight tiles: the box of block kb + 1 is
      // requested before the nine weight tiles of block kb (its slot was freed by block kb - 1).  Requested after
      // them it arrived ~1500 cycles late at every block boundary (step trace), in both co-resident CTAs at once.
      const bool ahead = a.n_raw >= 2;
      const int n_raw_pre = (a.n_raw >= 2 && a.k_blocks >= 2) ? 2 : 1, n_w_pre = min(a.n_b, a.taps);   // issued before the CTA barrier
      auto load_raw = [&](int kb) {
        if (kb >= n_raw_pre) {
          mbar_wait(&raw_empty[rb.slot], rb.phase ^ 1);
          mbar_expect_tx(&raw_full[rb.slot], a.raw_tx);
          tma_load_4d(s_raw + rb.slot * a.raw_bytes, &map_x, &raw_full[rb.slot], (pad && !a.flat) ? -4 : 0, a.flat ? 0 : h0 - pad, kb * kBlockK, b0);
        }
        rb.next();
      };
      if (ahead) load_raw(0);
      // The producer issues in program order, so a wait blocks everything behind it: the box of block kb + 1 (whose
      // slot frees up only when block kb - 1 is fully transformed) goes out AFTER the first weight tiles of block kb,
      // which would otherwise arrive ~4500 cycles late at every block boundary (step trace).
      // (1x1 convolutions have one weight tile per block and live off the boxes: there the box goes first.)
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        if (!ahead) load_raw(kb);
        else if (a.taps == 1 && kb + 1 < a.k_blocks) load_raw(kb + 1);
        for (int tap = 0; tap < a.taps; ++tap, sb.next()) {
          if (ahead && a.taps > 1 && tap == 2 && kb + 1 < a.k_blocks) load_raw(kb + 1);
          if (kb == 0 && tap < n_w_pre) continue;
          mbar_wait(&b_empty[sb.slot], sb.phase ^ 1);
          MFLOW_ST(6, kb * a.taps + tap, true);
          if (a.tail_k8 && kb == a.k_blocks - 1) {
            // short last channel block (at most 8 TF32 / 16 fp16 channels): 32-byte rows (32B swizzle) instead of full
            // ones - the kernel is bound by the L2 -> shared weight traffic, and this block would be mostly zero padding
            mbar_expect_tx(&b_full[sb.slot], 2 * a.w_tail_bytes);
            tma_load_3d(s_b + sb.slot * 2 * a.w_tile_bytes, &map_whi8, &b_full[sb.slot], kb * kBlockK, n0, tap);
            tma_load_3d(s_b + sb.slot * 2 * a.w_tile_bytes + a.w_tile_bytes, &map_wlo8, &b_full[sb.slot], kb * kBlockK, n0, tap);
          } else {
            mbar_expect_tx(&b_full[sb.slot], 2 * a.w_tile_bytes);
            tma_load_3d(s_b + sb.slot * 2 * a.w_tile_bytes, &map_whi, &b_full[sb.slot], kb * kBlockK, n0, tap);
            tma_load_3d(s_b + sb.slot * 2 * a.w_tile_bytes + a.w_tile_bytes, &map_wlo, &b_full[sb.slot], kb * kBlockK, n0, tap);
          }
        }
      }
      // Epilogue operand (residual or ReLU mask, shaped like the output): once the last MMA has drained the
      // weight ring, pull the tile's {W, rows, 32 channels, images} boxes into that memory with TMA - bulk,
      // asynchronous, 512-byte bursts - instead of 100 dependent 128-byte loads per warp.
      if (a.aux_chunks) {
        mbar_wait(&acc_full, 0);
        mbar_expect_tx(&aux_full, a.aux_chunks * kStageBytes);
        for (int c = 0; c < a.aux_chunks; ++c)
          tma_load_4d(s_b + c * kStageBytes, &map_aux, &aux_full, 0, h0, n0 + c * 32, b0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // elect.sync (not lane == 0): the compiler then issues the tcgen05 instructions back to back instead of
    // wrapping each one in a warp-election loop (the MMAs here are only 56 cycles long)
    if (elect_one()) {
      const int n_mma = MFLOW_ABL(4) ? 64 : a.n_tile;
      const uint32_t idesc = F16 ? idesc_f16(kTileM, n_mma, 0, 0) : idesc_tf32(kTileM, n_mma, /*a K-major*/ 0, /*b K-major*/ 0);
      constexpr int kStepK = F16 ? 16 : 8;   // channels per MMA instruction: 32 bytes of a weight row, 8 TMEM columns of A
      Ring sb(a.n_b), sa(kStagesA);
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        const int valid = min(kBlockK, a.Cin - kb * kBlockK);
        const int ksteps = (valid + kStepK - 1) / kStepK;
        for (int tap = 0; tap < a.taps; ++tap, sb.next(), sa.next()) {
          const bool first = kb == 0 && tap == 0;
          mbar_wait(&a_full[sa.slot], sa.phase);
          MFLOW_ST(3, kb * a.taps + tap, true);
          mbar_wait(&b_full[sb.slot], sb.phase);
          MFLOW_ST(4, kb * a.taps + tap, true);
          tc_fence_after();
          MFLOW_TL(2, first);
          const uint32_t ahi = tmem + kTileN + sa.slot * 2 * kAPart, alo = ahi + kAPart;
          const uint32_t bhi = smem_u32(s_b + sb.slot * 2 * a.w_tile_bytes);
          // K-major swizzled weight tiles (128-byte rows / 128B swizzle for TF32, 64-byte rows / 64B swizzle for fp16):
          // one MMA's K extent = 32 bytes (2 descriptor units) further along each row, 8 TMEM columns further in A
          const bool tail8 = a.tail_k8 && kb == a.k_blocks - 1;   // 32-byte rows, 32B swizzle: 8-row groups 256 bytes apart
          const uint32_t sbo = F16 ? 512 : 1024, lay = F16 ? 4 : 2;
          const uint64_t d_bhi = tail8 ? smem_desc(bhi, 16, 256, 6) : smem_desc(bhi, 16, sbo, lay);
          const uint64_t d_blo = tail8 ? smem_desc(bhi + a.w_tile_bytes, 16, 256, 6) : smem_desc(bhi + a.w_tile_bytes, 16, sbo, lay);
          auto mma3 = [&](int kk) {
            if (F16) {
              umma_f16_ts(tmem, ahi + kk * 8, d_bhi + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
              umma_f16_ts(tmem, alo + kk * 8, d_bhi + 2 * kk, idesc, 1);
              if (!MFLOW_ABL(1)) umma_f16_ts(tmem, ahi + kk * 8, d_blo + 2 * kk, idesc, 1);
            } else {
              umma_tf32_ts(tmem, ahi + kk * 8, d_bhi + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
              umma_tf32_ts(tmem, alo + kk * 8, d_bhi + 2 * kk, idesc, 1);
              umma_tf32_ts(tmem, ahi + kk * 8, d_blo + 2 * kk, idesc, 1);
            }
          };
          if (ksteps == kBlockK / kStepK) {
#pragma unroll
            for (int kk = 0; kk < kBlockK / kStepK; ++kk) mma3(kk);
          } else {
            for (int kk = 0; kk < ksteps; ++kk) mma3(kk);
          }
          umma_commit(&a_empty[sa.slot]);  // slots are free again once these MMAs have read them
          umma_commit(&b_empty[sb.slot]);
          MFLOW_ST(5, kb * a.taps + tap, true);
        }
      }
      umma_commit(&acc_full);
    }
  } else {
    // ===== activation transform: TMEM lane = pixel; the two warps of a lane quadrant alternate pipeline steps =====
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int pm = quad * 32 + lane;
    const int pw = pm % a.W, pg = pm / a.W;
    const int phl = pg % a.rows_per_tile, pbl = pg / a.rows_per_tile;
    const int chan_stride = CS ? CS : a.box_h * a.box_w;                    // floats between channels in the raw box
    const bool flat = CS == 0 && a.flat;   // flat boxes always take the run-time-stride instantiation: no predicate in the others
    const int raw_base = pbl * kBlockK * chan_stride + phl * a.box_w + pw;  // raw box layout [image][channel][row][w]
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    // The halves alternate pipeline steps (half h owns the TMEM stages h, h + 2, ...): two steps are in flight in
    // different warps, so one warp's gather / split / tcgen05.st latency hides behind the other's.
    int my_stage = 0;          // which of this half's stages the next step uses
    uint32_t my_phase = 0;     // bit i: phase of this half's stage i
    float sa = 1.f, inv_sa = 1.f;   // fp16 path: power-of-two scale of the activation operand and its inverse
    if (F16) pow2_scale_from_amax(a.in_amax, sa, inv_sa);
    Ring rb(a.n_raw);
    int step = 0;   // step parity
    for (int kb = 0; kb < a.k_blocks; ++kb, rb.next()) {
      const int nch = F16 ? (min(kBlockK, a.Cin - kb * kBlockK) + 15) & ~15
                          : (min(kBlockK, a.Cin - kb * kBlockK) + 7) & ~7;  // channels the MMAs of this block read
      mbar_wait(&raw_full[rb.slot], rb.phase);
      const float* raw = reinterpret_cast<const float*>(s_raw + rb.slot * a.raw_bytes) + raw_base;
      for (int tap = 0; tap < a.taps; ++tap, step ^= 1) {
        if (step != half) continue;
        // halo box: the tap's pixel sits (dy, dx + 3) further in the box.  Flat box (no halo): the neighbour pixel of the
        // same image, or zero when it lies outside the image (the convolution's padding)
        const int dy = tap / 3, dx = tap - 3 * dy;
        const float* src = raw + (!pad ? 0 : flat ? (dy - 1) * a.W + (dx - 1) : dy * a.box_w + dx + 3);
        const bool ok = !(pad && flat) || ((unsigned)(phl + dy - 1) < (unsigned)a.H && (unsigned)(pw + dx - 1) < (unsigned)a.W);
        const int stage = half + 2 * my_stage;
        const uint32_t t_a = t_lane + kTileN + stage * 2 * kAPart;
        MFLOW_ST(0, kb * a.taps + tap, quad == 0 && lane == 0);
        mbar_wait(&a_empty[stage], ((my_phase >> my_stage) & 1u) ^ 1u);
        MFLOW_ST(1, kb * a.taps + tap, quad == 0 && lane == 0);
        my_phase ^= 1u << my_stage;
        tc_fence_after();
        // gather the pixel's channels of the tap, optional ReLU, hi / lo split, and store both parts into the
        // pixel's TMEM lane (A_hi in the first kAPart columns of the stage, A_lo in the next kAPart)
        if (F16) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {     // 16 channels = 8 columns of packed fp16 pairs per part
            const int c0 = g * 16;
            if (c0 < nch) {
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                float v0 = MFLOW_ABL(2) ? 1.f : (ok ? src[(c0 + 2 * c) * chan_stride] : 0.f), v1 = MFLOW_ABL(2) ? 2.f : (ok ? src[(c0 + 2 * c + 1) * chan_stride] : 0.f);
                if (a.relu_in) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                split_f16x2(v0 * sa, v1 * sa, hi[c], lo[c]);
              }
              tmem_st_32x8u(t_a + g * 8, hi);
              tmem_st_32x8u(t_a + kAPart + g * 8, lo);
            }
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int c0 = g * 8;
            if (c0 < nch) {
              float hi[8], lo[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                float v = ok ? src[(c0 + c) * chan_stride] : 0.f;
                if (a.relu_in) v = fmaxf(v, 0.f);
                hi[c] = tf32_trunc(v);
                lo[c] = tf32_round(v - hi[c]);
              }
              tmem_st_32x8(t_a + c0, hi);
              tmem_st_32x8(t_a + kAPart + c0, lo);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[stage]);
        MFLOW_ST(2, kb * a.taps + tap, quad == 0 && lane == 0);
        if (kStagesA > 2) my_stage ^= 1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[rb.slot]);   // this warp's reads of the box are program-ordered before
    }
    // ===== epilogue: TMEM -> registers -> staging tile -> TMA store; bias / residual / ReLU / mask fused =====
    // Per-output-channel constants of this half's two 32-column chunks, one channel per lane, fetched while the last MMAs
    // run and broadcast by shuffles below (as dependent __ldg's inside the loop they cost ~2500 cycles per chunk):
    // weight unscale factor x activation unscale (fp16 path; w_unscale is padded like the weights) and bias.
    float fac_l[2], bias_l[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int co = n0 + half * 32 + 64 * q + lane;
      fac_l[q] = F16 ? inv_sa * __ldg(a.w_unscale + co) : 1.f;
      bias_l[q] = (a.bias != nullptr && co < a.Cout) ? __ldg(a.bias + co) : 0.f;
    }
    if constexpr (SPL != 0) {
      // ===== spline epilogue: thread = pixel, one 32-column chunk = the 32 parameters of one transformed channel =====
      const int b = b0 + pbl;
      const bool valid = b < a.B;
      const int pix = (h0 + phl) * a.W + pw;                       // pixel within the image plane
      const int64_t base = (int64_t)b * a.sp_so + pix;
      if (blockIdx.y == 0 && valid)                                 // identity channels pass through (coupling.py:116-118)
        for (int k = half; k < a.sp_n_id; k += 2) {
          const int64_t off = base + a.sp_id_off + k * a.sp_id_sc;
          a.sp_y[off] = __ldg(a.sp_x + off);
        }
      mbar_wait(&acc_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 32; c0 < kTileN; c0 += 64) {
        if (n0 + c0 >= a.Cout) break;
        float v[32];
        tmem_ld_32x32(t_lane + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float bj = __shfl_sync(0xffffffffu, c0 >= 64 ? bias_l[1] : bias_l[0], j);
          v[j] = F16 ? fmaf(v[j], __shfl_sync(0xffffffffu, c0 >= 64 ? fac_l[1] : fac_l[0], j), bj) : v[j] + bj;
        }
        const int ch = (n0 + c0) >> 5;
        float lad = 0.f;
        if (valid) {
          const int64_t off = base + ch * a.sp_sc;
          float out;
          int bin;
          rqs_eval<11, SPL == 2>(__ldg(a.sp_x + off), v, a.sp_c, out, lad, bin);
          a.sp_y[off] = out;
        }
        // 16 consecutive pixels belong to one image for every map size: one partial per half warp, summed in a fixed order
        lad += __shfl_xor_sync(0xffffffffu, lad, 1);
        lad += __shfl_xor_sync(0xffffffffu, lad, 2);
        lad += __shfl_xor_sync(0xffffffffu, lad, 4);
        lad += __shfl_xor_sync(0xffffffffu, lad, 8);
        if ((lane & 15) == 0 && valid) a.sp_lad[(int64_t)b * a.sp_slots + (pix >> 4) * (a.Cout >> 5) + ch] = lad;
      }
    } else {
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    MFLOW_TL(3, warp == 2 && lane == 0);
    // aux tile in shared memory: layout [chunk][image][channel 32][row][w] (the TMA boxes), thread = pixel
    const int tile_px = a.rows_per_tile * a.W;
    float* stage = reinterpret_cast<float*>(s_b) + pbl * 32 * tile_px + phl * a.W + pw;
    if (a.aux_chunks) mbar_wait(&aux_full, 0);
    const bool issuer = (warp == 2 + 4 * half) && lane == 0;
    float out_max = 0.f;
#pragma unroll 1
    for (int c0 = half * 32; c0 < kTileN; c0 += 64) {   // 32-column chunks alternate between the two halves
      if (n0 + c0 >= a.Cout) break;
      float v[32];
      tmem_ld_32x32(t_lane + c0, v);
      // the finished chunk overwrites its own slot of the staging area (same [image][channel][row][w] layout
      // as the aux boxes) and leaves as ONE asynchronous TMA tile store; channels >= Cout are clipped by TMA
      float* auxc = stage + (c0 / 32) * (kStageBytes / 4);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int co = n0 + c0 + j;
        // fp16 path: undo the operand scales (exact powers of two)
        const float bj = __shfl_sync(0xffffffffu, c0 >= 64 ? bias_l[1] : bias_l[0], j);
        float r = F16 ? fmaf(v[j], __shfl_sync(0xffffffffu, c0 >= 64 ? fac_l[1] : fac_l[0], j), bj) : v[j] + bj;
        if (a.aux_mode == 1) r += auxc[j * tile_px];                        // residual add
        if (a.relu_out) r = fmaxf(r, 0.f);
        if (a.aux_mode == 2) r = auxc[j * tile_px] > 0.f ? r : 0.f;         // ReLU mask of the backward pass
        auxc[j * tile_px] = r;
        if (co < a.Cout) out_max = fmaxf(out_max, fabsf(r));   // columns >= Cout hold whatever the MMA read past the weight rows
      }
      fence_proxy_async();
      if (half) named_bar_sync<2>(128); else named_bar_sync<1>(128);
      if (issuer) {
        tma_store_4d(&map_y, reinterpret_cast<const float*>(s_b) + (c0 / 32) * (kStageBytes / 4), 0, h0, n0 + c0, b0);
        tma_store_commit();
      }
    }
    if (a.out_amax != nullptr) {
      // absolute maximum of this launch's output -> the operand scale of the next (fp16-split) consumer.  Pixels beyond
      // the batch (partial last tile) computed on zero-filled inputs and are clipped by the store: bias-sized at most.
      const bool in_range = a.imgs_per_tile > 1 ? (b0 + pbl < a.B) : true;
      const uint32_t m = __reduce_max_sync(0xffffffffu, in_range ? __float_as_uint(out_max) : 0u);
      if (lane == 0) amax_publish(a.out_amax, __uint_as_float(m));
    }
    if (issuer) tma_store_wait_read<0>();  // shared memory must stay valid until the stores have read it
    MFLOW_TL(4, warp == 2 && lane == 0);
    }   // SPL == 0
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
  MFLOW_TL(5, threadIdx.x == 32);
#ifdef MFLOW_CONV_ABLATE
  if (threadIdx.x == 32 && blockIdx.x < 4096 && blockIdx.y == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_conv_tl[blockIdx.x][7] = (long long)t;
  }
#endif
}
#undef acc_full
#undef aux_full
#undef s_tmem
constexpr int kConvBarrierBytes = (2 * kMaxRaw + 2 * kMaxStagesB + 2 * kMaxStagesA + 2) * 8 + 16;

// ---- host side --------------------------------------------------------------------------------
}  // namespace mflow

using namespace mflow;

#ifdef MFLOW_CONV_ABLATE
extern "C" MFLOW_API int mflow_debug_conv_timeline(long long* out, int n_ctas) {
  return cudaMemcpyFromSymbol(out, mflow::g_conv_tl, sizeof(long long) * 9 * (n_ctas < 4096 ? n_ctas : 4096)) == cudaSuccess ? MFLOW_OK : MFLOW_E_LAUNCH;
}
#endif

#ifdef MFLOW_CONV_ABLATE
extern "C" MFLOW_API int mflow_debug_conv_steps(long long* out) {
  return cudaMemcpyFromSymbol(out, mflow::g_conv_steps, sizeof(long long) * 7 * 64) == cudaSuccess ? MFLOW_OK : MFLOW_E_LAUNCH;
}
#endif

// shared host side of both operand formats
template <bool F16>
static int launch_conv(const mflow_conv_desc* d, const float* x, const void* w_hi, const void* w_lo, const float* w_unscale,
                       const float* in_amax, const float* bias, const float* residual, const float* gate, float* y, float* out_amax,
                       void* stream, const mflow_rqs_desc* sd = nullptr, const float* sp_x = nullptr, float* sp_y = nullptr,
                       float* sp_lad = nullptr) {
  MFLOW_REQUIRE(d && x && w_hi && w_lo && (y || sd), "null pointer");
  MFLOW_REQUIRE(!F16 || w_unscale, "the fp16 path needs the per-channel weight unscale vector");
  MFLOW_REQUIRE(!(residual && gate), "residual and gate cannot be combined in one call");
  MFLOW_REQUIRE(d->kernel == 3 || d->kernel == 1, "kernel size must be 1 or 3");
  MFLOW_REQUIRE(d->width == 32 || d->width == 16 || d->width == 8 || d->width == 4, "width must be 4, 8, 16 or 32");
  MFLOW_REQUIRE(d->height == d->width, "square feature maps only");
  MFLOW_REQUIRE(d->c_in >= 1 && d->c_out >= 1 && d->batch >= 0, "bad extents");
  MFLOW_REQUIRE(d->k_pad % kBlockK == 0 && d->k_pad >= d->c_in && d->n_pad % kTileN == 0 && d->n_pad >= d->c_out,
                "weights must be padded to [taps][n_pad %% 128 == 0][k_pad %% 32 == 0]");
  if (d->batch == 0) return MFLOW_OK;
  constexpr int wsz = F16 ? 2 : 4;           // bytes per weight element
  constexpr int kTailK = F16 ? 16 : 8;       // channels in a 32-byte weight row
  const int W = (int)d->width, H = (int)d->height, B = (int)d->batch;
  ConvArgs a{};
  a.y = y; a.bias = bias; a.residual = residual; a.gate = gate;
  a.in_amax = in_amax; a.w_unscale = w_unscale; a.out_amax = out_amax;
  if (sd != nullptr) {   // spline epilogue: the final 1x1 convolution of an image coupling layer, K = 11 (P = 32)
    MFLOW_REQUIRE(F16 && d->kernel == 1 && d->c_out % 32 == 0 && sd->num_bins == 11 && !residual && !gate && !out_amax && sp_x && sp_y && sp_lad,
                  "spline epilogue: fp16 format, 1x1 convolution with C_out = 32 C_t, 11 bins");
    MFLOW_REQUIRE(sd->n_chan * 32 == d->c_out && sd->n_outer == d->batch && sd->n_inner == d->height * d->width && sd->x_so == sd->y_so &&
                      sd->x_sc == sd->y_sc && sd->tail_bound > 0.f && sd->wh_divisor > 0.f,
                  "spline epilogue: descriptor does not match the convolution");
    MFLOW_REQUIRE((double)sd->min_bin_width * 11 <= 1.0, "Minimal bin width too large for the number of bins");
    MFLOW_REQUIRE((double)sd->min_bin_height * 11 <= 1.0, "Minimal bin height too large for the number of bins");
    a.sp_x = sp_x; a.sp_y = sp_y; a.sp_lad = sp_lad;
    a.sp_so = sd->x_so; a.sp_sc = sd->x_sc; a.sp_id_off = sd->id_off; a.sp_id_sc = sd->id_sc; a.sp_n_id = (int)sd->n_id_chan;
    a.sp_slots = (int)(d->height * d->width / 16 * (d->c_out / 32));
    RqsConsts& c = a.sp_c;
    const double T = (double)sd->tail_bound;
    c.tail = (float)T; c.lo = (float)(-T); c.span = (float)(T - (-T));
    c.min_w = sd->min_bin_width; c.min_h = sd->min_bin_height; c.min_d = sd->min_derivative;
    c.cfac_w = (float)(1.0 - (double)sd->min_bin_width * 11);
    c.cfac_h = (float)(1.0 - (double)sd->min_bin_height * 11);
    c.edge = (float)log(exp(1.0 - (double)sd->min_derivative) - 1.0);
    c.divisor = sd->wh_divisor; c.rcp_divisor = 1.0f / sd->wh_divisor; c.use_divisor = sd->wh_divisor != 1.0f;
  }
#ifdef MFLOW_CONV_ABLATE
  { const char* e = getenv("MFLOW_CONV_ABLATE"); a.ablate = e ? atoi(e) : 0; }
#endif
  a.B = B; a.Cin = (int)d->c_in; a.Cout = (int)d->c_out; a.H = H; a.W = W;
  a.taps = d->kernel == 3 ? 9 : 1;
  a.relu_in = d->relu_in; a.relu_out = d->relu_out;
  const int groups = kTileM / W;  // (image, row) groups per tile
  a.rows_per_tile = groups < H ? groups : H;
  a.imgs_per_tile = groups / a.rows_per_tile;
  a.k_blocks = (a.Cin + kBlockK - 1) / kBlockK;
  // 8x8 and 4x4 maps: a tile is a few WHOLE images.  Their halo boxes have 64- / 48-byte rows (16-byte rows for the 1x1
  // boxes of 4x4 maps) and TMA moves a box row by row: the 4x4 halo box (1536 rows) took 11 000 cycles to arrive, four
  // times per tile, in launches that are a single wave anyway.  Flat mode describes those tensors as [B, C, 1, H*W]: a
  // (image, channel) plane is ONE row of 256 / 64 bytes, the box carries no halo, and the transform warps supply the
  // zero padding themselves (a neighbour outside the image reads as 0).  Outputs and aux tiles use the same description.
  a.flat = (a.rows_per_tile == H && W <= 8) ? 1 : 0;
  a.box_w = (a.taps == 9 && !a.flat) ? W + 8 : W;
  a.box_h = (a.taps == 9 && !a.flat) ? a.rows_per_tile + 2 : a.rows_per_tile;
  a.raw_tx = a.box_w * a.box_h * kBlockK * a.imgs_per_tile * 4;
  a.raw_bytes = (a.raw_tx + 1023) / 1024 * 1024;
  a.n_tile = a.Cout <= 112 ? 112 : 128;
  // Weight tiles: only the rows that hold output channels travel (rounded up to the 8-row swizzle atom); the
  // MMA's N extent may reach past them into whatever follows in shared memory - those accumulator columns
  // belong to channels >= C_out and are never stored.
  const int n_here_max = a.Cout < kTileN ? a.Cout : kTileN;
  const int w_rows = (n_here_max + 7) / 8 * 8;
  a.w_tile_bytes = w_rows * kBlockK * wsz;
  a.w_tail_bytes = w_rows * 32;
  a.aux_mode = residual ? 1 : (gate ? 2 : 0);
  const int out_chunks = (n_here_max + 31) / 32;   // the output (and aux) tile is staged in the weight ring, 16 KB per 32 channels
  a.aux_chunks = a.aux_mode ? out_chunks : 0;
  // Shared memory: weight ring + raw-box ring + barriers.  Prefer a footprint that lets two CTAs share an SM
  // (dynamic + 1 KB reserved <= 114 KB each: one CTA's prologue / epilogue hides behind the other's main loop),
  // and within it a second raw box before a third weight stage: a single raw box costs ~5000 idle cycles at every
  // 32-channel block boundary (measured with the step trace), the weight tiles arrive in ~700.  Else one CTA per SM
  // with both rings deep.  The epilogue stages its output / aux tile over the drained rings.
  // 228 KB of shared memory per SM, 1 KB reserved per resident CTA
  const int two_cta_budget = 228 * 1024 / 2 - 1024, one_cta_budget = 226 * 1024;
  const int steps = a.k_blocks * a.taps;
  // bytes of the two rings; never less than the epilogue's staging tile, which is laid over them
  auto rings = [&](int n_raw, int n_b) {
    const int r = n_b * 2 * a.w_tile_bytes + n_raw * a.raw_bytes, o = out_chunks * kStageBytes;
    return r > o ? r : o;
  };
  auto total = [&](int n_raw, int n_b) { return rings(n_raw, n_b) + kConvBarrierBytes; };
  const int prefs[][2] = {{2, 4}, {2, 3}, {2, 2}, {1, 4}, {1, 3}, {1, 2}};
  a.n_b = 0;
  for (auto& pr : prefs)
    if ((F16 || pr[1] < 4 || pr[0] == 1) && total(pr[0], pr[1]) <= two_cta_budget) { a.n_raw = pr[0]; a.n_b = pr[1]; break; }
  if (a.n_b == 0) {
    a.n_raw = 2; a.n_b = kMaxStagesB;
    while (a.n_b > 2 && total(a.n_raw, a.n_b) > one_cta_budget) --a.n_b;
    if (total(a.n_raw, a.n_b) > one_cta_budget) a.n_raw = 1;
  }
  if (a.n_b > steps) a.n_b = steps;
  if (a.n_raw > a.k_blocks) a.n_raw = a.k_blocks;
#ifdef MFLOW_CONV_ABLATE
  { const char* e = getenv("MFLOW_CONV_NB"); if (e && atoi(e) >= 1 && atoi(e) <= kMaxStagesB) a.n_b = atoi(e); }
  { const char* e = getenv("MFLOW_CONV_NRAW"); if (e && atoi(e) >= 1 && atoi(e) <= kMaxRaw) a.n_raw = atoi(e); }
#endif
  a.ring_bytes = rings(a.n_raw, a.n_b) - a.n_raw * a.raw_bytes;   // weight ring, padded when the staging tile needs more
  MFLOW_REQUIRE(total(a.n_raw, a.n_b) <= one_cta_budget, "tile does not fit in shared memory");

  CUtensorMap mx, mhi, mlo, mhi8, mlo8, maux, my;
  { const int tail = a.Cin % kBlockK; a.tail_k8 = (tail > 0 && tail <= kTailK) ? 1 : 0; }
  // NCHW tensors as 4-D maps (W, H, C, B) - or, in flat mode, (H*W, 1, C, B): same rank and coordinates, rows of a whole plane
  const cuuint64_t d0 = a.flat ? (cuuint64_t)H * W : (cuuint64_t)W, d1 = a.flat ? 1 : (cuuint64_t)H;
  const cuuint64_t s1 = a.flat ? (cuuint64_t)H * W * 4 : (cuuint64_t)W * 4;
  const cuuint32_t b0x = a.flat ? (cuuint32_t)(H * W) : (cuuint32_t)W, b1x = a.flat ? 1u : (cuuint32_t)a.rows_per_tile;
  {  // output, NCHW: box {W, rows, 32 channels, images}
    cuuint64_t dims[4] = {d0, d1, (cuuint64_t)a.Cout, (cuuint64_t)B};
    cuuint64_t str[3] = {s1, (cuuint64_t)H * W * 4, (cuuint64_t)a.Cout * H * W * 4};
    cuuint32_t box[4] = {b0x, b1x, 32u, (cuuint32_t)a.imgs_per_tile};
    if (sd == nullptr && !encode(&my, y, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return MFLOW_E_BADARG;
  }
  {  // aux (residual / gate), shaped like y: box {W, rows, 32 channels, images}
    const float* auxp = residual ? residual : (gate ? gate : (y ? y : x));
    cuuint64_t dims[4] = {d0, d1, (cuuint64_t)a.Cout, (cuuint64_t)B};
    cuuint64_t str[3] = {s1, (cuuint64_t)H * W * 4, (cuuint64_t)a.Cout * H * W * 4};
    cuuint32_t box[4] = {b0x, b1x, 32u, (cuuint32_t)a.imgs_per_tile};
    if (!encode(&maux, auxp, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return MFLOW_E_BADARG;
  }
  {  // activations, NCHW: dims (W, H, C, B); box lands in shared memory as [image][channel][row][w]
    cuuint64_t dims[4] = {d0, d1, (cuuint64_t)a.Cin, (cuuint64_t)B};
    cuuint64_t str[3] = {s1, (cuuint64_t)H * W * 4, (cuuint64_t)a.Cin * H * W * 4};
    cuuint32_t box[4] = {a.flat ? b0x : (cuuint32_t)a.box_w, a.flat ? 1u : (cuuint32_t)a.box_h, (cuuint32_t)kBlockK, (cuuint32_t)a.imgs_per_tile};
    if (!encode(&mx, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return MFLOW_E_BADARG;
  }
  {  // weights: [taps][n_pad][k_pad], K-major swizzled tiles of w_rows x 32 channels (TF32: 128-byte rows, fp16: 64-byte rows)
    const CUtensorMapDataType dt = F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t dims[3] = {(cuuint64_t)d->k_pad, (cuuint64_t)d->n_pad, (cuuint64_t)a.taps};
    cuuint64_t str[2] = {(cuuint64_t)d->k_pad * wsz, (cuuint64_t)d->k_pad * d->n_pad * wsz};
    cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)w_rows, 1};
    const CUtensorMapSwizzle sw = F16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (!encode(&mhi, w_hi, 3, dims, str, box, sw, dt)) return MFLOW_E_BADARG;
    if (!encode(&mlo, w_lo, 3, dims, str, box, sw, dt)) return MFLOW_E_BADARG;
    cuuint32_t box8[3] = {(cuuint32_t)kTailK, (cuuint32_t)w_rows, 1};
    if (!encode(&mhi8, w_hi, 3, dims, str, box8, CU_TENSOR_MAP_SWIZZLE_32B, dt)) return MFLOW_E_BADARG;
    if (!encode(&mlo8, w_lo, 3, dims, str, box8, CU_TENSOR_MAP_SWIZZLE_32B, dt)) return MFLOW_E_BADARG;
  }
  const int n_tiles = a.imgs_per_tile > 1 ? (B + a.imgs_per_tile - 1) / a.imgs_per_tile : B * (H / a.rows_per_tile);
  dim3 grid(n_tiles, (a.Cout + kTileN - 1) / kTileN);
  const size_t smem = (size_t)total(a.n_raw, a.n_b);
  // the opt-in shared-memory limit is a per-device function attribute
  int dev = 0;
  cudaGetDevice(&dev);
#define MFLOW_CONV_LAUNCH_(CS_, SPL_)                                                                                      \
  do {                                                                                                                     \
    static bool attr_set[64] = {};   /* the opt-in shared-memory limit is a per-device function attribute */               \
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {                                                                          \
      cudaFuncSetAttribute(conv_split3_kernel<F16, CS_, SPL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one_cta_budget); \
      attr_set[dev] = true;                                                                                                \
    }                                                                                                                      \
    conv_split3_kernel<F16, CS_, SPL_><<<grid, kConvThreads, smem, (cudaStream_t)stream>>>(mx, mhi, mlo, mhi8, mlo8, maux, my, a); \
  } while (0)
#define MFLOW_CONV_LAUNCH(CS_) MFLOW_CONV_LAUNCH_(CS_, 0)
  const int cs = a.flat ? 0 : a.box_w * a.box_h;
  if (sd != nullptr) {   // spline epilogue: 1x1 boxes of the 32x32 / 16x16 levels (128 floats per channel) or flat boxes
    if constexpr (F16) {
      my = mx;           // never used: the epilogue writes y itself
      if (cs == 128) { if (sd->inverse) MFLOW_CONV_LAUNCH_(128, 2); else MFLOW_CONV_LAUNCH_(128, 1); }
      else { if (sd->inverse) MFLOW_CONV_LAUNCH_(0, 2); else MFLOW_CONV_LAUNCH_(0, 1); }
      return check_launch("conv_split3_kernel<fp16, spline epilogue>");
    } else {
      return MFLOW_E_BADARG;   // (unreachable: checked above)
    }
  }
  if (F16 && cs == 240) MFLOW_CONV_LAUNCH(240);        // 3x3 on 32x32 and 16x16 maps
  else if (F16 && cs == 160) MFLOW_CONV_LAUNCH(160);   // 3x3 on 8x8 maps
  else if (F16 && cs == 128) MFLOW_CONV_LAUNCH(128);   // 1x1
  else MFLOW_CONV_LAUNCH(0);
#undef MFLOW_CONV_LAUNCH
#undef MFLOW_CONV_LAUNCH_
  return check_launch(F16 ? "conv_split3_kernel<fp16>" : "conv_split3_kernel<tf32>");
}

extern "C" MFLOW_API int mflow_conv2d_tf32x3(const mflow_conv_desc* d, const float* x, const float* w_hi, const float* w_lo,
                                             const float* bias, const float* residual, const float* gate, float* y,
                                             void* stream) {
  return launch_conv<false>(d, x, w_hi, w_lo, nullptr, nullptr, bias, residual, gate, y, nullptr, stream);
}

extern "C" MFLOW_API int mflow_conv2d_f16x3(const mflow_conv_desc* d, const float* x, const void* w_hi, const void* w_lo,
                                            const float* w_unscale, const float* in_amax, const float* bias, const float* residual,
                                            const float* gate, float* y, float* out_amax, void* stream) {
  return launch_conv<true>(d, x, w_hi, w_lo, w_unscale, in_amax, bias, residual, gate, y, out_amax, stream);
}

namespace mflow {
// lad_sample[b] = sum of the sample's partial log-dets in index order (one warp per sample: strided sums, shuffle tree)
__global__ void __launch_bounds__(256) rqs_lad_reduce_kernel(const float* __restrict__ partial, int slots, float* __restrict__ lad, int B) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float s = 0.f;
  for (int i = lane; i < slots; i += 32) s += partial[(int64_t)b * slots + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) lad[b] = s;
}
}  // namespace mflow

extern "C" MFLOW_API int64_t mflow_conv2d_rqs_workspace_bytes(const mflow_conv_desc* d) {
  if (!d || d->batch < 1) return 0;
  return d->batch * (d->height * d->width / 16) * (d->c_out / 32) * (int64_t)sizeof(float);
}

extern "C" MFLOW_API int mflow_conv2d_rqs_f16x3(const mflow_conv_desc* d, const float* x, const void* w_hi, const void* w_lo,
                                                const float* w_unscale, const float* in_amax, const float* bias,
                                                const mflow_rqs_desc* sd, const float* x_full, float* y_full, float* lad_sample,
                                                void* workspace, void* stream) {
  MFLOW_REQUIRE(d && sd && x_full && y_full && lad_sample && workspace, "null pointer");
  if (d->batch == 0) return MFLOW_OK;
  if (int rc = launch_conv<true>(d, x, w_hi, w_lo, w_unscale, in_amax, bias, nullptr, nullptr, nullptr, nullptr, stream, sd, x_full, y_full,
                                 (float*)workspace))
    return rc;
  const int slots = (int)(d->height * d->width / 16 * (d->c_out / 32));
  rqs_lad_reduce_kernel<<<(unsigned)((d->batch + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, slots, lad_sample, (int)d->batch);
  return check_launch("rqs_lad_reduce_kernel");
}

// max |x| over n floats -> *out (zero-initialised by the caller; bits compared as unsigned): the operand scale of the
// fp16-split kernels for tensors whose producer did not publish it
namespace mflow {
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, int64_t n, float* out) {
  float m = 0.f;
  const int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream4(x4 + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(x[n4 * 4 + threadIdx.x]));
  const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  if ((threadIdx.x & 31) == 0) ptx::amax_publish(out, __uint_as_float(w));
}
}  // namespace mflow

extern "C" MFLOW_API int mflow_absmax(const float* x, int64_t n, float* out, void* stream) {
  MFLOW_REQUIRE(x && out && n >= 0, "null pointer");
  MFLOW_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  if (n == 0) return MFLOW_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  absmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  return check_launch("absmax_kernel");
}
