// Conditioner convolutions on the 5th-generation tensor cores (kernel (b) of the north star).
//
//   y[b, co, h, w] = bias[co] + sum_{tap, ci} relu?(x)[b, ci, h + dy, w + dx] * W[co, ci, tap]   (+ residual)
//
// for the 3x3 (pad 1) and 1x1 convolutions of ConvResidualNet (reference: manifold_flow/nn/resnet.py:75-142),
// as an implicit GEMM per 128-pixel tile:  D[128 pixels, N out channels] += A_tap[128, 32 ci] * B_tap[N, 32 ci]^T.
//
// * Activations stay in the reference's NCHW fp32 layout: no im2col, no layout conversion and no channel
//   padding in HBM.  One TMA box {W, rows, 32 channels, images} per (tap, channel block) lands in shared
//   memory as [channel][pixel]; the tap shift is just the TMA coordinate, and the zero padding of the
//   convolution and of channels 100..127 is TMA's out-of-bounds fill.
// * fp32 parity on TF32 tensor cores by the 3xTF32 split: a = a_hi + a_lo with a_hi the TF32 truncation;
//   D += A_hi B_hi + A_lo B_hi + A_hi B_lo.  Weights are split once per optimizer step on the host side of
//   the ABI.  Activations are handled by the CTA's four transform warps between TMA arrival and the MMA:
//   one thread per pixel reads its 32 channels (conflict-free), applies the pre-activation ReLU, splits,
//   and writes its 128-byte row of the K-major, 128B-swizzled UMMA operand tiles (A_hi, A_lo) - i.e. the
//   NCHW -> [pixel][channel] transposition costs no extra pass.
// * Accumulator: 128 lanes x 128 columns of TMEM; epilogue tcgen05.ld 32x32b gives every thread one pixel
//   row, so for a fixed output channel a warp stores 32 consecutive pixels (coalesced NCHW writes) with
//   bias / residual / ReLU fused.
#include <stdarg.h>
#include <stdlib.h>

#include "mflow_common.cuh"
#include "tc_ptx.cuh"

namespace mflow {

using namespace ptx;

constexpr int kTileM = 128;   // pixels per CTA tile
constexpr int kTileN = 128;   // output channels per CTA tile (TMEM columns)
constexpr int kBlockK = 32;   // input channels per pipeline stage
constexpr int kStageBytes = kTileM * kBlockK * 4;  // 16 KB: one A or B operand tile
constexpr int kConvThreads = 256;  // warp 0: TMA + MMA issue; warps 4-7: transform + epilogue

struct ConvArgs {
  float* y;
  const float* bias;      // [Cout] or null
  const float* residual;  // same shape as y or null
  int B, Cin, Cout, H, W;
  int taps;               // 9 (3x3, pad 1) or 1
  int relu_in;            // apply relu to the input on the fly
  int relu_out;           // apply relu to the output
  int rows_per_tile;      // image rows per 128-pixel tile (box H extent)
  int imgs_per_tile;      // images per tile (box B extent)
  int k_blocks;           // ceil(Cin / 32)
  int box_w, box_h;       // raw activation box extents (W + 8, rows + 2 with a halo for 3x3; W, rows for 1x1)
  int raw_bytes;          // bytes of one raw box (all 32 channels, all images of the tile)
  int debug;              // bring-up: 0 = full kernel; 1 stop after setup; 2 after first TMA; 3 after transform; 4 after first MMA
};

__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// round to nearest TF32 (ties away): the tensor core truncates its fp32 inputs, so the low part of the
// split is pre-rounded to keep the truncation from biasing the sum
__device__ __forceinline__ float tf32_round(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kConvThreads, 1)
conv_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                   const __grid_constant__ CUtensorMap map_wlo, const ConvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* s_ahi = smem;
  uint8_t* s_alo = smem + kStageBytes;
  float* s_bhi = (float*)(smem + 2 * kStageBytes);
  float* s_blo = (float*)(smem + 3 * kStageBytes);
  float* s_raw = (float*)(smem + 4 * kStageBytes);
  __shared__ __align__(8) uint64_t bar_raw, bar_full, bar_mma;
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * kTileN;
  // tile -> first (image, row)
  const int tiles_per_img = a.imgs_per_tile > 1 ? 1 : a.H / a.rows_per_tile;
  const int b0 = a.imgs_per_tile > 1 ? tile * a.imgs_per_tile : tile / tiles_per_img;
  const int h0 = a.imgs_per_tile > 1 ? 0 : (tile % tiles_per_img) * a.rows_per_tile;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_whi);
    prefetch_tensormap(&map_wlo);
    mbar_init(&bar_raw, 1);
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&s_tmem, kTileN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const uint32_t idesc = idesc_tf32(kTileM, kTileN, /*a K-major*/ 0, /*b K-major*/ 0);
  // pixel owned by this transform / epilogue thread
  const int pm = (threadIdx.x - 128) & 127;
  const int pw = pm % a.W, pg = pm / a.W;
  const int phl = pg % a.rows_per_tile, pbl = pg / a.rows_per_tile;
  const int chan_stride = a.box_h * a.box_w;                            // floats between channels in the raw box
  const int raw_base = pbl * kBlockK * chan_stride + phl * a.box_w + pw;  // raw box layout [image][channel][row][w]
  const int pad = a.taps == 9 ? 1 : 0;
  uint32_t phase = 0, raw_phase = 0;
  uint32_t accumulate = 0;
  for (int kb = 0; kb < a.k_blocks; ++kb) {
    // one raw box per channel block: the tile's pixels plus a one-pixel halo (the innermost TMA
    // coordinate must be a multiple of 16 bytes, so the box starts 4 pixels to the left); all 9 taps
    // of the 3x3 stencil are cut out of it by the transform warps
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar_raw, a.raw_bytes);
      tma_load_4d(s_raw, &map_x, &bar_raw, pad ? -4 : 0, h0 - pad, kb * kBlockK, b0);
    }
    const int valid = min(kBlockK, a.Cin - kb * kBlockK);
    const int ksteps = (valid + 7) / 8;
    for (int tap = 0; tap < a.taps; ++tap) {
      const int dy = pad ? tap / 3 : 0, dx = pad ? tap % 3 : 0;
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_full, 2 * kStageBytes);
        tma_load_3d(s_bhi, &map_whi, &bar_full, kb * kBlockK, n0, tap);
        tma_load_3d(s_blo, &map_wlo, &bar_full, kb * kBlockK, n0, tap);
      }
      if (tap == 0) { mbar_wait(&bar_raw, raw_phase); raw_phase ^= 1; }
      // transform: thread = pixel; gather its 32 channels of this tap, optional ReLU, TF32 hi / lo split,
      // write the pixel's 128-byte row of both K-major operand tiles (16-byte chunk j -> slot j ^ (row & 7))
      if (warp >= 4) {
        uint8_t* row_hi = s_ahi + pm * 128;
        uint8_t* row_lo = s_alo + pm * 128;
        const float* src = s_raw + raw_base + (pad ? dy * a.box_w + dx + 3 : 0);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float4 h, l;
          float v0 = src[(4 * jj + 0) * chan_stride], v1 = src[(4 * jj + 1) * chan_stride];
          float v2 = src[(4 * jj + 2) * chan_stride], v3 = src[(4 * jj + 3) * chan_stride];
          if (a.relu_in) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
          h.x = tf32_trunc(v0); h.y = tf32_trunc(v1); h.z = tf32_trunc(v2); h.w = tf32_trunc(v3);
          l.x = tf32_round(v0 - h.x); l.y = tf32_round(v1 - h.y); l.z = tf32_round(v2 - h.z); l.w = tf32_round(v3 - h.w);
          const int slot = (jj ^ (pm & 7)) * 16;
          *reinterpret_cast<float4*>(row_hi + slot) = h;
          *reinterpret_cast<float4*>(row_lo + slot) = l;
        }
        fence_proxy_async();
      }
      mbar_wait(&bar_full, phase);
      __syncthreads();
      if (threadIdx.x == 0) {
        tc_fence_after();
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t d_ahi = smem_desc(smem_u32(s_ahi) + kk * 32, 16, 1024, 2);
          const uint64_t d_alo = smem_desc(smem_u32(s_alo) + kk * 32, 16, 1024, 2);
          const uint64_t d_bhi = smem_desc(smem_u32(s_bhi) + kk * 32, 16, 1024, 2);
          const uint64_t d_blo = smem_desc(smem_u32(s_blo) + kk * 32, 16, 1024, 2);
          umma_tf32(tmem, d_ahi, d_bhi, idesc, accumulate);
          umma_tf32(tmem, d_alo, d_bhi, idesc, 1);
          umma_tf32(tmem, d_ahi, d_blo, idesc, 1);
          accumulate = 1;
        }
        umma_commit(&bar_mma);
      }
      mbar_wait(&bar_mma, phase);  // operands consumed: the tiles (and, after the last tap, the raw box) may be overwritten
      phase ^= 1;
    }
  }
  // ---- epilogue: TMEM -> registers -> NCHW global, bias / residual / ReLU fused ----
  tc_fence_after();
  if (warp >= 4) {
    const int ew = warp - 4;           // TMEM lane quarter this warp may access
    const int w = pw;                  // pixel index inside the tile == TMEM lane == pm
    const int b = b0 + pbl, h = h0 + phl;
    const bool ok = b < a.B;
    const int64_t plane = (int64_t)a.H * a.W;
    const int64_t base = ((int64_t)b * a.Cout) * plane + (int64_t)h * a.W + w;
#pragma unroll 1
    for (int c0 = 0; c0 < kTileN; c0 += 32) {
      if (n0 + c0 >= a.Cout) break;
      float v[32];
      tmem_ld_32x32(tmem + ((uint32_t)(ew * 32) << 16) + c0, v);
      if (ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = n0 + c0 + j;
          if (co < a.Cout) {
            float r = v[j] + (a.bias ? a.bias[co] : 0.f);
            if (a.residual) r += a.residual[base + co * plane];
            if (a.relu_out) r = fmaxf(r, 0.f);
            a.y[base + co * plane] = r;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTileN);
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static bool encode(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = get_encode();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return false; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return false; }
  return true;
}

}  // namespace mflow

using namespace mflow;

extern "C" MFLOW_API int mflow_conv2d_tf32x3(const mflow_conv_desc* d, const float* x, const float* w_hi, const float* w_lo,
                                             const float* bias, const float* residual, float* y, void* stream) {
  MFLOW_REQUIRE(d && x && w_hi && w_lo && y, "null pointer");
  MFLOW_REQUIRE(d->kernel == 3 || d->kernel == 1, "kernel size must be 1 or 3");
  MFLOW_REQUIRE(d->width == 32 || d->width == 16 || d->width == 8 || d->width == 4, "width must be 4, 8, 16 or 32");
  MFLOW_REQUIRE(d->height == d->width, "square feature maps only");
  MFLOW_REQUIRE(d->c_in >= 1 && d->c_out >= 1 && d->batch >= 0, "bad extents");
  MFLOW_REQUIRE(d->k_pad % kBlockK == 0 && d->k_pad >= d->c_in && d->n_pad % kTileN == 0 && d->n_pad >= d->c_out,
                "weights must be padded to [taps][n_pad %% 128 == 0][k_pad %% 32 == 0]");
  if (d->batch == 0) return MFLOW_OK;
  const int W = (int)d->width, H = (int)d->height, B = (int)d->batch;
  ConvArgs a{};
  a.y = y; a.bias = bias; a.residual = residual;
  a.B = B; a.Cin = (int)d->c_in; a.Cout = (int)d->c_out; a.H = H; a.W = W;
  a.taps = d->kernel == 3 ? 9 : 1;
  a.relu_in = d->relu_in; a.relu_out = d->relu_out;
  const int groups = kTileM / W;                       // (image,row) groups per tile
  a.rows_per_tile = groups < H ? groups : H;
  a.imgs_per_tile = groups / a.rows_per_tile;
  a.k_blocks = (a.Cin + kBlockK - 1) / kBlockK;
  a.box_w = a.taps == 9 ? W + 8 : W;
  a.box_h = a.taps == 9 ? a.rows_per_tile + 2 : a.rows_per_tile;
  a.raw_bytes = a.box_w * a.box_h * kBlockK * a.imgs_per_tile * 4;
  a.debug = 0;
  CUtensorMap mx, mhi, mlo;
  {  // activations, NCHW: dims (W, H, C, B); box lands in shared memory as [image][channel][row][w]
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)a.Cin, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)a.Cin * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)a.box_w, (cuuint32_t)a.box_h, (cuuint32_t)kBlockK, (cuuint32_t)a.imgs_per_tile};
    if (!encode(&mx, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return MFLOW_E_BADARG;
  }
  {  // weights: [taps][n_pad][k_pad], K-major 128B-swizzled tiles of 128 x 32
    cuuint64_t dims[3] = {(cuuint64_t)d->k_pad, (cuuint64_t)d->n_pad, (cuuint64_t)a.taps};
    cuuint64_t str[2] = {(cuuint64_t)d->k_pad * 4, (cuuint64_t)d->k_pad * d->n_pad * 4};
    cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileN, 1};
    if (!encode(&mhi, w_hi, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MFLOW_E_BADARG;
    if (!encode(&mlo, w_lo, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MFLOW_E_BADARG;
  }
  const int n_tiles = a.imgs_per_tile > 1 ? (B + a.imgs_per_tile - 1) / a.imgs_per_tile : B * (H / a.rows_per_tile);
  dim3 grid(n_tiles, (a.Cout + kTileN - 1) / kTileN);
  const size_t smem = 4 * kStageBytes + a.raw_bytes + 1024;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(conv_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  conv_tf32x3_kernel<<<grid, kConvThreads, smem, (cudaStream_t)stream>>>(mx, mhi, mlo, a);
  return check_launch("conv_tf32x3_kernel");
}
