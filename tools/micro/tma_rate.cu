// Micro-benchmark: how fast can one SM pull a 104 x 128-byte weight tile out of L2 - as a tiled TMA box (one request
// per 128-byte row) or as one 1-D bulk copy of the same 13 KB - when all 148 SMs do it at once?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../manifold-flow_b200/csrc tma_rate.cu -o tma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
#include "tma_encode.cuh"
namespace mflow { void set_error(const char* fmt, ...) { (void)fmt; } }
using namespace mflow; using namespace mflow::ptx;

constexpr int kRows = 104, kTile = kRows * 128, kSlots = 4, kIters = 256;

__device__ __forceinline__ void bulk_load(void* smem, const void* gptr, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
               "l"(gptr), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void rate_kernel(const __grid_constant__ CUtensorMap map, const float* w, int mode, int n_tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kSlots];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    // keep kSlots loads in flight; wait for the oldest before reusing its slot
    for (int i = 0; i < kIters + kSlots; ++i) {
      const int s = i % kSlots;
      if (i >= kSlots) mbar_wait(&full[s], ((i / kSlots) - 1) & 1);
      if (i < kIters) {
        const int tile = (i * 7 + blockIdx.x) % n_tiles;
        mbar_expect_tx(&full[s], kTile);
        if (mode == 0) tma_load_3d(smem + s * kTile, &map, &full[s], 0, 0, tile);
        else bulk_load(smem + s * kTile, w + (size_t)tile * (kTile / 4), kTile, &full[s]);
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int n_tiles = 72;   // 36 steps x hi/lo of one convolution: 958 KB, L2 resident
  float* w; cudaMalloc(&w, (size_t)n_tiles * kTile); cudaMemset(w, 0, (size_t)n_tiles * kTile);
  long long* out; cudaMalloc(&out, 148 * sizeof(long long));
  CUtensorMap map;
  cuuint64_t dims[3] = {32, kRows, (cuuint64_t)n_tiles};
  cuuint64_t str[2] = {128, (cuuint64_t)kTile};
  cuuint32_t box[3] = {32, kRows, 1};
  if (!encode(&map, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("encode failed\n"); return 1; }
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kTile);
  for (int ctas = 148; ctas <= 296; ctas += 148)
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 2; ++rep) rate_kernel<<<ctas, 32, kSlots * kTile>>>(map, w, mode, n_tiles, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      long long h[296]; cudaMemcpy(h, out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("%3d CTAs, %s: %.0f cycles per 13 KB tile per CTA (%.1f B/clk/CTA)\n", ctas, mode ? "1-D bulk copy " : "tiled TMA box ",
             avg / kIters, kTile / (avg / kIters));
    }
  return 0;
}
