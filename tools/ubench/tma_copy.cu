// Microbenchmark: pure TMA streaming copy in[T][V][64] -> out[T][V][64] with the chain kernel's
// work decomposition (one warp per 32 voices walking time, warp-private stage ring), for
// different tile shapes.  Isolates what the memory system delivers for each access pattern.
//   mode 0: half-row tiles  box {32 samples, 32 voices}, 4 KB, order (t,h)   [round-1a kernel]
//   mode 1: full-row tiles through a 4-D map {32, V, 2 halves, T} box {32,32,2,1}: 8 KB per op,
//           lands as two standard 128B-swizzled 4 KB half tiles
//   mode 2: like mode 0 but the two half tiles of a block are requested back to back
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../madronalib_b200/csrc/tma.cuh"
using namespace mlb;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int MODE>
__global__ void copy_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                            int V, int T, int S, int touch)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int group = blockIdx.x * W + warp;
  const int v0 = group * 32;
  if (v0 >= V) return;
  constexpr int TB = (MODE == 1) ? 8192 : 4096;
  const uint32_t base = smem_u32(smem);
  const uint32_t tiles = base + warp * S * TB;
  const uint32_t bars = base + W * S * TB + warp * S * 8;
  if (lane == 0) { for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1); fence_mbar_init(); }
  __syncwarp();
  const int total = (MODE == 1) ? T : 2 * T;
  auto issue = [&](int k, int s) {
    const uint32_t bar = bars + 8 * s;
    mbar_arrive_expect_tx(bar, TB);
    if (MODE == 1) tma_load_4d(tiles + s * TB, &in_map, bar, 0, v0, 0, k);
    else tma_load_3d(tiles + s * TB, &in_map, bar, (k & 1) * 32, v0, k >> 1, kEvictFirst);
  };
  if (lane == 0) for (int k = 0; k < S - 1 && k < total; ++k) issue(k, k);
  int s = 0; uint32_t parity = 0; float sink = 0.f;
  for (int k = 0; k < total; ++k)
  {
    const uint32_t tile = tiles + s * TB;
    mbar_wait(bars + 8 * s, parity);
    if (touch) { // read + rewrite own row, like the real kernel
      for (int q = 0; q < TB / 4096; ++q)
        for (int j = 0; j < 8; ++j) { uint32_t a = tile + q * 4096 + lane * 128 + ((j ^ (lane & 7)) << 4); float4 x = lds128(a); x.x += 1.f; sts128(a, x); }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0)
    {
      if (MODE == 1) tma_store_4d(&out_map, tile, 0, v0, 0, k);
      else tma_store_3d(&out_map, tile, (k & 1) * 32, v0, k >> 1);
      bulk_commit();
      const int kn = k + S - 1;
      if (kn < total) { if (k >= 1) bulk_wait_read<1>(); issue(kn, s == 0 ? S - 1 : s - 1); }
    }
    if (++s == S) { s = 0; parity ^= 1; }
  }
  if (lane == 0) bulk_wait_read<0>();
  __syncwarp();
  if (sink == 123.f) printf("x");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
  const int V = 65536, T = 64;
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  float *in, *out; size_t n = (size_t)T * V * 64;
  cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4); cudaMemset(in, 0, n * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // reference: plain device-to-device memcpy
  for (int i = 0; i < 3; ++i) cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice);
  cudaEventRecord(e0); for (int i = 0; i < 10; ++i) cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice); cudaEventRecord(e1);
  cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemcpy D2D: %.3f ms  %.0f GB/s\n", ms / 10, 2.0 * n * 4 / (ms / 10 * 1e-3) / 1e9);
  for (int mode = 0; mode < 2; ++mode)
  {
    CUtensorMap mi, mo;
    for (int io = 0; io < 2; ++io)
    {
      CUtensorMap* m = io ? &mo : &mi; void* base = io ? (void*)out : (void*)in; CUresult r;
      if (mode == 1) {
        cuuint64_t dims[4] = {32, (cuuint64_t)V, 2, (cuuint64_t)T}; cuuint64_t str[3] = {256, 128, (cuuint64_t)V * 256};
        cuuint32_t box[4] = {32, 32, 2, 1}, es[4] = {1, 1, 1, 1};
        r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint64_t dims[3] = {64, (cuuint64_t)V, (cuuint64_t)T}; cuuint64_t str[2] = {256, (cuuint64_t)V * 256};
        cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
        r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (r != CUDA_SUCCESS) { printf("mode %d: encode failed %d\n", mode, (int)r); return 1; }
    }
    struct Cfg { int W, S; } cfgs[] = {{1, 3}, {2, 3}, {1, 2}, {14, 2}, {14, 3}, {1, 4}, {1, 6}};
    for (int touch = 0; touch < 2; ++touch)
      for (Cfg c : cfgs)
      {
        const int TB = mode == 1 ? 8192 : 4096;
        size_t smem = (size_t)c.W * c.S * TB + c.W * c.S * 8;
        if (smem > 227 * 1024) continue;
        int groups = V / 32, ctas = (groups + c.W - 1) / c.W;
        cudaError_t e;
        auto launch = [&]() {
          if (mode == 0) { cudaFuncSetAttribute(copy_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); copy_kernel<0><<<ctas, c.W * 32, smem>>>(mi, mo, V, T, c.S, touch); }
          else { cudaFuncSetAttribute(copy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); copy_kernel<1><<<ctas, c.W * 32, smem>>>(mi, mo, V, T, c.S, touch); }
        };
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0); for (int i = 0; i < 10; ++i) launch(); cudaEventRecord(e1);
        e = cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        int occ = 0;
        if (mode == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, copy_kernel<0>, c.W * 32, smem);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, copy_kernel<1>, c.W * 32, smem);
        printf("mode %d touch %d W=%2d S=%d smem/CTA %6zu B occ %2d CTAs/SM (%3d warps): %.3f ms  %.0f GB/s  %s\n", mode, touch, c.W, c.S,
               smem, occ, occ * c.W, ms / 10, 2.0 * n * 4 / (ms / 10 * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  }
  return 0;
}
