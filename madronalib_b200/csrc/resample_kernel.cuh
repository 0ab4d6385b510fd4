// resample_kernel.cuh -- K8: Upsampler / Downsampler banks (reference
// source/DSP/MLDSPFilters.h:1316-1473).  Lane per voice; the HalfBandFilter states of one voice are
// [octave][9] words in an SoA [octave*9 + k][V]; a launch runs n_blocks_in writes.
//  * Upsampler::write works in place at the end of its buffer list; the 2^octaves rows it leaves are
//    exactly what the following reads return, so the output planes themselves are the buffers.
//  * Downsampler::write keeps 2*octaves+1 rows per voice ([row][V][64] in delay memory) and a write
//    counter that is the same for every voice (kept on the host).
#pragma once
#include "functors.cuh"

namespace mlb
{
struct ResampleArgs
{
  const float* in;   // [T][V][64]
  float* out;        // up: [T << oct][V][64]; down: [produced][V][64]
  uint32_t* state;   // [oct * 9][V]
  float* buf;        // down only: [2 * oct + 1][V][64]
  int V, T, oct;
  unsigned counter;  // down: Downsampler::_counter at the start of the launch
};

MLB_DEV void hb_load(HalfBand& h, const uint32_t* st, int j, size_t V, int v)
{
#pragma unroll
  for (int i = 0; i < 9; ++i) h.s[i] = u2f(st[(size_t)(j * 9 + i) * V + v]);
}
MLB_DEV void hb_store(const HalfBand& h, uint32_t* st, int j, size_t V, int v)
{
#pragma unroll
  for (int i = 0; i < 9; ++i) st[(size_t)(j * 9 + i) * V + v] = f2u(h.s[i]);
}
// upsampleFirstHalf + upsampleSecondHalf of one row (F:1248-1270); d2 may be the source row itself
MLB_DEV void hb_up_row(HalfBand& h, const float* src, float* d1, float* d2)
{
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 1
  for (int q = 0; q < 16; ++q)
  {
    const float4 x = s4[q];
    float4 lo, hi;
    lo.x = h.a<true>(x.x), lo.y = h.b<true>(x.x), lo.z = h.a<true>(x.y), lo.w = h.b<true>(x.y);
    hi.x = h.a<true>(x.z), hi.y = h.b<true>(x.z), hi.z = h.a<true>(x.w), hi.w = h.b<true>(x.w);
    float4* d4 = reinterpret_cast<float4*>(q < 8 ? d1 : d2) + (q & 7) * 2;
    d4[0] = lo;
    d4[1] = hi;
  }
}
// downsample(vx1, vx2) (F:1272-1294)
MLB_DEV void hb_down_rows(HalfBand& h, const float* x1, const float* x2, float* dst)
{
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 1
  for (int q = 0; q < 16; ++q)
  {
    const float4* s4 = reinterpret_cast<const float4*>(q < 8 ? x1 : x2) + (q & 7) * 2;
    const float4 p0 = s4[0], p1 = s4[1];
    float4 y;
    float a0, b0;
    a0 = h.a<true>(p0.x), b0 = h.b<true>(p0.y), y.x = __fmul_rn(__fadd_rn(a0, h.s[8]), 0.5f), h.s[8] = b0;
    a0 = h.a<true>(p0.z), b0 = h.b<true>(p0.w), y.y = __fmul_rn(__fadd_rn(a0, h.s[8]), 0.5f), h.s[8] = b0;
    a0 = h.a<true>(p1.x), b0 = h.b<true>(p1.y), y.z = __fmul_rn(__fadd_rn(a0, h.s[8]), 0.5f), h.s[8] = b0;
    a0 = h.a<true>(p1.z), b0 = h.b<true>(p1.w), y.w = __fmul_rn(__fadd_rn(a0, h.s[8]), 0.5f), h.s[8] = b0;
    d4[q] = y;
  }
}
MLB_DEV void row_copy(const float* src, float* dst)
{
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
  for (int q = 0; q < 16; ++q) d4[q] = s4[q];
}

__global__ void __launch_bounds__(128) upsample_kernel(const ResampleArgs a)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= a.V) return;
  const size_t V = (size_t)a.V;
  const int N = 1 << a.oct;
  for (int t = 0; t < a.T; ++t)
  {
    // Upsampler::write, F:1428-1453: rows (t*N + k) of the output are bufferPtr(k)
    auto row = [&](int k) { return a.out + (((size_t)t * N + k) * V + v) * MLB_BLOCK; };
    row_copy(a.in + ((size_t)t * V + v) * MLB_BLOCK, row(N - 1));
    for (int j = 0; j < a.oct; ++j)
    {
      HalfBand h;
      hb_load(h, a.state, j, V, v);
      const int sourceBufs = 1 << j, srcStart = N - sourceBufs, destStart = N - 2 * sourceBufs;
      for (int i = 0; i < sourceBufs; ++i) hb_up_row(h, row(srcStart + i), row(destStart + 2 * i), row(destStart + 2 * i + 1));
      hb_store(h, a.state, j, V, v);
    }
  }
}

__global__ void __launch_bounds__(128) downsample_kernel(const ResampleArgs a)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= a.V) return;
  const size_t V = (size_t)a.V;
  const unsigned wrap = (1u << a.oct) - 1u;
  unsigned counter = a.counter;
  int produced = 0;
  auto buf = [&](int k) { return a.buf + ((size_t)k * V + v) * MLB_BLOCK; };
  for (int t = 0; t < a.T; ++t)
  {
    // Downsampler::write, F:1347-1386
    row_copy(a.in + ((size_t)t * V + v) * MLB_BLOCK, buf((int)(counter & 1u)));
    unsigned mask = 1;
    for (int hh = 0; hh < a.oct; ++hh)
    {
      if (!(counter & mask)) break;
      mask <<= 1;
      const int b1 = (counter & mask) != 0;
      HalfBand h;
      hb_load(h, a.state, hh, V, v);
      hb_down_rows(h, buf(hh * 2), buf(hh * 2 + 1), buf(hh * 2 + 2 + b1));
      hb_store(h, a.state, hh, V, v);
    }
    counter = (counter + 1u) & wrap;
    if (counter == 0) row_copy(buf(2 * a.oct), a.out + ((size_t)(produced++) * V + v) * MLB_BLOCK);
  }
}

}  // namespace mlb
