// chain_kernel.cuh -- fused linear voice chains: one kernel launch = n_blocks Bank::operator()
// calls (reference MLDSPFunctional.h:328-337) for every voice.
//
// Work decomposition (DESIGN.md "K1/K2"):
//   * one LANE per voice: all oscillator / filter state and coefficients live in registers
//     for the whole launch (loaded once, stored once, SoA, coalesced 4 B/lane);
//   * one WARP per 32 consecutive voices; warps are fully independent (no __syncthreads):
//     each warp owns a private ring of `stages` 4 KB shared-memory tiles and its own
//     mbarriers, and walks time sequentially (the recurrences cannot be split in time);
//   * a tile is 32 voices x 32 samples of one plane of the reference layout
//     [plane][V][64] f32, moved by ONE TMA instruction (cp.async.bulk.tensor.3d, box
//     {32 samples, 32 voices, 1 plane}, SWIZZLE_128B).  With the 128-byte swizzle lane r
//     finds 16-byte chunk j of its own row at  r*128 + ((j ^ (r & 7)) << 4): the eight
//     lanes of every quarter-warp hit eight different 16-byte bank groups, so the
//     per-lane LDS.128 / STS.128 row walk is bank-conflict free;
//   * results are written IN PLACE into the tile and leave through a TMA store
//     (UTMASTG); ragged V is clipped by the tensor map (zero fill on load, clip on store);
//   * optional mix bus: after a tile is computed, lane n sums column n over the 32 voice
//     rows (conflict-free through the same swizzle) into partial[plane][group][64];
//     a second tiny kernel adds the per-group partials in group order (deterministic).
#pragma once
#include "ops.cuh"
#include "tma.cuh"

namespace mlb
{
constexpr int kTileSamples = 32;
constexpr int kTileVoices = 32;
constexpr int kTileBytes = kTileSamples * kTileVoices * 4;  // 4096
constexpr int kMaxChainState = 8;
constexpr int kMaxChainCoef = 16;

__host__ __device__ constexpr int op_ns(int op)
{
  switch (op)
  {
#define MLB_X_NS(NAME, id, nin, nst, nco) \
  case id: return nst;
    MLB_OP_TABLE(MLB_X_NS)
#undef MLB_X_NS
  }
  return 0;
}
__host__ __device__ constexpr int op_nc(int op)
{
  switch (op)
  {
#define MLB_X_NC(NAME, id, nin, nst, nco) \
  case id: return nco;
    MLB_OP_TABLE(MLB_X_NC)
#undef MLB_X_NC
  }
  return 0;
}

enum ChainSrc
{
  SRC_INPUT = 0,  // generator/filter input = external signal plane (Contract R)
  SRC_PARAM = 1,  // = per-voice scalar broadcast, DSPVector(float) (Contract S)
  SRC_NONE = 2    // generator takes no input (NoiseGen)
};

struct ChainArgs
{
  uint32_t* state;     // [n_state_words][V]
  const float* coef;   // [n_coef_words][V]
  float* mix_partial;  // [T*n_out_planes][n_groups][64] or nullptr
  int V, T;
  int n_in_planes, in_plane;    // tile z = t*n_in_planes + in_plane
  int n_out_planes, out_plane;  // tile z = t*n_out_planes + out_plane
  int n_groups;                 // ceil(V/32)
  int write_out;                // store per-voice output planes
  int stages;
  int st_idx[kMaxChainState];  // SoA word index of each register state slot
  int co_idx[kMaxChainCoef];
};

// GEN: generator op id or -1 (the chain filters the source directly)
// F1, F2: filter op ids or -1;  GAIN: multiply by a PARAM at the end
template <int GEN, int SRC, int F1, int F2, bool GAIN, bool EX>
struct Chain
{
  static constexpr int NS_GEN = GEN >= 0 ? op_ns(GEN) : 0;
  static constexpr int NS_F1 = F1 >= 0 ? op_ns(F1) : 0;
  static constexpr int NS_F2 = F2 >= 0 ? op_ns(F2) : 0;
  static constexpr int NS = NS_GEN + NS_F1 + NS_F2;
  static constexpr int NC_SRC = (SRC == SRC_PARAM) ? 1 : 0;
  static constexpr int NC_F1 = F1 >= 0 ? op_nc(F1) : 0;
  static constexpr int NC_F2 = F2 >= 0 ? op_nc(F2) : 0;
  static constexpr int NC = NC_SRC + NC_F1 + NC_F2 + (GAIN ? 1 : 0);
  static constexpr bool HAS_IN = (SRC == SRC_INPUT);
  static_assert(NS <= kMaxChainState && NC <= kMaxChainCoef, "chain too large");

  static MLB_DEV float tick(float in, uint32_t (&st)[NS > 0 ? NS : 1],
                            const float (&co)[NC > 0 ? NC : 1])
  {
    float x = (SRC == SRC_INPUT) ? in : ((SRC == SRC_PARAM) ? co[0] : 0.0f);
    float y = x;
    if (GEN >= 0) y = gen_tick<EX>(GEN, x, 0.0f, &st[0]);
    if (F1 >= 0) y = filter_tick<EX>(F1, y, &st[NS_GEN], &co[NC_SRC]);
    if (F2 >= 0) y = filter_tick<EX>(F2, y, &st[NS_GEN + NS_F1], &co[NC_SRC + NC_F1]);
    if (GAIN) y = A<EX>::mul(y, co[NC - 1]);  // operator*(DSPVector, DSPVector(float)), O:345-348
    return y;
  }
};

template <class P>
__global__ void __launch_bounds__(128)
    chain_kernel(const __grid_constant__ CUtensorMap in_map,
                 const __grid_constant__ CUtensorMap out_map, const ChainArgs a)
{
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // SWIZZLE_128B atom = 1024 B
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int W = blockDim.x >> 5;
  const int S = a.stages;
  const int group = blockIdx.x * W + warp;
  const int v0 = group * kTileVoices;
  if (v0 >= a.V) return;  // warp-uniform; warps never synchronise with each other
  const int v = v0 + lane;
  const bool live = v < a.V;

  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();  // the swizzle formula below assumes 1024-byte aligned tiles
  const uint32_t tiles = base + (uint32_t)(warp * S) * kTileBytes;
  const uint32_t bars = base + (uint32_t)(W * S) * kTileBytes + (uint32_t)(warp * S) * 8u;

  if (lane == 0)
  {
    if (P::HAS_IN)
    {
      prefetch_tensormap(&in_map);
      for (int s = 0; s < S; ++s) mbar_init(bars + 8u * s, 1);
      fence_mbar_init();
    }
    if (a.write_out) prefetch_tensormap(&out_map);
  }
  __syncwarp();

  // ---- state and coefficients: HBM -> registers, once per launch ----
  uint32_t st[P::NS > 0 ? P::NS : 1];
  float co[P::NC > 0 ? P::NC : 1];
#pragma unroll
  for (int i = 0; i < P::NS; ++i) st[i] = live ? a.state[(size_t)a.st_idx[i] * a.V + v] : 0u;
#pragma unroll
  for (int i = 0; i < P::NC; ++i) co[i] = live ? a.coef[(size_t)a.co_idx[i] * a.V + v] : 0.0f;

  const int total = a.T * 2;  // two 32-sample tiles per 64-sample block
  // prologue: fill S-1 stages
  if (P::HAS_IN && lane == 0)
  {
    const int pre = (S - 1 < total) ? (S - 1) : total;
    for (int k = 0; k < pre; ++k)
    {
      const uint32_t bar = bars + 8u * k;
      mbar_arrive_expect_tx(bar, kTileBytes);
      tma_load_3d(tiles + (uint32_t)k * kTileBytes, &in_map, bar, (k & 1) * kTileSamples, v0,
                  (k >> 1) * a.n_in_planes + a.in_plane, kEvictFirst);
    }
  }

  const uint32_t row_off = (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7) << 4;

  uint32_t mix_off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    mix_off[j] = ((((uint32_t)lane >> 2) ^ (uint32_t)j) << 4) + (((uint32_t)lane & 3u) << 2);
  const bool full_group = (v0 + kTileVoices <= a.V);
  // partial[t*n_out + plane][group][h*32 + lane]; advanced by one block every second tile
  float* mix_row = a.mix_partial
                       ? a.mix_partial + ((size_t)a.out_plane * a.n_groups + group) * MLB_BLOCK + lane
                       : nullptr;
  const size_t mix_block_stride = (size_t)a.n_out_planes * a.n_groups * MLB_BLOCK;
  int s = 0;            // stage of tile k
  uint32_t parity = 0;  // parity of the current use of stage s
  for (int k = 0; k < total; ++k)
  {
    const uint32_t tile = tiles + (uint32_t)s * kTileBytes;
    if (P::HAS_IN)
    {
      mbar_wait(bars + 8u * s, parity);
    }
    else
    {
      // output-only ring: the store that last read this stage (tile k-S) must be done
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
    }

    // ---- 32 samples of this lane's voice, in place.  All eight LDS.128 are issued first so
    // their latency overlaps the arithmetic (the asm volatile accessors keep program order) ----
    float4 xin[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      xin[j] = P::HAS_IN ? lds128(tile + row_off + (((uint32_t)j << 4) ^ sw))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      float4 y;
      y.x = P::tick(xin[j].x, st, co);
      y.y = P::tick(xin[j].y, st, co);
      y.z = P::tick(xin[j].z, st, co);
      y.w = P::tick(xin[j].w, st, co);
      sts128(tile + row_off + (((uint32_t)j << 4) ^ sw), y);
    }

    const int t = k >> 1, h = k & 1;
    if (a.mix_partial != nullptr)
    {
      // lane n sums sample column n over the 32 voice rows, rows in voice order, from +0.
      // Element (r, n) sits at r*128 + (((n>>2) ^ (r&7)) << 4) + (n&3)*4: mix_off[r&7] holds the
      // lane-dependent part, the rest folds into LDS immediates.
      __syncwarp();
      float acc = 0.0f;
      if (full_group)
      {
#pragma unroll
        for (int r = 0; r < 32; ++r)
          acc = __fadd_rn(acc, lds32(tile + mix_off[r & 7] + (uint32_t)r * 128u));
      }
      else
      {
        // ragged last group: rows beyond V hold garbage computed from zero-filled input
        for (int r = 0; r < 32; ++r)
          if (v0 + r < a.V) acc = __fadd_rn(acc, lds32(tile + mix_off[r & 7] + (uint32_t)r * 128u));
      }
      mix_row[h * kTileSamples] = acc;
      if (h) mix_row += mix_block_stride;
    }

    // make this lane's generic-proxy writes visible to the TMA unit, then hand over
    fence_proxy_async();
    __syncwarp();
    if (lane == 0)
    {
      if (a.write_out)
      {
        tma_store_3d(&out_map, tile, h * kTileSamples, v0, t * a.n_out_planes + a.out_plane);
        bulk_commit();
      }
      if (P::HAS_IN)
      {
        const int kn = k + S - 1;  // next tile to fetch, into the stage tile k-1 used
        if (kn < total)
        {
          if (a.write_out && k >= 1) bulk_wait_read<1>();  // store k-1 finished reading smem
          const int sn = (s == 0) ? (S - 1) : (s - 1);
          const uint32_t bar = bars + 8u * sn;
          mbar_arrive_expect_tx(bar, kTileBytes);
          tma_load_3d(tiles + (uint32_t)sn * kTileBytes, &in_map, bar, (kn & 1) * kTileSamples, v0,
                      (kn >> 1) * a.n_in_planes + a.in_plane, kEvictFirst);
        }
      }
    }
    if (++s == S)
    {
      s = 0;
      parity ^= 1u;
    }
  }

  // ---- state back to HBM; shared memory must outlive the last bulk stores ----
#pragma unroll
  for (int i = 0; i < P::NS; ++i)
    if (live) a.state[(size_t)a.st_idx[i] * a.V + v] = st[i];
  if (lane == 0) bulk_wait_read<0>();
  __syncwarp();
}

// second stage of the mix bus.  Deterministic two-level sum (DESIGN.md "mix bus"):
//   chunk[c][n] = sum over the (up to) 64 group partials of chunk c, in group order, from +0
//   mix[p][n]   = sum over chunks c = 0..C-1, in order, from +0
// One CTA per plane, 64 x 16 threads: thread (n, w) owns chunks w, w+16, ...
constexpr int kMixChunkGroups = 64;  // 64 groups x 32 voices = 2048 voices per chunk
__global__ void __launch_bounds__(1024) mix_reduce_kernel(const float* __restrict__ partial,
                                                          float* __restrict__ chunk_scratch,
                                                          float* __restrict__ mix, int n_groups)
{
  const int p = blockIdx.x, n = threadIdx.x, w = threadIdx.y;
  const int n_chunks = (n_groups + kMixChunkGroups - 1) / kMixChunkGroups;
  const float* src = partial + (size_t)p * n_groups * MLB_BLOCK + n;
  float* scratch = chunk_scratch + (size_t)p * n_chunks * MLB_BLOCK + n;
  for (int c = w; c < n_chunks; c += blockDim.y)
  {
    const int g0 = c * kMixChunkGroups;
    const int g1 = min(g0 + kMixChunkGroups, n_groups);
    float acc = 0.0f;
    int g = g0;
    for (; g + 8 <= g1; g += 8)
    {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = src[(size_t)(g + i) * MLB_BLOCK];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = __fadd_rn(acc, x[i]);
    }
    for (; g < g1; ++g) acc = __fadd_rn(acc, src[(size_t)g * MLB_BLOCK]);
    scratch[(size_t)c * MLB_BLOCK] = acc;
  }
  __syncthreads();  // also orders the global scratch writes within the CTA
  if (w == 0)
  {
    float acc = 0.0f;
    for (int c = 0; c < n_chunks; ++c) acc = __fadd_rn(acc, scratch[(size_t)c * MLB_BLOCK]);
    mix[(size_t)p * MLB_BLOCK + n] = acc;
  }
}

}  // namespace mlb
