// fdn_kernel.cuh -- fused [3-operator FM ->] FDN<8> kernel (config 4; reference
// MLDSPFilters.h:1162-1239 FDN<SIZE>, :801-914 IntegerDelay, MLDSPGens.h:373-381 SineGen).
//
// Decomposition: one LANE per (voice, delay line): 8 lanes per voice, 4 voices per warp, warps
// independent.  Delay rings live in HBM as [voice][line][R] f32 (R = 2^k >= maxlen + 64), the
// genuinely bandwidth-heavy part of the path: 8 x (256 B read + 256 B write) per voice-block
// = 64 B/voice-sample of the 76 B/voice-sample total.  Per 64-sample block a warp
//   0. waits for its lanes' ring reads (cp.async.bulk 1-D, 16-byte-aligned 272 B over-read of
//      the 64-sample window that starts `len` samples behind the write position) and the
//      voices' input rows;
//   G. generator: the 8 lanes of a voice split the 64 samples (8 each); PhasorGen's u32 phase
//      accumulation is an integer prefix sum, so it is done exactly with a per-lane prefix and
//      a 3-step shuffle scan; three sines (two modulators, carrier) -> x row in shared memory;
//   B. Householder mixing: lanes re-distribute (lane k takes samples k, k+8, ...) and read all
//      8 lines of their samples from shared memory, so the sums  sumOfDelays, sumL, sumR are
//      formed in the reference's left-to-right order in registers, no shuffles;
//      v = d - (2/8) * sum is written back in place, sumL / sumR rows are built;
//   C. per line again: OnePole recurrence, * feedback gain, + x  -> next block's delay input,
//      written in place (LDS.128 / STS.128, bank-conflict free with the 68-float row stride);
//   S. each lane bulk-stores its 256 B vector into its ring at w + 64, two lanes store the
//      L / R rows; after the stores completed the next block's ring reads are issued (short
//      delays read what was just written).
// The reference's mDelayInputVectors (the vector a block hands to the next one) is never
// materialised separately: it is exactly what the next block writes into the ring at its write
// position, so it is written there directly at the end of the block that produces it.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ops.cuh"
#include "tma.cuh"

namespace mlb
{
struct FdnArgs
{
  int gen;                                  // 0: x = input plane, 1: 3-operator FM of the input plane
  int st_ph1, st_ph2, st_phc, st_y1;        // SoA state word indices (y1: first of 8)
  int co_r1, co_r2, co_i1, co_i2, co_one;   // PARAM coef word indices
  int co_fdn;                               // first of the 32 FDN8 coef words
  int in_plane;
};

struct FdnLaunch
{
  FdnArgs f;
  uint32_t* state;
  const float* coef;
  float* ring;
  const float* in;  // [T][n_in][V][64]
  float* out;       // [T][2][V][64] (L plane, R plane)
  int ring_len;
  long long blocks_done;
  int V, T, n_in;
};

constexpr int kFdnRow = 68;                      // floats per delay row: 16-B aligned, = 4 mod 32
constexpr int kFdnVoice = 8 * kFdnRow + 2 * 64 + 8;  // 8 delay rows + 2 input/x rows (+pad: = 8 mod 32)
constexpr int kFdnMaxWarpsPerCta = 4;  // warps per CTA (MLB_FDN_WARPS picks fewer for experiments)

template <bool EX>
__global__ void __launch_bounds__(128, 4) fdn8_kernel(const FdnLaunch a)
{
  extern __shared__ __align__(16) float fdn_smem[];
  using ar = A<EX>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;         // warps per CTA (run-time: chosen by the host per bank size)
  const int vq = lane >> 3, k = lane & 7;  // voice within the warp, delay line / sample slot
  const int v = (blockIdx.x * wpc + warp) * 4 + vq;
  if ((blockIdx.x * wpc + warp) * 4 >= a.V) return;  // warp-uniform
  const bool live = v < a.V;

  float* vbase = fdn_smem + (size_t)(warp * 4 + vq) * kFdnVoice;
  // rows: 8 delay rows (this lane's: k), then two input rows: block t's freq row lands in
  // xbuf[t & 1]; the generator turns it into the x row in place
  const uint32_t s_vbase = smem_u32(vbase), s_drow = s_vbase + (uint32_t)(k * kFdnRow) * 4u;
  const uint32_t s_xbuf0 = s_vbase + (uint32_t)(8 * kFdnRow) * 4u;
  const uint32_t bars = smem_u32(fdn_smem + (size_t)wpc * 4 * kFdnVoice) + warp * 24u;
  const uint32_t ringbar = bars, xbar0 = bars + 8u;  // xbar[b] = xbar0 + 8 b

  if (lane == 0)
  {
    mbar_init(ringbar, 32);
    mbar_init(xbar0, 4);
    mbar_init(xbar0 + 8u, 4);
    fence_mbar_init();
  }
  __syncwarp();

  // ---- per-lane constants and state ----
  const size_t V = (size_t)a.V;
  const int vv = live ? v : 0;
  const float a0 = a.coef[(size_t)(a.f.co_fdn + k) * V + vv];
  const float b1 = a.coef[(size_t)(a.f.co_fdn + 8 + k) * V + vv];
  const float gn = a.coef[(size_t)(a.f.co_fdn + 16 + k) * V + vv];
  const uint32_t len = (uint32_t)(int)a.coef[(size_t)(a.f.co_fdn + 24 + k) * V + vv];
  float y1 = u2f(a.state[(size_t)(a.f.st_y1 + k) * V + vv]);
  float r1 = 0, r2 = 0, i1 = 0, i2 = 0, one = 0;
  uint32_t ph1 = 0, ph2 = 0, phc = 0;
  if (a.f.gen == 1)
  {
    r1 = a.coef[(size_t)a.f.co_r1 * V + vv], r2 = a.coef[(size_t)a.f.co_r2 * V + vv];
    i1 = a.coef[(size_t)a.f.co_i1 * V + vv], i2 = a.coef[(size_t)a.f.co_i2 * V + vv];
    one = a.coef[(size_t)a.f.co_one * V + vv];
    ph1 = a.state[(size_t)a.f.st_ph1 * V + vv], ph2 = a.state[(size_t)a.f.st_ph2 * V + vv];
    phc = a.state[(size_t)a.f.st_phc * V + vv];
  }
  const uint32_t mask = (uint32_t)a.ring_len - 1u;
  float* ring = a.ring + ((size_t)vv * 8 + k) * a.ring_len;
  // a delay shorter than a block reads what the previous block has just written
  const bool warp_has_short = __any_sync(0xffffffffu, live && len < (uint32_t)MLB_BLOCK);

  // this lane's ring read for block t: the 64-sample window that starts len behind w_t,
  // over-read to 16-byte alignment (68 floats), split in two when it wraps
  auto issue_ring_load = [&](int t) -> uint32_t
  {
    const uint32_t w = (uint32_t)(((a.blocks_done + t) * MLB_BLOCK) & (long long)mask);
    const uint32_t r = (w - len) & mask;
    const uint32_t al = r & ~3u;
    if (live)
    {
      const uint32_t first = min((uint32_t)kFdnRow, (uint32_t)a.ring_len - al);  // floats before the wrap
      mbar_arrive_expect_tx(ringbar, kFdnRow * 4);
      bulk_load_1d(s_drow, ring + al, first * 4, ringbar);
      if (first < (uint32_t)kFdnRow) bulk_load_1d(s_drow + first * 4, ring, (kFdnRow - first) * 4, ringbar);
    }
    else
    {
      mbar_arrive_expect_tx(ringbar, 0);
    }
    return r & 3u;  // offset of the window inside the over-read row
  };
  // input row of block t (one lane per voice)
  auto issue_x_load = [&](int t)
  {
    if (k != 0) return;
    const uint32_t bar = xbar0 + 8u * (uint32_t)(t & 1);
    if (live)
    {
      mbar_arrive_expect_tx(bar, MLB_BLOCK * 4);
      bulk_load_1d(s_xbuf0 + (uint32_t)(t & 1) * 256u,
                   a.in + (((size_t)t * a.n_in + a.f.in_plane) * V + v) * MLB_BLOCK, MLB_BLOCK * 4, bar);
    }
    else
    {
      mbar_arrive_expect_tx(bar, 0);
    }
  };

  // ---- G: the three sines of block t, samples 8k .. 8k+7 of this voice, in place in xbuf[t&1] ----
  auto generate = [&](int t)
  {
    mbar_wait(xbar0 + 8u * (uint32_t)(t & 1), (uint32_t)((t >> 1) & 1));
    if (a.f.gen != 1) return;  // external input: the row already is x
    const uint32_t s_x = s_xbuf0 + (uint32_t)(t & 1) * 256u + (uint32_t)k * 32u;
    float f[8];
    {
      const float4 q0 = lds128(s_x), q1 = lds128(s_x + 16u);
      f[0] = q0.x, f[1] = q0.y, f[2] = q0.z, f[3] = q0.w, f[4] = q1.x, f[5] = q1.y, f[6] = q1.z, f[7] = q1.w;
    }
    // inclusive integer prefix of the per-sample phase increments, then across the 8 lanes
    auto scan_phases = [&](const float (&freq)[8], uint32_t& ph0, uint32_t (&phase)[8])
    {
      uint32_t run = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
      {
        run += (uint32_t)cvt_round(ar::mul(freq[j], 4294967296.0f));
        phase[j] = run;
      }
      uint32_t sc = run;
#pragma unroll
      for (int d = 1; d < 8; d <<= 1)
      {
        const uint32_t up = __shfl_up_sync(0xffffffffu, sc, d, 8);
        if (k >= d) sc += up;
      }
      const uint32_t excl = sc - run + ph0;
#pragma unroll
      for (int j = 0; j < 8; ++j) phase[j] += excl;
      ph0 += __shfl_sync(0xffffffffu, sc, 7, 8);
    };
    float fa[8], fb[8];
    uint32_t pa[8], pb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] = ar::mul(f[j], r1), fb[j] = ar::mul(f[j], r2);
    scan_phases(fa, ph1, pa);
    scan_phases(fb, ph2, pb);
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      const float m1 = phase_to_sine<EX>(pa[j]), m2 = phase_to_sine<EX>(pb[j]);
      const float bsum = ar::add(ar::add(one, ar::mul(m1, i1)), ar::mul(m2, i2));
      fa[j] = ar::mul(f[j], bsum);
    }
    scan_phases(fa, phc, pa);
    float4 x0, x1;
    x0.x = phase_to_sine<EX>(pa[0]), x0.y = phase_to_sine<EX>(pa[1]);
    x0.z = phase_to_sine<EX>(pa[2]), x0.w = phase_to_sine<EX>(pa[3]);
    x1.x = phase_to_sine<EX>(pa[4]), x1.y = phase_to_sine<EX>(pa[5]);
    x1.z = phase_to_sine<EX>(pa[6]), x1.w = phase_to_sine<EX>(pa[7]);
    sts128(s_x, x0);
    sts128(s_x + 16u, x1);
  };

  // ---- prologue: ring(0), x(0), x(1) in flight; generate x(0) while the ring read lands ----
  uint32_t off = issue_ring_load(0);
  issue_x_load(0);
  if (a.T > 1) issue_x_load(1);
  generate(0);
  __syncwarp();

  uint32_t rparity = 0;
  for (int t = 0; t < a.T; ++t)
  {
    const uint32_t s_xrow = s_xbuf0 + (uint32_t)(t & 1) * 256u;
    mbar_wait(ringbar, rparity);
    rparity ^= 1u;

    // ---- B: Householder mixing, lane k owns samples k, k+8, ..., k+56 of its voice ----
    {
      uint32_t offs[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) offs[n] = __shfl_sync(0xffffffffu, off, n, 8);
      float d[8][8];
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d[n][j] = lds32(s_vbase + (uint32_t)(n * kFdnRow + k + 8 * j) * 4u + offs[n] * 4u);
      __syncwarp();  // every lane holds its samples before rows are overwritten in place
      float* outL = a.out + (((size_t)t * 2 + 0) * V + vv) * MLB_BLOCK + k;
      float* outR = a.out + (((size_t)t * 2 + 1) * V + vv) * MLB_BLOCK + k;
#pragma unroll
      for (int j = 0; j < 8; ++j)
      {
        // DSPVector sumR, sumL, sumOfDelays start from zero and add lines left to right (F:1204-1227)
        float sumR = 0.f, sumL = 0.f, sum = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n)
        {
          if (n & 1)
            sumL = __fadd_rn(sumL, d[n][j]);
          else
            sumR = __fadd_rn(sumR, d[n][j]);
          sum = __fadd_rn(sum, d[n][j]);
        }
        sum = __fmul_rn(sum, 0.25f);  // 2 / SIZE
#pragma unroll
        for (int n = 0; n < 8; ++n)
          sts32(s_vbase + (uint32_t)(n * kFdnRow + k + 8 * j) * 4u, __fsub_rn(d[n][j], sum));
        if (live)
        {
          // concatRows(sumL, sumR) (F:1237): the 8 lanes of a voice write one full 32-B sector
          __stcs(outL + 8 * j, sumL);
          __stcs(outR + 8 * j, sumR);
        }
      }
    }
    __syncwarp();

    // ---- C: per line: OnePole, feedback gain, + x -> next block's delay input, in place ----
#pragma unroll 4
    for (int q = 0; q < 16; ++q)
    {
      const float4 vq4 = lds128(s_drow + (uint32_t)q * 16u);
      const float4 xq4 = lds128(s_xrow + (uint32_t)q * 16u);
      float4 u;
      y1 = ar::mul_add_mul(a0, vq4.x, b1, y1);
      u.x = ar::add(ar::mul(y1, gn), xq4.x);
      y1 = ar::mul_add_mul(a0, vq4.y, b1, y1);
      u.y = ar::add(ar::mul(y1, gn), xq4.y);
      y1 = ar::mul_add_mul(a0, vq4.z, b1, y1);
      u.z = ar::add(ar::mul(y1, gn), xq4.z);
      y1 = ar::mul_add_mul(a0, vq4.w, b1, y1);
      u.w = ar::add(ar::mul(y1, gn), xq4.w);
      sts128(s_drow + (uint32_t)q * 16u, u);
    }

    // ---- S: next block's delay input goes straight to the ring position it will be read from ----
    fence_proxy_async();
    __syncwarp();
    if (live)
    {
      const uint32_t wn = (uint32_t)(((a.blocks_done + t + 1) * MLB_BLOCK) & (long long)mask);
      bulk_store_1d(ring + wn, s_drow, MLB_BLOCK * 4);
      bulk_commit();
    }
    if (t + 1 < a.T)
    {
      // x(t) is consumed: its buffer takes the input row of block t+2
      if (t + 2 < a.T) issue_x_load(t + 2);
      if (!warp_has_short)
      {
        // long delays only need the row buffer back: start the next ring read at once and hide
        // its latency behind the generator of block t+1
        if (live) bulk_wait_read<0>();
        __syncwarp();
        off = issue_ring_load(t + 1);
        generate(t + 1);
      }
      else
      {
        // some line of this warp reads what was just written: generate first (hides the store
        // latency), then wait until the stores are globally performed
        generate(t + 1);
        if (live) bulk_wait_all<0>();
        __syncwarp();
        off = issue_ring_load(t + 1);
      }
      __syncwarp();
    }
  }

  if (live)
  {
    bulk_wait_all<0>();
    a.state[(size_t)(a.f.st_y1 + k) * V + v] = f2u(y1);
    if (a.f.gen == 1 && k == 0)
    {
      a.state[(size_t)a.f.st_ph1 * V + v] = ph1;
      a.state[(size_t)a.f.st_ph2 * V + v] = ph2;
      a.state[(size_t)a.f.st_phc * V + v] = phc;
    }
  }
}

// Mix-bus partials for kernels that do not produce them in their epilogue:
// partial[p][g][n] = sum over the 32 voices of group g (voice order, from +0) of plane[p][v][n].
__global__ void __launch_bounds__(64) mix_partial_from_planes_kernel(const float* __restrict__ planes,
                                                                      float* __restrict__ partial, int V,
                                                                      int n_groups)
{
  const int g = blockIdx.x, p = blockIdx.y, n = threadIdx.x;
  const float* src = planes + ((size_t)p * V + (size_t)g * 32) * MLB_BLOCK + n;
  const int cnt = min(32, V - g * 32);
  float acc = 0.f;
  for (int r = 0; r < cnt; ++r) acc = __fadd_rn(acc, src[(size_t)r * MLB_BLOCK]);
  partial[((size_t)p * n_groups + g) * MLB_BLOCK + n] = acc;
}

// ---- host side: recognise INPUT [-> 3-op FM] -> FDN8 (+ FDN8_R) ----
inline bool match_fm3_fdn8(const std::vector<mlb_node>& N, const std::vector<int32_t>& outs,
                           const std::vector<int32_t>& st_off, const std::vector<int32_t>& co_off,
                           FdnArgs* f)
{
  if (outs.size() != 2) return false;
  const int nl = outs[0], nr = outs[1];
  if (N[nl].op != MLB_OP_FDN8 || N[nr].op != MLB_OP_FDN8_R || N[nr].in[0] != nl) return false;
  std::vector<char> used(N.size(), 0);
  used[nl] = used[nr] = 1;
  f->st_y1 = st_off[nl];
  f->co_fdn = co_off[nl];
  const int x = N[nl].in[0];
  used[x] = 1;
  auto all_used = [&]()
  {
    for (char u : used)
      if (!u) return false;
    return true;
  };
  if (N[x].op == MLB_OP_INPUT)
  {
    f->gen = 0;
    f->in_plane = N[x].iarg;
    return all_used();
  }
  // carrier = SINE(MULTIPLY(f, ADD(ADD(one, MULTIPLY(m1, i1)), MULTIPLY(m2, i2))))
  auto is = [&](int n, int op) { return n >= 0 && N[n].op == op; };
  // split a commutative binary node into (the operand with op `want`, the other one)
  auto split = [&](int node, int want, int* a, int* b) -> bool
  {
    const int p = N[node].in[0], q = N[node].in[1];
    if (is(p, want) && !is(q, want)) { *a = p, *b = q; return true; }
    if (is(q, want) && !is(p, want)) { *a = q, *b = p; return true; }
    return false;
  };
  if (!is(x, MLB_OP_SINE)) return false;
  const int mulc = N[x].in[0];
  if (!is(mulc, MLB_OP_MULTIPLY)) return false;
  used[mulc] = 1;
  int fin, bsum;
  if (!split(mulc, MLB_OP_INPUT, &fin, &bsum)) return false;
  if (!is(bsum, MLB_OP_ADD)) return false;
  used[fin] = used[bsum] = 1;
  // bsum = ADD(a1, mul2) with a1 = ADD(one, mul1): the reference order (one + m1*i1) + m2*i2
  const int a1 = N[bsum].in[0], mul2 = N[bsum].in[1];
  if (!is(a1, MLB_OP_ADD) || !is(mul2, MLB_OP_MULTIPLY)) return false;
  const int onep = N[a1].in[0], mul1 = N[a1].in[1];
  if (!is(onep, MLB_OP_PARAM) || !is(mul1, MLB_OP_MULTIPLY)) return false;
  used[a1] = used[mul2] = used[onep] = used[mul1] = 1;
  int s1, p1, s2, p2;
  if (!split(mul1, MLB_OP_SINE, &s1, &p1) || !is(p1, MLB_OP_PARAM)) return false;
  if (!split(mul2, MLB_OP_SINE, &s2, &p2) || !is(p2, MLB_OP_PARAM)) return false;
  used[s1] = used[p1] = used[s2] = used[p2] = 1;
  const int m1 = N[s1].in[0], m2 = N[s2].in[0];
  if (!is(m1, MLB_OP_MULTIPLY) || !is(m2, MLB_OP_MULTIPLY)) return false;
  used[m1] = used[m2] = 1;
  int f1, q1, f2, q2;
  if (!split(m1, MLB_OP_INPUT, &f1, &q1) || !is(q1, MLB_OP_PARAM) || f1 != fin) return false;
  if (!split(m2, MLB_OP_INPUT, &f2, &q2) || !is(q2, MLB_OP_PARAM) || f2 != fin) return false;
  used[q1] = used[q2] = 1;
  if (!all_used()) return false;
  f->gen = 1;
  f->in_plane = N[fin].iarg;
  f->st_ph1 = st_off[s1], f->st_ph2 = st_off[s2], f->st_phc = st_off[x];
  f->co_r1 = co_off[q1], f->co_r2 = co_off[q2], f->co_i1 = co_off[p1], f->co_i2 = co_off[p2];
  f->co_one = co_off[onep];
  return true;
}

inline int launch_fm3_fdn8(const FdnArgs& f, bool exact, uint32_t* state, const float* coef, float* ring,
                           float* /*carry*/, int ring_len, long long blocks_done, const float* in, float* out,
                           int V, int T, int n_in, cudaStream_t stream)
{
  FdnLaunch a;
  a.f = f;
  a.state = state, a.coef = coef, a.ring = ring, a.in = in, a.out = out;
  a.ring_len = ring_len, a.blocks_done = blocks_done;
  a.V = V, a.T = T, a.n_in = n_in;
  // Warps per CTA.  Config 4 (16 384 voices = 1 024 four-warp CTAs over 4 x 148 slots) runs 1.73 waves, but fitting
  // whole waves does not pay: measured at 16 384 voices x 16 blocks (MLB_FDN_WARPS, round 2) 4 warps per CTA
  // (16 resident warps per SM) 0.255 ms, 5: 0.282, 6: 0.352, 7 (14 resident warps, two even waves): 0.281,
  // 8: 0.279 -- nor does a fifth CTA per SM at 96 registers (0.276 ms): the launch shape is not the lever; the
  // serialised per-lane bulk copies and the ring read that starts only after the previous block's store are.
  const void* fn = exact ? (const void*)fdn8_kernel<true> : (const void*)fdn8_kernel<false>;
  const int units = (V + 3) / 4;
  int best_wpc = 4;
  static bool attr_set[2] = {false, false};
  if (!attr_set[exact ? 1 : 0])
  {
    const size_t smem_max = (size_t)kFdnMaxWarpsPerCta * 4 * kFdnVoice * 4 + kFdnMaxWarpsPerCta * 24;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess) return MLB_ERR_CUDA;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    attr_set[exact ? 1 : 0] = true;
  }
  const char* force = getenv("MLB_FDN_WARPS");
  if (force && *force) best_wpc = std::min(std::max(atoi(force), 1), kFdnMaxWarpsPerCta);
  const int wpc = best_wpc;
  const int n_ctas = (units + wpc - 1) / wpc;
  const size_t smem = (size_t)wpc * 4 * kFdnVoice * 4 + (size_t)wpc * 24;
  if (exact)
    fdn8_kernel<true><<<n_ctas, wpc * 32, smem, stream>>>(a);
  else
    fdn8_kernel<false><<<n_ctas, wpc * 32, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? MLB_OK : MLB_ERR_CUDA;
}
}  // namespace mlb
