// generic_kernel.cuh -- graph interpreter: evaluates ANY voice graph (the "procs" launcher,
// reference stub source/procs/MLProcMultiply.cpp:29-46) in one kernel launch per call, for
// graphs that have no fused specialisation.  Lane per voice, warp per 32 voices; the rows a
// node produces live in shared memory as [slot][lane][68 floats] (per-lane rows walked with
// LDS.128 / STS.128, bank-conflict free); stateful nodes keep their state in registers across
// the 64-sample loop and read/write the SoA state once per block.  The op switch is hoisted
// out of the sample loop for every node kind (one specialised row loop per op); node
// descriptors are read through the read-only path.  Results are identical to the fused kernels (same device functions).
#pragma once
#include "ops.cuh"
#include "tma.cuh"
#include "functors.cuh"

namespace mlb
{
// rows live in shared memory as [slot][lane][68 floats]: each lane walks its own row with
// LDS.128 / STS.128 (row stride 68 words = 4 mod 32 -> the 8 lanes of a quarter-warp cover all
// 32 banks), and the mix-bus column walk ([voice rr][sample = lane]) is conflict free too
constexpr uint32_t kRowFloats = 68u;
constexpr uint32_t kRowStride = kRowFloats * 4u;      // bytes between lanes
constexpr uint32_t kSlotBytes = 32u * kRowStride;     // one row slot for a 32-voice group

enum
{
  OPERAND_NONE = 0,
  OPERAND_SLOT = 1,   // row in shared memory
  OPERAND_PARAM = 2   // per-voice scalar in the coef SoA (DSPVector(float) broadcast)
};

struct GNode
{
  int op;
  int in_kind[3];
  int in_ref[3];   // slot index or coef word index
  int out_slot;    // -1: node output never read as a row (PARAM)
  int out_slot2;   // FDN8: slot of the sumR row (read by FDN8_R), else -1
  int st_off, co_off;
  int out_plane;   // >= 0: also written to out / mix plane
  int iarg;
  // delay memory of this node (float offsets into GenericArgs::dmem)
  unsigned ring_stride;          // floats between the rings of consecutive voices (power of two)
  unsigned long long row_off;    // [V][64] member row (Allpass::vy1, LinearGlide::mCurrVec, feedback)
  unsigned long long ring_off;   // [V][ring_stride] IntegerDelay ring
};

struct GenericArgs
{
  const GNode* nodes;
  int n_nodes;
  uint32_t* state;
  const float* coef;
  const float* in;     // [T][n_in][V][64]
  float* out;          // [T][n_out][V][64] or nullptr
  float* mix_partial;  // [T*n_out][n_groups][64] or nullptr
  int V, T, n_in, n_out, n_groups, n_slots;
  // FDN delay memory (one FDN8 node per graph supported): rings [V][8][ring], carry [V][8][64]
  float* fdn_ring;
  float* fdn_carry;
  int fdn_ring_len;        // power of two
  long long blocks_done;   // IntegerDelay write index = (64 * blocks_done) & (ring_len - 1)
  float* dmem;             // delay memory of the section-8(f) functors (rows + rings)
  int scratch_slot;        // first of 4 scratch row slots (delay input, two tap streams, old block)
};

struct RowRef
{
  uint32_t addr;  // shared address of this lane's row
  float k;
  bool is_row;
  MLB_DEV float get(int n) const { return is_row ? lds32(addr + (uint32_t)n * 4u) : k; }
  MLB_DEV float4 get4(int q) const
  {
    return is_row ? lds128(addr + (uint32_t)q * 16u) : make_float4(k, k, k, k);
  }
};

template <int OP, bool EX>
MLB_DEV void run_filter_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x,
                             uint32_t out_addr)
{
  constexpr int NS = op_ns(OP), NC = op_nc(OP);
  uint32_t st[NS > 0 ? NS : 1];
  float co[NC > 0 ? NC : 1];
#pragma unroll
  for (int i = 0; i < NS; ++i) st[i] = live ? a.state[(size_t)(nd.st_off + i) * a.V + v] : 0u;
#pragma unroll
  for (int i = 0; i < NC; ++i) co[i] = live ? a.coef[(size_t)(nd.co_off + i) * a.V + v] : 0.f;
#pragma unroll 2
  for (int q = 0; q < 16; ++q)
  {
    const float4 xi = x.get4(q);
    float4 y;
    y.x = filter_tick<EX>(OP, xi.x, st, co);
    y.y = filter_tick<EX>(OP, xi.y, st, co);
    y.z = filter_tick<EX>(OP, xi.z, st, co);
    y.w = filter_tick<EX>(OP, xi.w, st, co);
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
#pragma unroll
  for (int i = 0; i < NS; ++i)
    if (live) a.state[(size_t)(nd.st_off + i) * a.V + v] = st[i];
}

template <int OP, bool EX>
MLB_DEV void run_gen_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef f,
                          RowRef w, uint32_t out_addr)
{
  uint32_t st[1];
  st[0] = live ? a.state[(size_t)nd.st_off * a.V + v] : 0u;
#pragma unroll 2
  for (int q = 0; q < 16; ++q)
  {
    const float4 fi = (OP == MLB_OP_NOISE) ? make_float4(0.f, 0.f, 0.f, 0.f) : f.get4(q);
    const float4 wi = (OP == MLB_OP_PULSE) ? w.get4(q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 y;
    y.x = gen_tick<EX>(OP, fi.x, wi.x, st);
    y.y = gen_tick<EX>(OP, fi.y, wi.y, st);
    y.z = gen_tick<EX>(OP, fi.z, wi.z, st);
    y.w = gen_tick<EX>(OP, fi.w, wi.w, st);
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
  if (live) a.state[(size_t)nd.st_off * a.V + v] = st[0];
}

// stateless elementwise node with the op known at compile time (the switch in op_apply folds)
template <int OP, bool EX>
MLB_DEV void run_stateless_node(RowRef x, RowRef b, RowRef c, uint32_t out_addr)
{
  constexpr int NIN = op_nin(OP);
#pragma unroll 4
  for (int q = 0; q < 16; ++q)
  {
    const float4 xi = x.get4(q);
    const float4 bi = NIN >= 2 ? b.get4(q) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 ci = NIN >= 3 ? c.get4(q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 y;
    y.x = op_apply<EX>(OP, xi.x, bi.x, ci.x);
    y.y = op_apply<EX>(OP, xi.y, bi.y, ci.y);
    y.z = op_apply<EX>(OP, xi.z, bi.z, ci.z);
    y.w = op_apply<EX>(OP, xi.w, bi.w, ci.w);
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
}

// every stateless op gets its own specialised row loop
template <bool EX>
MLB_DEV void dispatch_stateless(int op, RowRef x, RowRef b, RowRef c, uint32_t out_addr)
{
  switch (op)
  {
#define MLB_X_STATELESS(NAME, id, nin, nst, nco)                           \
  case id:                                                                 \
    if constexpr (id >= MLB_OP_MAP_FIRST && id < MLB_OP_MAP_END)         \
      run_stateless_node<id, EX>(x, b, c, out_addr);                       \
    break;
    MLB_OP_TABLE_STATELESS(MLB_X_STATELESS)
#undef MLB_X_STATELESS
    default: break;
  }
}


// ---- SURVEY 8(f) row 2 node runners (device functions in functors.cuh) ----

template <int OP, bool EX>
MLB_DEV float functor_tick(float x, uint32_t* st, const float* co)
{
  if constexpr (OP == MLB_OP_ONESHOT) return oneshot_tick<EX>(x, st);
  if constexpr (OP == MLB_OP_PEAK) return peak_tick<EX>(x, st, co);
  if constexpr (OP == MLB_OP_RMS) return rms_tick<EX>(x, st, co);
  if constexpr (OP == MLB_OP_ADSR) return adsr_tick<EX>(x, st, co);
  if constexpr (OP == MLB_OP_SAMPLE_GLIDE) return sample_glide_tick<EX>(x, st, co);
  if constexpr (OP == MLB_OP_ALLPASS1)
  {
    float x1 = u2f(st[0]), y1 = u2f(st[1]);
    const float y = allpass1_tick<EX>(x, x1, y1, co[0]);
    st[0] = f2u(x1), st[1] = f2u(y1);
    return y;
  }
  return x;
}

// register-state functor with one signal input
template <int OP, bool EX>
MLB_DEV void run_functor_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x,
                              uint32_t out_addr)
{
  constexpr int NS = op_ns(OP), NC = op_nc(OP);
  uint32_t st[NS > 0 ? NS : 1];
  float co[NC > 0 ? NC : 1];
#pragma unroll
  for (int i = 0; i < NS; ++i) st[i] = live ? a.state[(size_t)(nd.st_off + i) * a.V + v] : 0u;
#pragma unroll
  for (int i = 0; i < NC; ++i) co[i] = live ? a.coef[(size_t)(nd.co_off + i) * a.V + v] : 0.f;
#pragma unroll 1
  for (int q = 0; q < 16; ++q)
  {
    const float4 xi = x.get4(q);
    float4 y;
    y.x = functor_tick<OP, EX>(xi.x, st, co);
    y.y = functor_tick<OP, EX>(xi.y, st, co);
    y.z = functor_tick<OP, EX>(xi.z, st, co);
    y.w = functor_tick<OP, EX>(xi.w, st, co);
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
  if constexpr (OP == MLB_OP_PEAK)  // F:607-610
    if ((int32_t)st[1] > 0) st[1] = (uint32_t)((int32_t)st[1] - MLB_BLOCK);
#pragma unroll
  for (int i = 0; i < NS; ++i)
    if (live) a.state[(size_t)(nd.st_off + i) * a.V + v] = st[i];
}

MLB_DEV float* node_row(const GNode& nd, const GenericArgs& a, int v)
{
  return a.dmem + nd.row_off + (size_t)v * MLB_BLOCK;
}
MLB_DEV void row_global_to_smem(const float* src, uint32_t dst)
{
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 4
  for (int q = 0; q < 16; ++q) sts128(dst + (uint32_t)q * 16u, s4[q]);
}
MLB_DEV void row_smem_to_global(RowRef x, float* dst)
{
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
  for (int q = 0; q < 16; ++q) d4[q] = x.get4(q);
}

// LinearGlide::operator()(float), G:459-505.  The scalar argument is sample 0 of the operand.
template <bool EX>
MLB_DEV void run_glide_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x,
                            uint32_t out_addr)
{
  using ar = A<EX>;
  if (!live) return;
  float* curr = node_row(nd, a, v);
  float step = u2f(a.state[(size_t)nd.st_off * a.V + v]);
  float target = u2f(a.state[(size_t)(nd.st_off + 1) * a.V + v]);
  int32_t remaining = (int32_t)a.state[(size_t)(nd.st_off + 2) * a.V + v];
  const int32_t per = cvt_trunc(a.coef[(size_t)nd.co_off * a.V + v]);
  const float dy = a.coef[(size_t)(nd.co_off + 1) * a.V + v];
  const float f = x.get(0);
  if (f != target)
  {
    target = f;
    remaining = per;
  }
  if (remaining < 0)
    row_global_to_smem(curr, out_addr);
  else
  {
    float cv = 0.f;
    const int mode = (remaining == 0) ? 0 : (remaining == per) ? 1 : 2;
    if (mode == 0) step = 0.f;
    if (mode == 1)
    {
      cv = curr[MLB_BLOCK - 1];
      step = ar::mul(ar::sub(target, cv), dy);
    }
    float4* c4 = reinterpret_cast<float4*>(curr);
#pragma unroll 2
    for (int q = 0; q < 16; ++q)
    {
      float4 y;
      if (mode == 0)
        y = make_float4(target, target, target, target);
      else if (mode == 1)
      {
        y.x = ar::add(cv, ar::mul(unity_ramp(4 * q), step));
        y.y = ar::add(cv, ar::mul(unity_ramp(4 * q + 1), step));
        y.z = ar::add(cv, ar::mul(unity_ramp(4 * q + 2), step));
        y.w = ar::add(cv, ar::mul(unity_ramp(4 * q + 3), step));
      }
      else
      {
        y = c4[q];
        y.x = ar::add(y.x, step), y.y = ar::add(y.y, step), y.z = ar::add(y.z, step), y.w = ar::add(y.w, step);
      }
      c4[q] = y;
      sts128(out_addr + (uint32_t)q * 16u, y);
    }
    remaining--;
  }
  a.state[(size_t)nd.st_off * a.V + v] = f2u(step);
  a.state[(size_t)(nd.st_off + 1) * a.V + v] = f2u(target);
  a.state[(size_t)(nd.st_off + 2) * a.V + v] = (uint32_t)remaining;
}

// Interpolator1::operator()(float), G:416-422
template <bool EX>
MLB_DEV void run_interp1_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x,
                              uint32_t out_addr)
{
  using ar = A<EX>;
  if (!live) return;
  const float cur = u2f(a.state[(size_t)nd.st_off * a.V + v]);
  const float f = x.get(0);
  const float dydt = ar::sub(f, cur);
#pragma unroll 4
  for (int n = 0; n < MLB_BLOCK; ++n) sts32(out_addr + (uint32_t)n * 4u, ar::add(cur, ar::mul(unity_ramp(n), dydt)));
  a.state[(size_t)nd.st_off * a.V + v] = f2u(f);
}

template <bool EX>
MLB_DEV RingRef node_ring(const GNode& nd, const GenericArgs& a, int v, int t, float max_delay)
{
  RingRef r;
  r.mask = ring_mask_of(max_delay);
  r.w = (uint32_t)(((a.blocks_done + t) * MLB_BLOCK) & (long long)r.mask);
  r.p = a.dmem + nd.ring_off + (size_t)v * nd.ring_stride;
  return r;
}
// IntegerDelay block write, F:836-851 (w is a multiple of 64 and the ring holds >= 64 samples)
MLB_DEV void ring_write_block(const RingRef& r, RowRef x)
{
  float4* d4 = reinterpret_cast<float4*>(r.p + r.w);
#pragma unroll 4
  for (int q = 0; q < 16; ++q) d4[q] = x.get4(q);
}
// IntegerDelay::operator()(vx) read half, F:853-869: 64 samples from (w - d) & mask
MLB_DEV void ring_read_block(const RingRef& r, int32_t d, uint32_t dst)
{
  const uint32_t rd = (r.w - (uint32_t)d) & r.mask;
#pragma unroll 8
  for (int n = 0; n < MLB_BLOCK; ++n) sts32(dst + (uint32_t)n * 4u, r.p[(rd + (uint32_t)n) & r.mask]);
}
// in-place Allpass1 over a shared row
template <bool EX>
MLB_DEV void allpass1_row(uint32_t row, int n0, int n1, float& x1, float& y1, float coeff)
{
#pragma unroll 4
  for (int n = n0; n < n1; ++n)
    sts32(row + (uint32_t)n * 4u, allpass1_tick<EX>(lds32(row + (uint32_t)n * 4u), x1, y1, coeff));
}

// IntegerDelay::operator()(vx), F:834-875.  coef: delay, maxDelay
template <bool EX>
MLB_DEV void run_int_delay_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                uint32_t out_addr)
{
  if (!live) return;
  const int32_t d = cvt_trunc(a.coef[(size_t)nd.co_off * a.V + v]);
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)(nd.co_off + 1) * a.V + v]);
  ring_write_block(r, x);
  ring_read_block(r, d, out_addr);
}
// IntegerDelay::operator()(x, delay), F:877-896.  coef: maxDelay
template <bool EX>
MLB_DEV void run_int_delay_var_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                    RowRef dl, uint32_t out_addr, uint32_t old_addr)
{
  if (!live) return;
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)nd.co_off * a.V + v]);
  bool ahead = false;
  for (int n = 0; n < MLB_BLOCK; ++n) ahead |= delay_reads_ahead(cvt_trunc(dl.get(n)), r.mask);
  if (ahead) row_global_to_smem(r.p + r.w, old_addr);
  ring_write_block(r, x);
#pragma unroll 4
  for (int n = 0; n < MLB_BLOCK; ++n)
    sts32(out_addr + (uint32_t)n * 4u, ring_read(r, n, cvt_trunc(dl.get(n)), ahead, old_addr));
}
// FractionalDelay::operator()(vx), F:1014.  state: allpass x1,y1 at st_off; delay d
template <bool EX>
MLB_DEV void frac_delay_block(const RingRef& r, const GenericArgs& a, int st_off, int v, float d, RowRef x,
                              uint32_t dst)
{
  int32_t di;
  float coeff;
  frac_split<EX>(d, di, coeff);
  ring_write_block(r, x);
  ring_read_block(r, di, dst);
  float x1 = u2f(a.state[(size_t)st_off * a.V + v]), y1 = u2f(a.state[(size_t)(st_off + 1) * a.V + v]);
  allpass1_row<EX>(dst, 0, MLB_BLOCK, x1, y1, coeff);
  a.state[(size_t)st_off * a.V + v] = f2u(x1);
  a.state[(size_t)(st_off + 1) * a.V + v] = f2u(y1);
}
template <bool EX>
MLB_DEV void run_frac_delay_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                 uint32_t out_addr)
{
  if (!live) return;
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)(nd.co_off + 1) * a.V + v]);
  frac_delay_block<EX>(r, a, nd.st_off, v, a.coef[(size_t)nd.co_off * a.V + v], x, out_addr);
}
// FractionalDelay::operator()(vx, vDelay), F:1033-1042
template <bool EX>
MLB_DEV void run_frac_delay_var_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                     RowRef dl, uint32_t out_addr, uint32_t old_addr)
{
  if (!live) return;
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)nd.co_off * a.V + v]);
  bool ahead = false;
  for (int n = 0; n < MLB_BLOCK; ++n)
  {
    int32_t di;
    float coeff;
    frac_split<EX>(dl.get(n), di, coeff);
    ahead |= delay_reads_ahead(di, r.mask);
  }
  if (ahead) row_global_to_smem(r.p + r.w, old_addr);
  ring_write_block(r, x);
  float x1 = u2f(a.state[(size_t)nd.st_off * a.V + v]), y1 = u2f(a.state[(size_t)(nd.st_off + 1) * a.V + v]);
#pragma unroll 2
  for (int n = 0; n < MLB_BLOCK; ++n)
  {
    int32_t di;
    float coeff;
    frac_split<EX>(dl.get(n), di, coeff);
    sts32(out_addr + (uint32_t)n * 4u, allpass1_tick<EX>(ring_read(r, n, di, ahead, old_addr), x1, y1, coeff));
  }
  a.state[(size_t)nd.st_off * a.V + v] = f2u(x1);
  a.state[(size_t)(nd.st_off + 1) * a.V + v] = f2u(y1);
}

// PitchbendableDelay::operator(), F:1097-1104: two allpass-interpolated taps of ONE ring (both
// FractionalDelays receive the same input, so their rings are identical), delay 1 retuned at
// n % 32 == 16, delay 2 at n % 32 == 0 (F:1053-1076), crossfaded by the triangle kvFade.
// Leaves tap 1 in row B and tap 2 in row C; DL(n) is the delay-time operand.
template <bool EX, class DelayAt>
MLB_DEV void pitchbend_taps(const RingRef& r, const GenericArgs& a, int st_off, int v, RowRef x, DelayAt DL,
                            uint32_t B, uint32_t C, uint32_t old_addr)
{
  uint32_t st[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) st[i] = a.state[(size_t)(st_off + i) * a.V + v];
  int32_t di1[3], di2[2];
  float ac1[3], ac2[2];
  di1[0] = (int32_t)st[2], ac1[0] = u2f(st[3]);
  frac_split<EX>(DL(16), di1[1], ac1[1]);
  frac_split<EX>(DL(48), di1[2], ac1[2]);
  frac_split<EX>(DL(0), di2[0], ac2[0]);
  frac_split<EX>(DL(32), di2[1], ac2[1]);
  const bool ahead = delay_reads_ahead(di1[0], r.mask) | delay_reads_ahead(di1[1], r.mask) |
                     delay_reads_ahead(di1[2], r.mask) | delay_reads_ahead(di2[0], r.mask) |
                     delay_reads_ahead(di2[1], r.mask);
  if (ahead) row_global_to_smem(r.p + r.w, old_addr);
  ring_write_block(r, x);
  // gather both tap streams first (independent loads), then run the two recurrences
#pragma unroll
  for (int s = 0; s < 3; ++s)
  {
    const int n0 = s == 0 ? 0 : (s == 1 ? 16 : 48), n1 = s == 0 ? 16 : (s == 1 ? 48 : 64);
#pragma unroll 8
    for (int n = n0; n < n1; ++n) sts32(B + (uint32_t)n * 4u, ring_read(r, n, di1[s], ahead, old_addr));
  }
#pragma unroll
  for (int s = 0; s < 2; ++s)
  {
#pragma unroll 8
    for (int n = 32 * s; n < 32 * s + 32; ++n) sts32(C + (uint32_t)n * 4u, ring_read(r, n, di2[s], ahead, old_addr));
  }
  float x1 = u2f(st[0]), y1 = u2f(st[1]);
  allpass1_row<EX>(B, 0, 16, x1, y1, ac1[0]);
  allpass1_row<EX>(B, 16, 48, x1, y1, ac1[1]);
  allpass1_row<EX>(B, 48, 64, x1, y1, ac1[2]);
  st[0] = f2u(x1), st[1] = f2u(y1), st[2] = (uint32_t)di1[2], st[3] = f2u(ac1[2]);
  x1 = u2f(st[4]), y1 = u2f(st[5]);
  allpass1_row<EX>(C, 0, 32, x1, y1, ac2[0]);
  allpass1_row<EX>(C, 32, 64, x1, y1, ac2[1]);
  st[4] = f2u(x1), st[5] = f2u(y1), st[6] = (uint32_t)di2[1], st[7] = f2u(ac2[1]);
#pragma unroll
  for (int i = 0; i < 8; ++i) a.state[(size_t)(st_off + i) * a.V + v] = st[i];
}
// kvFade, F:1056-1062: 2 * (r > 16 ? 1 - r/32 : r/32), r = n % 32 (all values exact)
MLB_DEV float pitchbend_fade(int n)
{
  const int rr = n & 31;
  const float u = __int2float_rn(rr) * 0.03125f;
  return 2.f * (rr > 16 ? 1.0f - u : u);
}
template <bool EX>
MLB_DEV float pitchbend_mix(uint32_t B, uint32_t C, int n)
{
  const float b = lds32(B + (uint32_t)n * 4u), c = lds32(C + (uint32_t)n * 4u);
  return A<EX>::add(b, A<EX>::mul(pitchbend_fade(n), A<EX>::sub(c, b)));  // lerp, O:744
}

template <bool EX>
MLB_DEV void run_pitchbend_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                RowRef dl, uint32_t out_addr, uint32_t B, uint32_t C, uint32_t old_addr)
{
  if (!live) return;
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)nd.co_off * a.V + v]);
  pitchbend_taps<EX>(r, a, nd.st_off, v, x, [&](int n) { return dl.get(n); }, B, C, old_addr);
#pragma unroll 4
  for (int n = 0; n < MLB_BLOCK; ++n) sts32(out_addr + (uint32_t)n * 4u, pitchbend_mix<EX>(B, C, n));
}

// Allpass<DELAY>::operator(), F:1135-1153: din = x - vy1 * (-g); y = din * (-g) + vy1; the row
// `din` goes to scratch A, y to the node's output row.  vy1 is this node's member row.
template <bool EX>
MLB_DEV void allpass_pre(const float* vy1, float gain, RowRef x, uint32_t A_addr, uint32_t out_addr)
{
  using ar = A<EX>;
  const float g = -gain;
  const float4* v4 = reinterpret_cast<const float4*>(vy1);
#pragma unroll 2
  for (int q = 0; q < 16; ++q)
  {
    const float4 xi = x.get4(q), yi = v4[q];
    float4 din, y;
    din.x = ar::sub(xi.x, ar::mul(yi.x, g)), y.x = ar::add(ar::mul(din.x, g), yi.x);
    din.y = ar::sub(xi.y, ar::mul(yi.y, g)), y.y = ar::add(ar::mul(din.y, g), yi.y);
    din.z = ar::sub(xi.z, ar::mul(yi.z, g)), y.z = ar::add(ar::mul(din.z, g), yi.z);
    din.w = ar::sub(xi.w, ar::mul(yi.w, g)), y.w = ar::add(ar::mul(din.w, g), yi.w);
    sts128(A_addr + (uint32_t)q * 16u, din);
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
}
// OP = ALLPASS_INT / ALLPASS_FRAC (coef mGain, delay, maxDelay) or ALLPASS_PB (coef mGain, maxDelay;
// second operand = delay times)
template <int OP, bool EX>
MLB_DEV void run_allpass_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                              RowRef dl, uint32_t out_addr, uint32_t SA, uint32_t B, uint32_t C,
                              uint32_t old_addr)
{
  if (!live) return;
  float* vy1 = node_row(nd, a, v);
  allpass_pre<EX>(vy1, a.coef[(size_t)nd.co_off * a.V + v], x, SA, out_addr);
  RowRef din;
  din.is_row = true, din.addr = SA, din.k = 0.f;
  const float blk = (float)MLB_BLOCK;
  if constexpr (OP == MLB_OP_ALLPASS_PB)
  {
    const RingRef r = node_ring<EX>(nd, a, v, t, A<EX>::sub(a.coef[(size_t)(nd.co_off + 1) * a.V + v], blk));
    pitchbend_taps<EX>(r, a, nd.st_off, v, din, [&](int n) { return A<EX>::sub(dl.get(n), blk); }, B, C,
                       old_addr);
#pragma unroll 4
    for (int n = 0; n < MLB_BLOCK; ++n) vy1[n] = pitchbend_mix<EX>(B, C, n);
  }
  else
  {
    const float d = A<EX>::sub(a.coef[(size_t)(nd.co_off + 1) * a.V + v], blk);  // setDelayInSamples(d - 64)
    const RingRef r = node_ring<EX>(nd, a, v, t, A<EX>::sub(a.coef[(size_t)(nd.co_off + 2) * a.V + v], blk));
    if constexpr (OP == MLB_OP_ALLPASS_INT)
    {
      ring_write_block(r, din);
      ring_read_block(r, cvt_trunc(d), B);
    }
    else
      frac_delay_block<EX>(r, a, nd.st_off, v, d, din, B);
    RowRef res;
    res.is_row = true, res.addr = B, res.k = 0.f;
    row_smem_to_global(res, vy1);
  }
}

// FDN<8>::operator() for one voice per lane (reference F:1195-1238); rings in HBM.
template <bool EX>
MLB_DEV void run_fdn8_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                           uint32_t outL, uint32_t outR)
{
  using ar = A<EX>;
  if (!live) return;
  const uint32_t mask = (uint32_t)a.fdn_ring_len - 1u;
  const uint32_t w = (uint32_t)(((a.blocks_done + t) * MLB_BLOCK) & (long long)mask);
  float* ring = a.fdn_ring + (size_t)v * 8 * a.fdn_ring_len;
  float* carry = a.fdn_carry + (size_t)v * 8 * MLB_BLOCK;
  float a0[8], b1[8], gn[8], y1[8];
  uint32_t rd[8];
#pragma unroll
  for (int l = 0; l < 8; ++l)
  {
    a0[l] = a.coef[(size_t)(nd.co_off + l) * a.V + v];
    b1[l] = a.coef[(size_t)(nd.co_off + 8 + l) * a.V + v];
    gn[l] = a.coef[(size_t)(nd.co_off + 16 + l) * a.V + v];
    const int len = (int)a.coef[(size_t)(nd.co_off + 24 + l) * a.V + v];
    rd[l] = (w - (uint32_t)len) & mask;
    y1[l] = u2f(a.state[(size_t)(nd.st_off + l) * a.V + v]);
    // IntegerDelay: write the carried input vector at w (F:836-851) ...
    for (int i = 0; i < MLB_BLOCK; ++i)
      ring[(size_t)l * a.fdn_ring_len + ((w + i) & mask)] = carry[l * MLB_BLOCK + i];
  }
  for (int i = 0; i < MLB_BLOCK; ++i)
  {
    float d[8];
#pragma unroll
    for (int l = 0; l < 8; ++l)  // ... then read at (w - len) & mask (F:853-869)
      d[l] = ring[(size_t)l * a.fdn_ring_len + ((rd[l] + i) & mask)];
    float sumR = 0.f, sumL = 0.f, sum = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l)
    {
      if (l & 1)
        sumL = __fadd_rn(sumL, d[l]);
      else
        sumR = __fadd_rn(sumR, d[l]);
      sum = __fadd_rn(sum, d[l]);
    }
    sts32(outL + (uint32_t)i * 4u, sumL);
    sts32(outR + (uint32_t)i * 4u, sumR);
    sum = __fmul_rn(sum, 0.25f);  // 2/SIZE, exact
    const float xi = x.get(i);
#pragma unroll
    for (int l = 0; l < 8; ++l)
    {
      const float vin = __fsub_rn(d[l], sum);
      y1[l] = ar::mul_add_mul(a0[l], vin, b1[l], y1[l]);
      carry[l * MLB_BLOCK + i] = ar::add(ar::mul(y1[l], gn[l]), xi);
    }
  }
#pragma unroll
  for (int l = 0; l < 8; ++l) a.state[(size_t)(nd.st_off + l) * a.V + v] = f2u(y1[l]);
}

template <bool EX>
__global__ void __launch_bounds__(32) generic_graph_kernel(const GenericArgs a)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x;
  const int group = blockIdx.x;
  const int v0 = group * 32;
  const int v = v0 + lane;
  const bool live = v < a.V;
  const uint32_t rows = smem_u32(smem_raw) + (uint32_t)lane * kRowStride;  // [slot][lane][68]
  // scratch rows of the delay functors: delay input, two tap streams, the ring's oldest block
  const uint32_t sA = rows + (uint32_t)a.scratch_slot * kSlotBytes, sB = sA + kSlotBytes;
  const uint32_t sC = sB + kSlotBytes, sOld = sC + kSlotBytes;

  for (int t = 0; t < a.T; ++t)
  {
    for (int i = 0; i < a.n_nodes; ++i)
    {
      GNode nd = a.nodes[i];
      RowRef r[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
      {
        r[k].is_row = nd.in_kind[k] == OPERAND_SLOT;
        r[k].addr = rows + (uint32_t)nd.in_ref[k] * kSlotBytes;
        r[k].k = (nd.in_kind[k] == OPERAND_PARAM && live) ? a.coef[(size_t)nd.in_ref[k] * a.V + v]
                                                           : 0.f;
      }
      const uint32_t o = rows + (uint32_t)(nd.out_slot < 0 ? 0 : nd.out_slot) * kSlotBytes;
      switch (nd.op)
      {
        case MLB_OP_PARAM:
        case MLB_OP_FDN8_R: break;  // PARAM is an operand kind; FDN8_R aliases FDN8's 2nd slot
        case MLB_OP_INPUT:
        {
          const float4* src = reinterpret_cast<const float4*>(
              a.in + (((size_t)t * a.n_in + nd.iarg) * a.V + (live ? v : 0)) * MLB_BLOCK);
#pragma unroll 4
          for (int q = 0; q < 16; ++q)
            sts128(o + (uint32_t)q * 16u, live ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f));
          break;
        }
#define MLB_GEN_CASE(OPN) \
  case OPN: run_gen_node<OPN, EX>(nd, a, v, live, r[0], r[1], o); break;
          MLB_GEN_CASE(MLB_OP_NOISE)
          MLB_GEN_CASE(MLB_OP_PHASOR)
          MLB_GEN_CASE(MLB_OP_SINE)
          MLB_GEN_CASE(MLB_OP_SAW)
          MLB_GEN_CASE(MLB_OP_PULSE)
          MLB_GEN_CASE(MLB_OP_TICK)
#undef MLB_GEN_CASE
#define MLB_FLT_CASE(OPN) \
  case OPN: run_filter_node<OPN, EX>(nd, a, v, live, r[0], o); break;
          MLB_FLT_CASE(MLB_OP_LOPASS)
          MLB_FLT_CASE(MLB_OP_HIPASS)
          MLB_FLT_CASE(MLB_OP_BANDPASS)
          MLB_FLT_CASE(MLB_OP_LOSHELF)
          MLB_FLT_CASE(MLB_OP_HISHELF)
          MLB_FLT_CASE(MLB_OP_BELL)
          MLB_FLT_CASE(MLB_OP_ONEPOLE)
          MLB_FLT_CASE(MLB_OP_DCBLOCKER)
          MLB_FLT_CASE(MLB_OP_DIFFERENTIATOR)
          MLB_FLT_CASE(MLB_OP_INTEGRATOR)
#undef MLB_FLT_CASE
        case MLB_OP_FDN8:
          run_fdn8_node<EX>(nd, a, v, live, t, r[0], o,
                            rows + (uint32_t)nd.out_slot2 * kSlotBytes);
          break;
#define MLB_FUN_CASE(OPN) \
  case OPN: run_functor_node<OPN, EX>(nd, a, v, live, r[0], o); break;
          MLB_FUN_CASE(MLB_OP_ONESHOT)
          MLB_FUN_CASE(MLB_OP_PEAK)
          MLB_FUN_CASE(MLB_OP_RMS)
          MLB_FUN_CASE(MLB_OP_ADSR)
          MLB_FUN_CASE(MLB_OP_ALLPASS1)
          MLB_FUN_CASE(MLB_OP_SAMPLE_GLIDE)
#undef MLB_FUN_CASE
        case MLB_OP_GLIDE: run_glide_node<EX>(nd, a, v, live, r[0], o); break;
        case MLB_OP_INTERPOLATOR1: run_interp1_node<EX>(nd, a, v, live, r[0], o); break;
        case MLB_OP_INTEGER_DELAY: run_int_delay_node<EX>(nd, a, v, live, t, r[0], o); break;
        case MLB_OP_INTEGER_DELAY_VAR: run_int_delay_var_node<EX>(nd, a, v, live, t, r[0], r[1], o, sOld); break;
        case MLB_OP_FRACTIONAL_DELAY: run_frac_delay_node<EX>(nd, a, v, live, t, r[0], o); break;
        case MLB_OP_FRACTIONAL_DELAY_VAR:
          run_frac_delay_var_node<EX>(nd, a, v, live, t, r[0], r[1], o, sOld);
          break;
        case MLB_OP_PITCHBEND_DELAY: run_pitchbend_node<EX>(nd, a, v, live, t, r[0], r[1], o, sB, sC, sOld); break;
        case MLB_OP_ALLPASS_INT:
          run_allpass_node<MLB_OP_ALLPASS_INT, EX>(nd, a, v, live, t, r[0], r[1], o, sA, sB, sC, sOld);
          break;
        case MLB_OP_ALLPASS_FRAC:
          run_allpass_node<MLB_OP_ALLPASS_FRAC, EX>(nd, a, v, live, t, r[0], r[1], o, sA, sB, sC, sOld);
          break;
        case MLB_OP_ALLPASS_PB:
          run_allpass_node<MLB_OP_ALLPASS_PB, EX>(nd, a, v, live, t, r[0], r[1], o, sA, sB, sC, sOld);
          break;
        case MLB_OP_FEEDBACK_READ:
          if (live) row_global_to_smem(node_row(nd, a, v), o);
          break;
        case MLB_OP_FEEDBACK_WRITE:  // row_off is the FEEDBACK_READ node's row
          if (live)
          {
            row_smem_to_global(r[0], node_row(nd, a, v));
#pragma unroll 4
            for (int q = 0; q < 16; ++q) sts128(o + (uint32_t)q * 16u, r[0].get4(q));
          }
          break;
        default: dispatch_stateless<EX>(nd.op, r[0], r[1], r[2], o); break;
      }

      if (nd.out_plane >= 0)
      {
        // source row of this output: PARAM nodes broadcast their scalar
        RowRef y;
        y.is_row = nd.op != MLB_OP_PARAM;
        y.addr = o;
        y.k = (nd.op == MLB_OP_PARAM && live) ? a.coef[(size_t)nd.co_off * a.V + v] : 0.f;
        if (a.out != nullptr && live)
        {
          float4* dst = reinterpret_cast<float4*>(
              a.out + (((size_t)t * a.n_out + nd.out_plane) * a.V + v) * MLB_BLOCK);
#pragma unroll 4
          for (int q = 0; q < 16; ++q) dst[q] = y.get4(q);
        }
        if (a.mix_partial != nullptr)
        {
          // lane n sums samples n and n+32 over the 32 voice rows of this group, in voice order
          __syncwarp();
          const uint32_t colbase = o - (uint32_t)lane * kRowStride;  // [voice 0][sample 0]
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
          {
            const int n = hh * 32 + lane;
            float acc = 0.f;
            for (int rr = 0; rr < 32; ++rr)
            {
              float xv;
              if (nd.op == MLB_OP_PARAM)
                xv = __shfl_sync(0xffffffffu, y.k, rr);
              else
                xv = lds32(colbase + (uint32_t)rr * kRowStride + (uint32_t)n * 4u);
              if (v0 + rr < a.V) acc = __fadd_rn(acc, xv);
            }
            a.mix_partial[((size_t)(t * a.n_out + nd.out_plane) * a.n_groups + group) * MLB_BLOCK +
                          n] = acc;
          }
          __syncwarp();
        }
      }
      __syncwarp();
    }
  }
}

// ---- K3: stateless elementwise ops over n_rows*64 elements (MLDSPOps.h:567-918) ----
// One instantiation per op (the switch in op_apply folds away); float4 per thread, grid-stride.
template <int OP, bool EX>
__global__ void __launch_bounds__(256)
    map_kernel(const float4* __restrict__ x1, const float4* __restrict__ x2,
               const float4* __restrict__ x3, float4* __restrict__ y, size_t n4)
{
  constexpr int NIN = op_nin(OP);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
  {
    const float4 a = __ldcs(x1 + i);
    const float4 b = NIN >= 2 ? __ldcs(x2 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 c = NIN >= 3 ? __ldcs(x3 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r;
    r.x = op_apply<EX>(OP, a.x, b.x, c.x);
    r.y = op_apply<EX>(OP, a.y, b.y, c.y);
    r.z = op_apply<EX>(OP, a.z, b.z, c.z);
    r.w = op_apply<EX>(OP, a.w, b.w, c.w);
    __stcs(y + i, r);
  }
}

typedef void (*MapKernelFn)(const float4*, const float4*, const float4*, float4*, size_t);
inline MapKernelFn map_kernel_for(int op)
{
  switch (op)
  {
#define MLB_X_MAP(NAME, id, nin, nst, nco)                                 \
  case id:                                                                 \
    if constexpr (id >= MLB_OP_MAP_FIRST && id < MLB_OP_MAP_END)         \
      return map_kernel<id, true>;                                         \
    break;
    MLB_OP_TABLE_STATELESS(MLB_X_MAP)
#undef MLB_X_MAP
    default: break;
  }
  return nullptr;
}

}  // namespace mlb
