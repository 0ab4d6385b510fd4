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
  int in_kind[MLB_MAX_INS];
  int in_ref[MLB_MAX_INS];   // slot index or coef word index
  int out_slot;    // -1: node output never read as a row (PARAM)
  int out_slot2;   // FDN8: slot of the sumR row (read by FDN8_R), else -1
  int st_off, co_off;
  int out_plane;   // >= 0: also written to out / mix plane
  int iarg;
  int src_node;    // index in the caller's node list, -1 for an import pseudo-node
  int chan_out;    // >= 0: the row is also stored to this channel plane (read by a later stage)
  // delay memory of this node (float offsets into GenericArgs::dmem)
  unsigned ring_stride;          // floats between the rings of consecutive voices (power of two)
  unsigned long long row_off;    // [V][64] member row (Allpass::vy1, LinearGlide::mCurrVec, feedback)
  unsigned long long ring_off;   // [V][ring_stride] IntegerDelay ring
};

// internal op of the stage pipeline: load a row another stage stored in channel plane `iarg`
#define MLB_OP_IMPORT_ROW 200

// One pipeline stage = the nodes [node_begin, node_end) of the program.  Before block t the stage
// waits until every stage in wait_stage[] has finished block t, and every stage in fbwait_stage[]
// (writers of the feedback rows it reads) has finished block t-1.
struct GStage
{
  static constexpr int kMaxWait = 8;
  int node_begin, node_end;
  int n_wait, n_fbwait;
  int wait_stage[kMaxWait];
  int fbwait_stage[kMaxWait];
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
  int scratch_slot;        // first of 3 scratch row slots (graphs with Allpass<> / PitchbendableDelay nodes)
  const GStage* stages;
  int n_stages, n_chan;
  unsigned* sync;          // [0] ticket counter; [1 + s*n_groups + g] blocks finished by (stage s, group g)
  float* chan;             // channel planes [T][n_chan][V][64]
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
  // the shared-memory accessors are ordered asm statements: load four quads of every operand first,
  // then compute, then store, so that the LDS latencies overlap instead of adding up 16 times
#pragma unroll 1
  for (int q0 = 0; q0 < 16; q0 += 4)
  {
    float4 xi[4], bi[4], ci[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      xi[j] = x.get4(q0 + j);
      bi[j] = NIN >= 2 ? b.get4(q0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      ci[j] = NIN >= 3 ? c.get4(q0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      float4 y;
      y.x = op_apply<EX>(OP, xi[j].x, bi[j].x, ci[j].x);
      y.y = op_apply<EX>(OP, xi[j].y, bi[j].y, ci[j].y);
      y.z = op_apply<EX>(OP, xi[j].z, bi[j].z, ci[j].z);
      y.w = op_apply<EX>(OP, xi[j].w, bi[j].w, ci[j].w);
      sts128(out_addr + (uint32_t)(q0 + j) * 16u, y);
    }
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


// filters with coefficient ROWS (LOPASS_V / LOSHELF_V / HISHELF_V) and the device-designed LOPASS_MOD:
// operand 0 is the audio row, operands 1.. the per-sample coefficient rows (or PARAM constants)
template <int OP, bool EX>
MLB_DEV void run_vfilter_node(const GNode& nd, const GenericArgs& a, int v, bool live, uint32_t rows,
                              uint32_t out_addr)
{
  constexpr int NR = op_nin(OP) - 1;
  RowRef x, c[NR];
  x.is_row = nd.in_kind[0] == OPERAND_SLOT;
  x.addr = rows + (uint32_t)nd.in_ref[0] * kSlotBytes;
  x.k = (nd.in_kind[0] == OPERAND_PARAM && live) ? a.coef[(size_t)nd.in_ref[0] * a.V + v] : 0.f;
#pragma unroll
  for (int k = 0; k < NR; ++k)
  {
    c[k].is_row = nd.in_kind[k + 1] == OPERAND_SLOT;
    c[k].addr = rows + (uint32_t)nd.in_ref[k + 1] * kSlotBytes;
    c[k].k = (nd.in_kind[k + 1] == OPERAND_PARAM && live) ? a.coef[(size_t)nd.in_ref[k + 1] * a.V + v] : 0.f;
  }
  uint32_t st[2];
  st[0] = live ? a.state[(size_t)nd.st_off * a.V + v] : 0u;
  st[1] = live ? a.state[(size_t)(nd.st_off + 1) * a.V + v] : 0u;
#pragma unroll 1
  for (int q = 0; q < 16; ++q)
  {
    const float4 xi = x.get4(q);
    float4 ci[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) ci[k] = c[k].get4(q);
    float cv[NR];
    float4 y;
#pragma unroll
    for (int k = 0; k < NR; ++k) cv[k] = ci[k].x;
    y.x = vfilter_tick<EX>(OP, xi.x, st, cv);
#pragma unroll
    for (int k = 0; k < NR; ++k) cv[k] = ci[k].y;
    y.y = vfilter_tick<EX>(OP, xi.y, st, cv);
#pragma unroll
    for (int k = 0; k < NR; ++k) cv[k] = ci[k].z;
    y.z = vfilter_tick<EX>(OP, xi.z, st, cv);
#pragma unroll
    for (int k = 0; k < NR; ++k) cv[k] = ci[k].w;
    y.w = vfilter_tick<EX>(OP, xi.w, st, cv);
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
  if (live)
  {
    a.state[(size_t)nd.st_off * a.V + v] = st[0];
    a.state[(size_t)(nd.st_off + 1) * a.V + v] = st[1];
  }
}

// RAMP(start, end) = interpolateDSPVectorLinear(start[0], end[0]), O:986-990
template <bool EX>
MLB_DEV void run_ramp_node(RowRef s, RowRef e, uint32_t out_addr)
{
  const float s0 = s.get(0), e0 = e.get(0);
#pragma unroll 4
  for (int q = 0; q < 16; ++q)
  {
    float4 y;
    y.x = ramp_sample<EX>(s0, e0, 4 * q);
    y.y = ramp_sample<EX>(s0, e0, 4 * q + 1);
    y.z = ramp_sample<EX>(s0, e0, 4 * q + 2);
    y.w = ramp_sample<EX>(s0, e0, 4 * q + 3);
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
}

// ---- SURVEY 8(f) row 2 node runners (device functions in functors.cuh) ----

template <int OP, bool EX>
MLB_DEV float functor_tick(float x, uint32_t* st, const float* co)
{
  if constexpr (OP == MLB_OP_ONESHOT) return oneshot_tick<EX>(x, st);
  if constexpr (OP == MLB_OP_IMPULSE) return impulse_tick<EX>(x, st);
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
// (the shared-memory stores are asm volatile with a memory clobber: global loads are issued in
// batches into registers first so that they overlap instead of serialising behind each store)
MLB_DEV void row_global_to_smem(const float* src, uint32_t dst)
{
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int h = 0; h < 2; ++h)
  {
    float4 buf[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) buf[q] = s4[h * 8 + q];
#pragma unroll
    for (int q = 0; q < 8; ++q) sts128(dst + (uint32_t)(h * 8 + q) * 16u, buf[q]);
  }
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
// The fixed-delay and per-sample-delay runners keep 16-sample batches in registers (16 independent ring
// loads in flight; the shared-memory stores are asm volatile + memory clobber and would serialise loads
// issued between them).  In the per-sample-delay runners a lane whose delay can reach into the block being
// written ("ahead", functors.cuh) keeps the ring's oldest block in a per-thread local array -- rare, so it
// lives in local memory.  The pitch-bendable delay / allpass nodes use cp.async instead (below).
struct OldBlock
{
  float v[MLB_BLOCK];
  MLB_DEV void save(const RingRef& r)
  {
    const float4* s4 = reinterpret_cast<const float4*>(r.p + r.w);
#pragma unroll 4
    for (int q = 0; q < 16; ++q)
    {
      const float4 t = s4[q];
      v[4 * q] = t.x, v[4 * q + 1] = t.y, v[4 * q + 2] = t.z, v[4 * q + 3] = t.w;
    }
  }
};

// IntegerDelay block write, F:836-851 (w is a multiple of 64 and the ring holds >= 64 samples)
MLB_DEV void ring_write_block(const RingRef& r, RowRef x)
{
  float4* d4 = reinterpret_cast<float4*>(r.p + r.w);
#pragma unroll 4
  for (int q = 0; q < 16; ++q) d4[q] = x.get4(q);
}
MLB_DEV void store16(uint32_t row, int n0, const float (&y)[16])
{
#pragma unroll
  for (int q = 0; q < 4; ++q)
    sts128(row + (uint32_t)(n0 + 4 * q) * 4u, make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]));
}
MLB_DEV void store16(float* row, int n0, const float (&y)[16])
{
  float4* d4 = reinterpret_cast<float4*>(row + n0);
#pragma unroll
  for (int q = 0; q < 4; ++q) d4[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
}

// IntegerDelay::operator()(vx) read half (F:853-869), optionally followed by the Allpass1 of a
// FractionalDelay (F:1014); DST = shared row address or global row pointer
template <bool EX, bool AP, class DST>
MLB_DEV void ring_block_out(const RingRef& r, int32_t d, float& x1, float& y1, float coeff, DST dst)
{
  const uint32_t rd = (r.w - (uint32_t)d) & r.mask;
  float cur[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) cur[j] = r.p[(rd + (uint32_t)j) & r.mask];
#pragma unroll
  for (int s = 0; s < 4; ++s)
  {
    float nxt[16];
    if (s < 3)
    {
#pragma unroll
      for (int j = 0; j < 16; ++j) nxt[j] = r.p[(rd + (uint32_t)(16 * s + 16 + j)) & r.mask];
    }
    if (AP)
    {
#pragma unroll
      for (int j = 0; j < 16; ++j) cur[j] = allpass1_tick<EX>(cur[j], x1, y1, coeff);
    }
    store16(dst, 16 * s, cur);
    if (s < 3)
    {
#pragma unroll
      for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
    }
  }
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
  float x1 = 0.f, y1 = 0.f;
  ring_block_out<EX, false>(r, d, x1, y1, 0.f, out_addr);
}
// IntegerDelay::operator()(x, delay), F:877-896 (coef: maxDelay) and
// FractionalDelay::operator()(vx, vDelay), F:1033-1042 (coef: maxDelay; state: allpass x1, y1)
template <bool EX, bool FRAC>
MLB_DEV void run_delay_var_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                RowRef dl, uint32_t out_addr)
{
  if (!live) return;
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)nd.co_off * a.V + v]);
  OldBlock old;
  bool ahead = false;
#pragma unroll 4
  for (int n = 0; n < MLB_BLOCK; ++n)
  {
    int32_t di;
    float coeff;
    if (FRAC)
      frac_split<EX>(dl.get(n), di, coeff);
    else
      di = cvt_trunc(dl.get(n));
    ahead |= delay_reads_ahead(di, r.mask);
  }
  if (ahead) old.save(r);
  ring_write_block(r, x);
  float x1 = 0.f, y1 = 0.f;
  if (FRAC) x1 = u2f(a.state[(size_t)nd.st_off * a.V + v]), y1 = u2f(a.state[(size_t)(nd.st_off + 1) * a.V + v]);
#pragma unroll 1
  for (int n0 = 0; n0 < MLB_BLOCK; n0 += 16)
  {
    float buf[16], co[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
    {
      int32_t di;
      co[j] = 0.f;
      if (FRAC)
        frac_split<EX>(dl.get(n0 + j), di, co[j]);
      else
        di = cvt_trunc(dl.get(n0 + j));
      const uint32_t m = ((uint32_t)(n0 + j) - (uint32_t)di) & r.mask;
      buf[j] = r.p[(r.w + m) & r.mask];
      if (ahead && m < (uint32_t)MLB_BLOCK && m > (uint32_t)(n0 + j)) buf[j] = old.v[m];
    }
    if (FRAC)
    {
#pragma unroll
      for (int j = 0; j < 16; ++j) buf[j] = allpass1_tick<EX>(buf[j], x1, y1, co[j]);
    }
    store16(out_addr, n0, buf);
  }
  if (FRAC)
  {
    a.state[(size_t)nd.st_off * a.V + v] = f2u(x1);
    a.state[(size_t)(nd.st_off + 1) * a.V + v] = f2u(y1);
  }
}
// FractionalDelay::operator()(vx), F:1014.  coef: delay, maxDelay; state: allpass x1, y1
template <bool EX>
MLB_DEV void run_frac_delay_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                 uint32_t out_addr)
{
  if (!live) return;
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)(nd.co_off + 1) * a.V + v]);
  int32_t di;
  float coeff;
  frac_split<EX>(a.coef[(size_t)nd.co_off * a.V + v], di, coeff);
  ring_write_block(r, x);
  float x1 = u2f(a.state[(size_t)nd.st_off * a.V + v]), y1 = u2f(a.state[(size_t)(nd.st_off + 1) * a.V + v]);
  ring_block_out<EX, true>(r, di, x1, y1, coeff, out_addr);
  a.state[(size_t)nd.st_off * a.V + v] = f2u(x1);
  a.state[(size_t)(nd.st_off + 1) * a.V + v] = f2u(y1);
}

// kvFade, F:1056-1062: 2 * (r > 16 ? 1 - r/32 : r/32), r = n % 32 (all values exact)
MLB_DEV float pitchbend_fade(int n)
{
  const int rr = n & 31;
  const float u = __int2float_rn(rr) * 0.03125f;
  return 2.f * (rr > 16 ? 1.0f - u : u);
}

// PitchbendableDelay::operator(), F:1097-1104: two allpass-interpolated taps of ONE ring (both
// FractionalDelays receive the same input, so their rings are identical), delay 1 retuned at
// n % 32 == 16, delay 2 at n % 32 == 0 (F:1053-1076), crossfaded by the triangle kvFade.
struct PitchbendPlan
{
  uint32_t st[8];
  int32_t di1[3], di2[2];
  float ac1[3], ac2[2];
};
template <bool EX, class DelayAt>
MLB_DEV void pitchbend_prepare(PitchbendPlan& p, const GenericArgs& a, int st_off, int v, DelayAt DL)
{
#pragma unroll
  for (int i = 0; i < 8; ++i) p.st[i] = a.state[(size_t)(st_off + i) * a.V + v];
  p.di1[0] = (int32_t)p.st[2], p.ac1[0] = u2f(p.st[3]);
  frac_split<EX>(DL(16), p.di1[1], p.ac1[1]);
  frac_split<EX>(DL(48), p.di1[2], p.ac1[2]);
  frac_split<EX>(DL(0), p.di2[0], p.ac2[0]);
  frac_split<EX>(DL(32), p.di2[1], p.ac2[1]);
}

// cp.async: global -> shared without a register round trip; completion is per thread
MLB_DEV void cp_async4(uint32_t smem_dst, const float* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(src) : "memory");
}
MLB_DEV void cp_async16(uint32_t smem_dst, const float* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(src) : "memory");
}
MLB_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Tap sample n with delay d reads ring slot w + m, m = (n - d) & mask.  A slot of the CURRENT block
// that the per-sample loop (F:898-912) has already written when it reads it (m <= n) holds this
// block's input sample m; every other slot still holds what was in the ring before this block --
// including slots of the current block the loop has NOT written yet (m > n, only for delays beyond the
// ring).  So all "old" values can be fetched before the block is stored, all at once and asynchronously,
// and the "new" ones are taken from the input row afterwards.
MLB_DEV bool tap_is_new(uint32_t m, int n) { return m <= (uint32_t)n; }
MLB_DEV void tap_prefetch16(const RingRef& r, int n0, int32_t d, uint32_t row)
{
#pragma unroll
  for (int j = 0; j < 16; ++j)
  {
    const uint32_t m = ((uint32_t)(n0 + j) - (uint32_t)d) & r.mask;
    if (!tap_is_new(m, n0 + j)) cp_async4(row + (uint32_t)(n0 + j) * 4u, r.p + ((r.w + m) & r.mask));
  }
}
template <class InputAt>
MLB_DEV float tap_value(const RingRef& r, int n, int32_t d, uint32_t row, InputAt X)
{
  const uint32_t m = ((uint32_t)n - (uint32_t)d) & r.mask;
  return tap_is_new(m, n) ? X((int)m) : lds32(row + (uint32_t)n * 4u);
}

// both taps of a PitchbendableDelay into rows B and C (issue only; wait with cp_async_wait_all)
MLB_DEV void pitchbend_prefetch(const PitchbendPlan& p, const RingRef& r, uint32_t B, uint32_t C)
{
  tap_prefetch16(r, 0, p.di1[0], B);
  tap_prefetch16(r, 16, p.di1[1], B);
  tap_prefetch16(r, 32, p.di1[1], B);
  tap_prefetch16(r, 48, p.di1[2], B);
  tap_prefetch16(r, 0, p.di2[0], C);
  tap_prefetch16(r, 16, p.di2[0], C);
  tap_prefetch16(r, 32, p.di2[1], C);
  tap_prefetch16(r, 48, p.di2[1], C);
}
// the two allpass recurrences and the crossfade; X(m) = this block's input sample m
template <bool EX, class DST, class InputAt>
MLB_DEV void pitchbend_run(PitchbendPlan& p, const RingRef& r, uint32_t B, uint32_t C, InputAt X,
                           const GenericArgs& a, int st_off, int v, DST dst)
{
  float xb = u2f(p.st[0]), yb = u2f(p.st[1]), xc = u2f(p.st[4]), yc = u2f(p.st[5]);
#pragma unroll
  for (int s = 0; s < 4; ++s)
  {
    // quarter s: tap 1 uses segment {0,1,1,2}[s], tap 2 segment {0,0,1,1}[s]
    const int s1 = s == 0 ? 0 : (s == 3 ? 2 : 1), s2 = s >= 2 ? 1 : 0;
    const float cb = p.ac1[s1], cc = p.ac2[s2];
    float y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
    {
      const int n = 16 * s + j;
      const float b = allpass1_tick<EX>(tap_value(r, n, p.di1[s1], B, X), xb, yb, cb);
      const float c = allpass1_tick<EX>(tap_value(r, n, p.di2[s2], C, X), xc, yc, cc);
      y[j] = A<EX>::add(b, A<EX>::mul(pitchbend_fade(n), A<EX>::sub(c, b)));  // lerp, O:744
    }
    store16(dst, 16 * s, y);
  }
  p.st[0] = f2u(xb), p.st[1] = f2u(yb), p.st[2] = (uint32_t)p.di1[2], p.st[3] = f2u(p.ac1[2]);
  p.st[4] = f2u(xc), p.st[5] = f2u(yc), p.st[6] = (uint32_t)p.di2[1], p.st[7] = f2u(p.ac2[1]);
#pragma unroll
  for (int i = 0; i < 8; ++i) a.state[(size_t)(st_off + i) * a.V + v] = p.st[i];
}

template <bool EX>
MLB_DEV void run_pitchbend_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                                RowRef dl, uint32_t out_addr, uint32_t B, uint32_t C)
{
  if (!live) return;
  const RingRef r = node_ring<EX>(nd, a, v, t, a.coef[(size_t)nd.co_off * a.V + v]);
  PitchbendPlan p;
  pitchbend_prepare<EX>(p, a, nd.st_off, v, [&](int n) { return dl.get(n); });
  pitchbend_prefetch(p, r, B, C);
  cp_async_wait_all();
  ring_write_block(r, x);
  pitchbend_run<EX>(p, r, B, C, [&](int m) { return x.get(m); }, a, nd.st_off, v, out_addr);
}

// Allpass<DELAY>::operator(), F:1135-1153: din = x - vy1 * (-g); y = din * (-g) + vy1;
// vy1 = DELAY(din).  din goes into the ring (and stays in scratch row SA for taps that land in the
// current block), y to the node's output row; vy1 is this node's member row in delay memory.
// OP = ALLPASS_INT / ALLPASS_FRAC (coef mGain, delay, maxDelay) or ALLPASS_PB (coef mGain, maxDelay;
// second operand = delay times)
template <int OP, bool EX>
MLB_DEV void run_allpass_node(const GNode& nd, const GenericArgs& a, int v, bool live, int t, RowRef x,
                              RowRef dl, uint32_t out_addr, uint32_t SA, uint32_t B, uint32_t C)
{
  using ar = A<EX>;
  if (!live) return;
  float* vy1 = node_row(nd, a, v);
  const float blk = (float)MLB_BLOCK;
  const float md = a.coef[(size_t)(nd.co_off + (OP == MLB_OP_ALLPASS_PB ? 1 : 2)) * a.V + v];
  const RingRef r = node_ring<EX>(nd, a, v, t, ar::sub(md, blk));  // setMaxDelayInSamples(d - 64), F:1125-1128
  // everything this node reads from HBM is requested up front: vy1 -> SA, the two tap streams -> B, C
#pragma unroll
  for (int q = 0; q < 16; ++q) cp_async16(SA + (uint32_t)q * 16u, vy1 + 4 * q);
  PitchbendPlan p;
  if constexpr (OP == MLB_OP_ALLPASS_PB)
  {
    pitchbend_prepare<EX>(p, a, nd.st_off, v, [&](int n) { return ar::sub(dl.get(n), blk); });  // F:1151
    pitchbend_prefetch(p, r, B, C);
  }
  cp_async_wait_all();
  {
    const float g = -a.coef[(size_t)nd.co_off * a.V + v];
    float4* ring4 = reinterpret_cast<float4*>(r.p + r.w);
#pragma unroll 4
    for (int q = 0; q < 16; ++q)
    {
      const float4 xi = x.get4(q), yi = lds128(SA + (uint32_t)q * 16u);
      float4 din, y;
      din.x = ar::sub(xi.x, ar::mul(yi.x, g)), y.x = ar::add(ar::mul(din.x, g), yi.x);
      din.y = ar::sub(xi.y, ar::mul(yi.y, g)), y.y = ar::add(ar::mul(din.y, g), yi.y);
      din.z = ar::sub(xi.z, ar::mul(yi.z, g)), y.z = ar::add(ar::mul(din.z, g), yi.z);
      din.w = ar::sub(xi.w, ar::mul(yi.w, g)), y.w = ar::add(ar::mul(din.w, g), yi.w);
      ring4[q] = din;                            // IntegerDelay block write of the delay input
      sts128(SA + (uint32_t)q * 16u, din);       // ... kept for taps inside the current block
      sts128(out_addr + (uint32_t)q * 16u, y);
    }
  }
  if constexpr (OP == MLB_OP_ALLPASS_PB)
    pitchbend_run<EX>(p, r, B, C, [&](int m) { return lds32(SA + (uint32_t)m * 4u); }, a, nd.st_off, v, vy1);
  else
  {
    const float d = ar::sub(a.coef[(size_t)(nd.co_off + 1) * a.V + v], blk);  // setDelayInSamples(d - 64), F:1123
    if constexpr (OP == MLB_OP_ALLPASS_INT)
    {
      float x1 = 0.f, y1 = 0.f;
      ring_block_out<EX, false>(r, cvt_trunc(d), x1, y1, 0.f, vy1);
    }
    else
    {
      int32_t di;
      float coeff;
      frac_split<EX>(d, di, coeff);
      float x1 = u2f(a.state[(size_t)nd.st_off * a.V + v]), y1 = u2f(a.state[(size_t)(nd.st_off + 1) * a.V + v]);
      ring_block_out<EX, true>(r, di, x1, y1, coeff, vy1);
      a.state[(size_t)nd.st_off * a.V + v] = f2u(x1);
      a.state[(size_t)(nd.st_off + 1) * a.V + v] = f2u(y1);
    }
  }
}

// HalfBandFilter::upsampleFirstHalf + upsampleSecondHalf, F:1248-1270: one input row -> two rows
template <bool EX>
MLB_DEV void run_halfband_up_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x,
                                  uint32_t out1, uint32_t out2)
{
  if (!live) return;
  HalfBand h;
#pragma unroll
  for (int i = 0; i < 9; ++i) h.s[i] = u2f(a.state[(size_t)(nd.st_off + i) * a.V + v]);
#pragma unroll 1
  for (int q = 0; q < 16; ++q)
  {
    const float4 xi = x.get4(q);
    float4 lo, hi;
    lo.x = h.a<EX>(xi.x), lo.y = h.b<EX>(xi.x), lo.z = h.a<EX>(xi.y), lo.w = h.b<EX>(xi.y);
    hi.x = h.a<EX>(xi.z), hi.y = h.b<EX>(xi.z), hi.z = h.a<EX>(xi.w), hi.w = h.b<EX>(xi.w);
    const uint32_t dst = (q < 8 ? out1 : out2) + (uint32_t)((q & 7) * 2) * 16u;
    sts128(dst, lo);
    sts128(dst + 16u, hi);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) a.state[(size_t)(nd.st_off + i) * a.V + v] = f2u(h.s[i]);
}
// HalfBandFilter::downsample(vx1, vx2), F:1272-1294: two rows -> one row
template <bool EX>
MLB_DEV void run_halfband_down_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x1,
                                    RowRef x2, uint32_t out_addr)
{
  if (!live) return;
  HalfBand h;
#pragma unroll
  for (int i = 0; i < 9; ++i) h.s[i] = u2f(a.state[(size_t)(nd.st_off + i) * a.V + v]);
#pragma unroll 1
  for (int q = 0; q < 16; ++q)  // output float4 q <- input float4s 2q, 2q+1 of row (q < 8 ? x1 : x2)
  {
    const RowRef& x = q < 8 ? x1 : x2;
    const float4 p0 = x.get4((q & 7) * 2), p1 = x.get4((q & 7) * 2 + 1);
    float4 y;
    float a0, b0;
    a0 = h.a<EX>(p0.x), b0 = h.b<EX>(p0.y), y.x = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
    a0 = h.a<EX>(p0.z), b0 = h.b<EX>(p0.w), y.y = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
    a0 = h.a<EX>(p1.x), b0 = h.b<EX>(p1.y), y.z = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
    a0 = h.a<EX>(p1.z), b0 = h.b<EX>(p1.w), y.w = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
    sts128(out_addr + (uint32_t)q * 16u, y);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) a.state[(size_t)(nd.st_off + i) * a.V + v] = f2u(h.s[i]);
}
// Downsample2xFunction (MLDSPFunctional.h:166-223) around a stateless fn: the statements before fn(...)
// state: HalfBandFilter mDowners[0] (9 words), mPhase; member row mInputBuffer in delay memory
template <bool EX>
MLB_DEV void run_down2x_in_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x, uint32_t out_addr)
{
  if (!live) return;
  float* buffer = node_row(nd, a, v);
  const uint32_t phase = a.state[(size_t)(nd.st_off + 9) * a.V + v];
  if (phase)
  {
    row_global_to_smem(buffer, out_addr);  // mInputBuffer; then downsample(mInputBuffer, vx) in place
    HalfBand h;
#pragma unroll
    for (int i = 0; i < 9; ++i) h.s[i] = u2f(a.state[(size_t)(nd.st_off + i) * a.V + v]);
    RowRef x1;
    x1.is_row = true, x1.addr = out_addr, x1.k = 0.f;
#pragma unroll 1
    for (int q = 0; q < 16; ++q)  // output quad q needs input quads 2q, 2q+1 (>= q): safe in place
    {
      const RowRef& src = q < 8 ? x1 : x;
      const float4 p0 = src.get4((q & 7) * 2), p1 = src.get4((q & 7) * 2 + 1);
      float4 y;
      float a0, b0;
      a0 = h.a<EX>(p0.x), b0 = h.b<EX>(p0.y), y.x = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
      a0 = h.a<EX>(p0.z), b0 = h.b<EX>(p0.w), y.y = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
      a0 = h.a<EX>(p1.x), b0 = h.b<EX>(p1.y), y.z = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
      a0 = h.a<EX>(p1.z), b0 = h.b<EX>(p1.w), y.w = A<EX>::mul(A<EX>::add(a0, h.s[8]), 0.5f), h.s[8] = b0;
      sts128(out_addr + (uint32_t)q * 16u, y);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) a.state[(size_t)(nd.st_off + i) * a.V + v] = f2u(h.s[i]);
  }
  else
  {
    row_smem_to_global(x, buffer);  // mInputBuffer = vx; nothing for fn on this block
#pragma unroll 4
    for (int q = 0; q < 16; ++q) sts128(out_addr + (uint32_t)q * 16u, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  a.state[(size_t)(nd.st_off + 9) * a.V + v] = phase ? 0u : 1u;
}
// ... and the statements after fn(...): state HalfBandFilter mUppers[0], mPhase; member row mOutputBuffer
template <bool EX>
MLB_DEV void run_down2x_out_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x, uint32_t out_addr)
{
  if (!live) return;
  float* buffer = node_row(nd, a, v);
  const uint32_t phase = a.state[(size_t)(nd.st_off + 9) * a.V + v];
  if (phase)
  {
    HalfBand h;
#pragma unroll
    for (int i = 0; i < 9; ++i) h.s[i] = u2f(a.state[(size_t)(nd.st_off + i) * a.V + v]);
    float4* b4 = reinterpret_cast<float4*>(buffer);
#pragma unroll 1
    for (int q = 0; q < 16; ++q)
    {
      const float4 xi = x.get4(q);
      float4 lo, hi;
      lo.x = h.a<EX>(xi.x), lo.y = h.b<EX>(xi.x), lo.z = h.a<EX>(xi.y), lo.w = h.b<EX>(xi.y);
      hi.x = h.a<EX>(xi.z), hi.y = h.b<EX>(xi.z), hi.z = h.a<EX>(xi.w), hi.w = h.b<EX>(xi.w);
      if (q < 8)  // upsampleFirstHalf -> returned
      {
        sts128(out_addr + (uint32_t)(2 * q) * 16u, lo);
        sts128(out_addr + (uint32_t)(2 * q + 1) * 16u, hi);
      }
      else  // upsampleSecondHalf -> mOutputBuffer
        b4[2 * (q - 8)] = lo, b4[2 * (q - 8) + 1] = hi;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) a.state[(size_t)(nd.st_off + i) * a.V + v] = f2u(h.s[i]);
  }
  else
    row_global_to_smem(buffer, out_addr);
  a.state[(size_t)(nd.st_off + 9) * a.V + v] = phase ? 0u : 1u;
}
// TempoLock::operator()(x, dydx, isr), F:1494-1578
template <bool EX>
MLB_DEV void run_tempo_lock_node(const GNode& nd, const GenericArgs& a, int v, bool live, RowRef x, RowRef ratio,
                                 uint32_t out_addr)
{
  if (!live) return;
  float omega = u2f(a.state[(size_t)nd.st_off * a.V + v]), x1v = u2f(a.state[(size_t)(nd.st_off + 1) * a.V + v]);
  float dydt = 0.f;
  const bool running = tempo_lock_prepare<EX>(x.get(0), x.get(1), ratio.get(0), a.coef[(size_t)nd.co_off * a.V + v],
                                              omega, x1v, dydt);
#pragma unroll 1
  for (int q = 0; q < 16; ++q)
  {
    float y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      y[j] = running ? omega : 0.f;
      if (running)
      {
        omega = A<EX>::add(omega, dydt);
        if (omega > 1.0f) omega = A<EX>::sub(omega, 1.0f);
      }
    }
    sts128(out_addr + (uint32_t)q * 16u, make_float4(y[0], y[1], y[2], y[3]));
  }
  a.state[(size_t)nd.st_off * a.V + v] = f2u(omega);
  a.state[(size_t)(nd.st_off + 1) * a.V + v] = f2u(x1v);
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

MLB_DEV unsigned ld_acquire_u32(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
MLB_DEV void st_release_u32(unsigned* p, unsigned v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool EX>
__global__ void __launch_bounds__(32) generic_graph_kernel(const GenericArgs a)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x;
  // (stage, group) by ticket, stage-major: a CTA only ever waits for CTAs with smaller tickets,
  // which are already running or done -- no assumption about block scheduling order
  unsigned ticket = 0;
  if (lane == 0) ticket = atomicAdd(a.sync, 1u);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  const int stage = (int)(ticket / (unsigned)a.n_groups);
  const int group = (int)(ticket % (unsigned)a.n_groups);
  const GStage sg = a.stages[stage];
  unsigned* const progress = a.sync + 1;
  const int v0 = group * 32;
  const int v = v0 + lane;
  const bool live = v < a.V;
  const uint32_t rows = smem_u32(smem_raw) + (uint32_t)lane * kRowStride;  // [slot][lane][68]
  // scratch rows of the allpass / pitch-bendable delay nodes: delay input, the two tap streams
  const uint32_t sA = rows + (uint32_t)a.scratch_slot * kSlotBytes, sB = sA + kSlotBytes, sC = sB + kSlotBytes;

  for (int t = 0; t < a.T; ++t)
  {
    if (a.n_stages > 1)
    {
      if (lane < sg.n_wait)
        while (ld_acquire_u32(progress + (size_t)sg.wait_stage[lane] * a.n_groups + group) <= (unsigned)t)
          __nanosleep(100);
      if (lane >= 8 && lane - 8 < sg.n_fbwait)
        while (ld_acquire_u32(progress + (size_t)sg.fbwait_stage[lane - 8] * a.n_groups + group) < (unsigned)t)
          __nanosleep(100);
      __syncwarp();
    }
    GNode nxt = a.nodes[sg.node_begin];
    for (int i = sg.node_begin; i < sg.node_end; ++i)
    {
      const GNode nd = nxt;
      if (i + 1 < sg.node_end) nxt = a.nodes[i + 1];  // descriptor of the next node is in flight while this one runs
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
        case MLB_OP_HALFBAND_UP_2:
        case MLB_OP_FDN8_R: break;  // PARAM is an operand kind; FDN8_R / HALFBAND_UP_2 alias their producer's 2nd slot
        case MLB_OP_HALFBAND_UP:
          run_halfband_up_node<EX>(nd, a, v, live, r[0], o, rows + (uint32_t)nd.out_slot2 * kSlotBytes);
          break;
        case MLB_OP_HALFBAND_DOWN: run_halfband_down_node<EX>(nd, a, v, live, r[0], r[1], o); break;
        case MLB_OP_TEMPO_LOCK: run_tempo_lock_node<EX>(nd, a, v, live, r[0], r[1], o); break;
        case MLB_OP_DOWN2X_IN: run_down2x_in_node<EX>(nd, a, v, live, r[0], o); break;
        case MLB_OP_DOWN2X_OUT: run_down2x_out_node<EX>(nd, a, v, live, r[0], o); break;
        case MLB_OP_IMPORT_ROW:
        {
          const float4* src = reinterpret_cast<const float4*>(
              a.chan + (((size_t)t * a.n_chan + nd.iarg) * a.V + (live ? v : 0)) * MLB_BLOCK);
#pragma unroll
          for (int h = 0; h < 2; ++h)
          {
            float4 buf[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) buf[q] = live ? __ldcg(src + h * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) sts128(o + (uint32_t)(h * 8 + q) * 16u, buf[q]);
          }
          break;
        }
        case MLB_OP_INPUT:
        {
          const float4* src = reinterpret_cast<const float4*>(
              a.in + (((size_t)t * a.n_in + nd.iarg) * a.V + (live ? v : 0)) * MLB_BLOCK);
#pragma unroll
          for (int h = 0; h < 2; ++h)
          {
            float4 buf[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) buf[q] = live ? __ldg(src + h * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) sts128(o + (uint32_t)(h * 8 + q) * 16u, buf[q]);
          }
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
          MLB_FUN_CASE(MLB_OP_IMPULSE)
          MLB_FUN_CASE(MLB_OP_PEAK)
          MLB_FUN_CASE(MLB_OP_RMS)
          MLB_FUN_CASE(MLB_OP_ADSR)
          MLB_FUN_CASE(MLB_OP_ALLPASS1)
          MLB_FUN_CASE(MLB_OP_SAMPLE_GLIDE)
#undef MLB_FUN_CASE
        case MLB_OP_LOPASS_V: run_vfilter_node<MLB_OP_LOPASS_V, EX>(nd, a, v, live, rows, o); break;
        case MLB_OP_LOSHELF_V: run_vfilter_node<MLB_OP_LOSHELF_V, EX>(nd, a, v, live, rows, o); break;
        case MLB_OP_HISHELF_V: run_vfilter_node<MLB_OP_HISHELF_V, EX>(nd, a, v, live, rows, o); break;
        case MLB_OP_LOPASS_MOD: run_vfilter_node<MLB_OP_LOPASS_MOD, EX>(nd, a, v, live, rows, o); break;
        case MLB_OP_RAMP: run_ramp_node<EX>(r[0], r[1], o); break;
        case MLB_OP_GLIDE: run_glide_node<EX>(nd, a, v, live, r[0], o); break;
        case MLB_OP_INTERPOLATOR1: run_interp1_node<EX>(nd, a, v, live, r[0], o); break;
        case MLB_OP_INTEGER_DELAY: run_int_delay_node<EX>(nd, a, v, live, t, r[0], o); break;
        case MLB_OP_INTEGER_DELAY_VAR: run_delay_var_node<EX, false>(nd, a, v, live, t, r[0], r[1], o); break;
        case MLB_OP_FRACTIONAL_DELAY: run_frac_delay_node<EX>(nd, a, v, live, t, r[0], o); break;
        case MLB_OP_FRACTIONAL_DELAY_VAR: run_delay_var_node<EX, true>(nd, a, v, live, t, r[0], r[1], o); break;
        case MLB_OP_PITCHBEND_DELAY: run_pitchbend_node<EX>(nd, a, v, live, t, r[0], r[1], o, sB, sC); break;
        case MLB_OP_ALLPASS_INT: run_allpass_node<MLB_OP_ALLPASS_INT, EX>(nd, a, v, live, t, r[0], r[1], o, sA, sB, sC); break;
        case MLB_OP_ALLPASS_FRAC: run_allpass_node<MLB_OP_ALLPASS_FRAC, EX>(nd, a, v, live, t, r[0], r[1], o, sA, sB, sC); break;
        case MLB_OP_ALLPASS_PB: run_allpass_node<MLB_OP_ALLPASS_PB, EX>(nd, a, v, live, t, r[0], r[1], o, sA, sB, sC); break;
        case MLB_OP_FEEDBACK_READ:
          if (live)
          {
            // the row may have been written by another CTA (the FEEDBACK_WRITE's stage): L2 loads
            const float4* s4 = reinterpret_cast<const float4*>(node_row(nd, a, v));
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
              float4 buf[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) buf[q] = __ldcg(s4 + h * 8 + q);
#pragma unroll
              for (int q = 0; q < 8; ++q) sts128(o + (uint32_t)(h * 8 + q) * 16u, buf[q]);
            }
          }
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

      if (nd.chan_out >= 0 && live)
      {
        float4* dst = reinterpret_cast<float4*>(a.chan + (((size_t)t * a.n_chan + nd.chan_out) * a.V + v) * MLB_BLOCK);
#pragma unroll 4
        for (int q = 0; q < 16; ++q) __stcg(dst + q, lds128(o + (uint32_t)q * 16u));
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
    if (a.n_stages > 1)
    {
      // publish block t of this stage: channel rows, feedback rows and delay memory written above
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_u32(progress + (size_t)stage * a.n_groups + group, (unsigned)(t + 1));
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
