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
    if constexpr (nst == 0 && nco == 0 && nin >= 1 && id != MLB_OP_FDN8_R) \
      run_stateless_node<id, EX>(x, b, c, out_addr);                       \
    break;
    MLB_OP_TABLE_STATELESS(MLB_X_STATELESS)
#undef MLB_X_STATELESS
    default: break;
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
    if constexpr (nst == 0 && nco == 0 && nin >= 1 && id != MLB_OP_FDN8_R) \
      return map_kernel<id, true>;                                         \
    break;
    MLB_OP_TABLE_STATELESS(MLB_X_MAP)
#undef MLB_X_MAP
    default: break;
  }
  return nullptr;
}

}  // namespace mlb
