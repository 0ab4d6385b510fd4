// mlb200.cu -- host side of the C ABI declared in include/mlb200.h.
//
// Owns device memory for per-voice state / coefficients / delay lines, recognises linear
// chains that have a fused sm_100a kernel (chain_kernel.cuh), builds the TMA tensor maps
// and launches ONE kernel per process call.  Everything else goes through the graph
// interpreter kernel (generic_kernel.cuh).  There is no CPU compute path in this library:
// without an sm_100 device every processing entry point returns MLB_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mlb200.h"
#include "chain_kernel.cuh"
#include "fdn_kernel.cuh"
#include "voice_kernel.cuh"
#include "resample_kernel.cuh"
#include "generic_kernel.cuh"

using namespace mlb;

// ------------------------------------------------------------------------------------------
// errors, counters

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU_CHECK(expr)                                                                      \
  do                                                                                        \
  {                                                                                         \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(MLB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                  __FILE__, __LINE__);                                                      \
  } while (0)

extern "C" const char* mlb_last_error(void) { return g_err; }
extern "C" int mlb_abi_version(void) { return MLB_ABI_VERSION; }
extern "C" long long mlb_kernel_launches(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------
// op metadata and graph layout (host only)

extern "C" int mlb_op_info(int op, int* n_in, int* n_state, int* n_coef)
{
  switch (op)
  {
#define MLB_X_INFO(NAME, id, nin, nst, nco) \
  case id:                                  \
    if (n_in) *n_in = nin;                  \
    if (n_state) *n_state = nst;            \
    if (n_coef) *n_coef = nco;              \
    return MLB_OK;
    MLB_OP_TABLE(MLB_X_INFO)
#undef MLB_X_INFO
  }
  return MLB_ERR_INVALID;
}

extern "C" const char* mlb_op_name(int op)
{
  switch (op)
  {
#define MLB_X_NAME(NAME, id, nin, nst, nco) \
  case id: return #NAME;
    MLB_OP_TABLE(MLB_X_NAME)
#undef MLB_X_NAME
  }
  return "?";
}

// delay memory of an op: 64-float member rows and IntegerDelay rings per voice (MLB_OP_MEM_TABLE)
static void op_mem(int op, int* n_rows, int* n_rings);

// MLB_AGAIN (mlb200.h): may a node of this op be a further call of an earlier functor?
static bool again_allowed(int op)
{
  int nin = 0, nst = 0, nco = 0, rows = 0, rings = 0;
  if (mlb_op_info(op, &nin, &nst, &nco) != MLB_OK) return false;
  op_mem(op, &rows, &rings);
  if (rings != 0 || (nst == 0 && nco == 0)) return false;
  switch (op)
  {
    case MLB_OP_INPUT: case MLB_OP_PARAM: case MLB_OP_FEEDBACK_READ: case MLB_OP_FEEDBACK_WRITE: case MLB_OP_FDN8:
    case MLB_OP_FDN8_R: case MLB_OP_HALFBAND_UP_2: case MLB_OP_DOWN2X_IN: case MLB_OP_DOWN2X_OUT:
      return false;  // (a HALFBAND_UP called again has its own HALFBAND_UP_2 reading ITS second row)
    default: return true;
  }
}
// the node whose functor node i calls again, or -1
static int again_target(const mlb_node* nodes, int i)
{
  const int op = nodes[i].op;
  if (op == MLB_OP_INPUT || op == MLB_OP_PARAM || op == MLB_OP_FEEDBACK_WRITE) return -1;  // iarg has its own meaning / none
  return nodes[i].iarg < 0 ? MLB_AGAIN_TARGET(nodes[i].iarg) : -1;
}

extern "C" int mlb_graph_layout(const mlb_node* nodes, int n_nodes, mlb_layout* layout,
                                int32_t* state_off, int32_t* coef_off)
{
  if (!nodes || n_nodes <= 0 || !layout) return fail(MLB_ERR_INVALID, "null graph");
  int ns = 0, nc = 0, nin_planes = 0, n_fdn = 0;
  for (int i = 0; i < n_nodes; ++i)
  {
    int nin, nst, nco;
    if (mlb_op_info(nodes[i].op, &nin, &nst, &nco) != MLB_OK)
      return fail(MLB_ERR_INVALID, "node %d: unknown op %d", i, nodes[i].op);
    for (int k = 0; k < MLB_MAX_INS; ++k)
    {
      const int src = nodes[i].in[k];
      if (k < nin)
      {
        if (src < 0 || src >= i)
          return fail(MLB_ERR_INVALID, "node %d (%s): input %d must be an earlier node", i,
                      mlb_op_name(nodes[i].op), k);
      }
      else if (src != -1)
        return fail(MLB_ERR_INVALID, "node %d (%s): unused input %d must be -1", i,
                    mlb_op_name(nodes[i].op), k);
    }
    if (nodes[i].op == MLB_OP_INPUT)
    {
      if (nodes[i].iarg < 0 || nodes[i].iarg >= 64)
        return fail(MLB_ERR_INVALID, "node %d: INPUT plane %d out of range", i, nodes[i].iarg);
      nin_planes = std::max(nin_planes, nodes[i].iarg + 1);
    }
    if (nodes[i].op == MLB_OP_FDN8_R && nodes[nodes[i].in[0]].op != MLB_OP_FDN8)
      return fail(MLB_ERR_INVALID, "node %d: FDN8_R must read an FDN8 node", i);
    if (nodes[i].op == MLB_OP_HALFBAND_UP_2 && nodes[nodes[i].in[0]].op != MLB_OP_HALFBAND_UP)
      return fail(MLB_ERR_INVALID, "node %d: HALFBAND_UP_2 must read a HALFBAND_UP node", i);
    if (nodes[i].op == MLB_OP_FDN8) ++n_fdn;
    if (nodes[i].op == MLB_OP_FEEDBACK_WRITE)
    {
      const int rd = nodes[i].iarg;
      if (rd < 0 || rd >= i || nodes[rd].op != MLB_OP_FEEDBACK_READ)
        return fail(MLB_ERR_INVALID, "node %d: FEEDBACK_WRITE.iarg must name an earlier FEEDBACK_READ node", i);
      for (int q = 0; q < i; ++q)
        if (nodes[q].op == MLB_OP_FEEDBACK_WRITE && nodes[q].iarg == rd)
          return fail(MLB_ERR_INVALID, "node %d: FEEDBACK_READ %d already has a FEEDBACK_WRITE (node %d)", i, rd, q);
    }
    const int again = again_target(nodes, i);
    if (again >= 0)
    {
      if (!again_allowed(nodes[i].op))
        return fail(MLB_ERR_INVALID, "node %d (%s): iarg %d -- this op cannot be called again in a vector (MLB_AGAIN)", i,
                    mlb_op_name(nodes[i].op), nodes[i].iarg);
      if (again >= i || nodes[again].op != nodes[i].op || again_target(nodes, again) >= 0)
        return fail(MLB_ERR_INVALID, "node %d (%s): MLB_AGAIN(%d) must name an earlier node of the same op that is not "
                    "itself an AGAIN node", i, mlb_op_name(nodes[i].op), again);
      // a further call of node `again`: its words, no words of its own
      if (state_off) state_off[i] = state_off[again];
      if (coef_off) coef_off[i] = coef_off[again];
      continue;
    }
    if (state_off) state_off[i] = ns;
    if (coef_off) coef_off[i] = nc;
    ns += nst;
    nc += nco;
  }
  if (n_fdn > 1) return fail(MLB_ERR_UNSUPPORTED, "at most one FDN8 node per graph");
  layout->n_state_words = ns;
  layout->n_coef_words = nc;
  layout->n_inputs = nin_planes;
  return MLB_OK;
}

// ------------------------------------------------------------------------------------------
// coefficient design on the host with glibc libm -- the same calls the reference makes, so
// coefficients are bit-identical by construction (SURVEY 8c "third-party arithmetic").

static const float kPiF = 3.1415926535897932384626433f;     // ScalarMath.h:24
static const float kTwoPiF = 6.2831853071795864769252867f;  // ScalarMath.h:23

static void svf_g(float omega, float k, float* g0, float* g1, float* g2)
{
  volatile float piOmega = kPiF * omega;
  volatile float s1 = sinf(piOmega);
  volatile float s2 = sinf(2.0f * piOmega);
  volatile float nrm = 1.0f / (2.f + k * s2);
  *g0 = s2 * nrm;
  volatile float a = -2.f * s1 * s1;
  volatile float b = k * s2;
  volatile float c = a - b;
  *g1 = c * nrm;
  volatile float d = 2.0f * s1 * s1;
  *g2 = d * nrm;
}
extern "C" void mlb_coeffs_lopass(float omega, float k, float o[3]) { svf_g(omega, k, &o[0], &o[1], &o[2]); }
extern "C" void mlb_coeffs_hipass(float omega, float k, float o[4])
{
  svf_g(omega, k, &o[0], &o[1], &o[2]);
  o[3] = k;
}
extern "C" void mlb_coeffs_bandpass(float omega, float k, float o[3]) { svf_g(omega, k, &o[0], &o[1], &o[2]); }
extern "C" void mlb_coeffs_loshelf(float omega, float k, float A, float r[5])
{
  volatile float piOmega = kPiF * omega;
  volatile float g = tanf(piOmega) / sqrtf(A);
  volatile float t0 = g + k;
  volatile float t1 = g * t0;
  volatile float t2 = 1.f + t1;
  r[0] = 1.f / t2;
  r[1] = g * r[0];
  r[2] = g * r[1];
  volatile float am1 = A - 1.f;
  r[3] = k * am1;
  volatile float aa = A * A;
  r[4] = aa - 1.f;
}
extern "C" void mlb_coeffs_hishelf(float omega, float k, float A, float r[6])
{
  volatile float piOmega = kPiF * omega;
  volatile float g = tanf(piOmega) * sqrtf(A);
  volatile float t0 = g + k;
  volatile float t1 = g * t0;
  volatile float t2 = 1.f + t1;
  r[0] = 1.f / t2;
  r[1] = g * r[0];
  r[2] = g * r[1];
  r[3] = A * A;
  volatile float oma = 1.f - A;
  volatile float koma = k * oma;
  r[4] = koma * A;
  volatile float aa = A * A;
  r[5] = 1.f - aa;
}
extern "C" void mlb_coeffs_bell(float omega, float k, float A, float r[4])
{
  volatile float kc = k / A;
  volatile float piOmega = kPiF * omega;
  volatile float g = tanf(piOmega);
  volatile float t0 = g + kc;
  volatile float t1 = g * t0;
  volatile float t2 = 1.f + t1;
  volatile float a1 = 1.f / t2;
  volatile float a2 = g * a1;
  volatile float a3 = g * a2;
  volatile float aa = A * A;
  volatile float aam1 = aa - 1.f;
  r[0] = a1, r[1] = a2, r[2] = a3, r[3] = kc * aam1;
}
extern "C" void mlb_coeffs_onepole(float omega, float r[2])
{
  volatile float arg = -omega * kTwoPiF;
  volatile float x = expf(arg);
  r[0] = 1.f - x;
  r[1] = x;
}
extern "C" float mlb_coeffs_dcblocker(float omega) { return cosf(omega); }
extern "C" float mlb_db_to_gain(float dB)
{
  volatile float e = dB / 40.f;
  return powf(10.f, e);
}
extern "C" void mlb_coeffs_peak(float omega, float r[2]) { mlb_coeffs_onepole(omega, r); }
extern "C" void mlb_coeffs_rms(float omega, float r[2]) { mlb_coeffs_onepole(omega, r); }
extern "C" void mlb_coeffs_adsr(float a, float d, float s, float rel, float sr, float o[4])
{
  const float minSegmentTime = 0.0002f;
  volatile float invSr = 1.0f / sr;
  volatile float num = kTwoPiF * invSr;
  o[0] = num / (a > minSegmentTime ? a : minSegmentTime);
  o[1] = num / (d > minSegmentTime ? d : minSegmentTime);
  o[2] = s;
  o[3] = num / (rel > minSegmentTime ? rel : minSegmentTime);
}
extern "C" float mlb_coeffs_allpass1(float d)
{
  volatile float xm1 = d - 1.f;
  volatile float t0 = -0.53f * xm1;
  volatile float t1 = 0.24f * xm1;
  volatile float t2 = t1 * xm1;
  return t0 + t2;
}
// ImpulseGen's constructor, G:64-78 (window: MLDSPUtils.h:22-35; normalize / sum: MLDSPOps.h:995-1049 with
// vecSumH's (v0 + v2) + (v1 + v3) order, MLDSPMathSSE.h:246-251)
extern "C" void mlb_impulse_table(float t17[17])
{
  float row[MLB_BLOCK];
  memset(row, 0, sizeof(row));
  for (int i = 0; i < 17; ++i)
  {
    volatile float m = (1.f - 0.f) / (16.f - 0.f);
    volatile float x = m * ((float)i - 0.f) + 0.f;
    volatile float c1 = cosf(kTwoPiF * x);
    volatile float c2 = cosf(2.f * kTwoPiF * x);
    volatile float w0 = 0.5f * c1;
    volatile float w1 = 0.42f - w0;
    volatile float w2 = 0.08f * c2;
    volatile float w = w1 + w2;
    const int idx = i - 8;
    volatile float pi_x = kTwoPiF * 0.25f * (float)idx;
    volatile float sinc = (idx == 0) ? 1.f : sinf(pi_x) / pi_x;
    row[i] = sinc * w;
  }
  volatile float sum = 0.f;
  for (int n = 0; n < MLB_BLOCK; n += 4)
  {
    volatile float a = row[n] + row[n + 2];
    volatile float b = row[n + 1] + row[n + 3];
    volatile float ab = a + b;
    sum = sum + ab;
  }
  for (int i = 0; i < 17; ++i) t17[i] = row[i] / sum;
}
extern "C" void mlb_coeffs_glide(float t, float o[2])
{
  int n = (int)(t / (float)MLB_BLOCK);
  if (n < 1) n = 1;
  o[0] = (float)n;
  o[1] = 1.0f / ((float)n + 0.f);
}
extern "C" void mlb_coeffs_sample_glide(float t, float o[2])
{
  int n = (int)t;
  if (n < 1) n = 1;
  o[0] = (float)n;
  o[1] = 1.0f / (float)n;
}
// Lopass::makeCoeffsVec, F:97-115 (clamps by _mm_min_ps / _mm_max_ps operand order, then per-sample makeCoeffs)
extern "C" void mlb_coeffs_lopass_vec(const float omega[64], const float k[64], float out[3 * 64])
{
  for (int n = 0; n < MLB_BLOCK; ++n)
  {
    const float om = omega[n] < 0.5f ? omega[n] : 0.5f;  // min(omega, DSPVector(0.5f))
    const float kk = k[n] > 0.01f ? k[n] : 0.01f;        // max(k, DSPVector(0.01f))
    svf_g(om, kk, &out[n], &out[MLB_BLOCK + n], &out[2 * MLB_BLOCK + n]);
  }
}
extern "C" void mlb_coeffs_lopass_vec_n(const float* omega, const float* k, float* out, size_t n_rows)
{
  for (size_t r = 0; r < n_rows; ++r)
    mlb_coeffs_lopass_vec(omega + r * MLB_BLOCK, k + r * MLB_BLOCK, out + r * 3 * MLB_BLOCK);
}
// interpolateCoeffsLinear, F:32-44; interpolateDSPVectorLinear, O:986-990 (columnIndex() * interval + (start + interval))
extern "C" void mlb_interpolate_coeffs_linear(const float* c0, const float* c1, int n_coeffs, float* out)
{
  for (int i = 0; i < n_coeffs; ++i)
  {
    volatile float diff = c1[i] - c0[i];
    volatile float interval = diff / (float)MLB_BLOCK;
    volatile float base = c0[i] + interval;
    for (int n = 0; n < MLB_BLOCK; ++n)
    {
      volatile float ramp = (float)n * interval;
      out[(size_t)i * MLB_BLOCK + n] = ramp + base;
    }
  }
}

// n voices at once: out[word][n] (the SoA layout of mlb_graph_set_coefs).  kind = the node op.
extern "C" int mlb_coeffs_batch(int op, size_t n, const float* omega, const float* k, const float* A, float* out)
{
  if (!omega || !out) return fail(MLB_ERR_INVALID, "null argument");
  int nco = 0;
  switch (op)
  {
    case MLB_OP_LOPASS: case MLB_OP_BANDPASS: nco = 3; break;
    case MLB_OP_HIPASS: case MLB_OP_BELL: nco = 4; break;
    case MLB_OP_LOSHELF: nco = 5; break;
    case MLB_OP_HISHELF: nco = 6; break;
    case MLB_OP_ONEPOLE: nco = 2; break;
    default: return fail(MLB_ERR_INVALID, "mlb_coeffs_batch: op %d has no omega/k/A coefficient design", op);
  }
  if (op != MLB_OP_ONEPOLE && !k) return fail(MLB_ERR_INVALID, "k is null");
  if ((op == MLB_OP_BELL || op == MLB_OP_LOSHELF || op == MLB_OP_HISHELF) && !A) return fail(MLB_ERR_INVALID, "A is null");
  for (size_t v = 0; v < n; ++v)
  {
    float c[6];
    switch (op)
    {
      case MLB_OP_LOPASS: mlb_coeffs_lopass(omega[v], k[v], c); break;
      case MLB_OP_BANDPASS: mlb_coeffs_bandpass(omega[v], k[v], c); break;
      case MLB_OP_HIPASS: mlb_coeffs_hipass(omega[v], k[v], c); break;
      case MLB_OP_BELL: mlb_coeffs_bell(omega[v], k[v], A[v], c); break;
      case MLB_OP_LOSHELF: mlb_coeffs_loshelf(omega[v], k[v], A[v], c); break;
      case MLB_OP_HISHELF: mlb_coeffs_hishelf(omega[v], k[v], A[v], c); break;
      default: mlb_coeffs_onepole(omega[v], c); break;
    }
    for (int w = 0; w < nco; ++w) out[(size_t)w * n + v] = c[w];
  }
  return MLB_OK;
}
extern "C" void mlb_coeffs_fdn8(const float times[8], const float cutoffs[8], const float gains[8],
                                float out32[32])
{
  for (int n = 0; n < 8; ++n)
  {
    float c[2];
    mlb_coeffs_onepole(cutoffs[n], c);
    out32[n] = c[0];
    out32[8 + n] = c[1];
    out32[16 + n] = gains[n];
    // FDN::setDelaysInSamples, F:1173-1183: int len = times[n] - 64; len = max(1, len)
    int len = (int)(times[n] - (float)MLB_BLOCK);
    out32[24 + n] = (float)std::max(1, len);
  }
}

// ------------------------------------------------------------------------------------------
// device

static int g_device = -1;
static std::atomic<int> g_live_handles{0};  // graphs + voice banks + resamplers alive (they hold pointers of g_device)
static int g_sm_count = 0;
static size_t g_smem_optin = 0;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

extern "C" int mlb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int mlb_init(int device)
{
  int n = mlb_device_count();
  if (n <= 0)
    return fail(MLB_ERR_NO_DEVICE,
                "no CUDA device visible: this library has no CPU fallback (sm_100a kernels only)");
  if (device < 0 || device >= n) return fail(MLB_ERR_INVALID, "device %d out of range (%d)", device, n);
  // one device per process (one process per GPU is the multi-GPU model): retargeting while
  // handles of the current device are alive would leave them with foreign pointers
  if (g_device >= 0 && device != g_device && g_live_handles.load() > 0)
    return fail(MLB_ERR_INVALID, "mlb_init(%d): %d handle(s) of device %d are still alive (one device per process)",
                device, g_live_handles.load(), g_device);
  CU_CHECK(cudaSetDevice(device));
  cudaDeviceProp p;
  CU_CHECK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return fail(MLB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library ships sm_100a code only",
                device, p.major, p.minor);
  g_sm_count = p.multiProcessorCount;
  g_smem_optin = p.sharedMemPerBlockOptin;
  if (!g_encode)
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn)
      return fail(MLB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
  }
  {
    float table[17];
    mlb_impulse_table(table);  // ImpulseGen's windowed sinc, host libm
    CU_CHECK(cudaMemcpyToSymbol(c_impulse_table, table, sizeof(table)));
  }
  // mix_reduce_kernel runs between two chain kernels that need the maximum shared-memory carve-out: keep the
  // SMs in that configuration instead of switching the L1 / shared split twice per step
  cudaFuncSetAttribute((const void*)mix_reduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                       (int)cudaSharedmemCarveoutMaxShared);
  g_device = device;
  return MLB_OK;
}

static int ensure_init()
{
  if (g_device >= 0)
  {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != g_device) cudaSetDevice(g_device);
    return MLB_OK;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess)
  {
    cudaGetLastError();
    dev = 0;
  }
  return mlb_init(dev);
}

// Tensor map over planes of [V][64] f32 that moves FULL 256-byte rows yet lands as two
// 128B-swizzled half tiles: dims {32 samples, V voices, 2 halves, n_planes} with strides
// {256 B, 128 B, plane} and box {32, 32, 2, 1} (8 KB per TMA op).
static int make_block_map(CUtensorMap* map, const float* base, int V, long long n_planes,
                          long long plane_stride_floats)
{
  cuuint64_t dims[4] = {(cuuint64_t)kTileSamples, (cuuint64_t)V, 2, (cuuint64_t)n_planes};
  cuuint64_t strides[3] = {(cuuint64_t)MLB_BLOCK * 4, (cuuint64_t)kTileSamples * 4,
                           (cuuint64_t)plane_stride_floats * 4};
  cuuint32_t box[4] = {(cuuint32_t)kTileSamples, (cuuint32_t)kTileVoices, 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MLB_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
  return MLB_OK;
}

// Encoded tensor maps are pure functions of (base, V, planes, plane stride): a graph keeps the few it has
// built (an audio callback re-uses the same buffers call after call), so a steady-state process call does
// no driver work besides the launches.
struct MapKey
{
  const void* base;
  int V;
  long long n_planes, stride;
  bool operator==(const MapKey& o) const { return base == o.base && V == o.V && n_planes == o.n_planes && stride == o.stride; }
};
struct MapCache
{
  static constexpr int kEntries = 40;  // 2 maps x 16 host slices + device-path buffers
  MapKey key[kEntries];
  CUtensorMap map[kEntries];
  int n = 0, next = 0;
  long long hits = 0, misses = 0;
};
static int cached_block_map(MapCache* c, CUtensorMap* out, const float* base, int V, long long n_planes,
                            long long plane_stride_floats)
{
  const MapKey k{base, V, n_planes, plane_stride_floats};
  for (int i = 0; i < c->n; ++i)
    if (c->key[i] == k)
    {
      *out = c->map[i];
      ++c->hits;
      return MLB_OK;
    }
  int rc = make_block_map(out, base, V, n_planes, plane_stride_floats);
  if (rc != MLB_OK) return rc;
  const int slot = c->n < MapCache::kEntries ? c->n++ : (c->next++ % MapCache::kEntries);
  c->key[slot] = k;
  c->map[slot] = *out;
  ++c->misses;
  return MLB_OK;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and size, not once per launch
static int ensure_func_smem(const void* fn, size_t smem)
{
  static std::mutex mu;
  static std::vector<std::pair<const void*, size_t>> seen;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : seen)
    if (e.first == fn)
    {
      if (e.second >= smem) return MLB_OK;
      CU_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      e.second = smem;
      return MLB_OK;
    }
  CU_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // every kernel of this library wants the SM's L1 / shared split at "maximum shared": the rings of several CTAs
  // (or one 197-KB CTA) per SM set the occupancy, and kernels that follow each other then never make the SM switch
  CU_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  seen.emplace_back(fn, smem);
  return MLB_OK;
}

// ------------------------------------------------------------------------------------------
// fused chain registry

typedef void (*ChainKernelFn)(const CUtensorMap, const CUtensorMap, const ChainArgs);
struct FusedEntry
{
  int gen, src, f1, f2, gain, exact;
  ChainKernelFn fn;
  const char* name;
  bool has_in;
  int n_planes;  // 8-KB planes a ring stage holds (Chain::NP): the input row and/or coefficient rows
  ChainKernelFn team_fn;  // two-warp team variant for small banks (chain_team_kernel), or nullptr
};
template <class P>
static constexpr ChainKernelFn team_fn_of()
{
  if constexpr (P::SPLIT)
    return chain_team_kernel<P>;
  else
    return nullptr;
}

#define N_ (-1)
// (GEN, SRC, F1, F2, GAIN)
#define MLB_FUSED_LIST(X)                                                   \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_LOPASS, N_, true)   /* config 1 / A */   \
  X(MLB_OP_SINE, SRC_PARAM, MLB_OP_LOPASS, N_, true)   /* contract S / M */ \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_LOPASS, N_, false)  /* config 2 */       \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_HIPASS, N_, false)                       \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_BANDPASS, N_, false)                     \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_LOSHELF, N_, false)                      \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_HISHELF, N_, false)                      \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_BELL, N_, false)                         \
  X(MLB_OP_PHASOR, SRC_INPUT, MLB_OP_LOPASS, MLB_OP_ONEPOLE, false) /* config 3 */ \
  X(MLB_OP_SINE, SRC_INPUT, N_, N_, false)                                  \
  X(MLB_OP_PHASOR, SRC_INPUT, N_, N_, false)                                \
  X(MLB_OP_SAW, SRC_INPUT, N_, N_, false)                                   \
  X(MLB_OP_NOISE, SRC_NONE, N_, N_, false)                                  \
  X(MLB_OP_NOISE, SRC_NONE, MLB_OP_LOPASS, N_, true)                        \
  X(N_, SRC_INPUT, MLB_OP_LOPASS, N_, false)                                \
  X(N_, SRC_INPUT, MLB_OP_ONEPOLE, N_, false)                               \
  X(N_, SRC_INPUT, MLB_OP_DCBLOCKER, N_, false)                             \
  /* swept filter: coefficient ROWS as three more input planes (20 B per voice-sample) */ \
  X(MLB_OP_SINE, SRC_INPUT, MLB_OP_LOPASS_V, N_, true)                      \
  X(MLB_OP_SINE, SRC_PARAM, MLB_OP_LOPASS_V, N_, true)                      \
  X(N_, SRC_INPUT, MLB_OP_LOPASS_V, N_, false)

static const FusedEntry g_fused[] = {
#define MLB_X_ENTRY(G, S, F1, F2, GN)                                                          \
  {G, S, F1, F2, GN, 1, chain_kernel<Chain<G, S, F1, F2, GN, true>>, #G "+" #F1 "+" #F2 "+" #GN, \
   S == SRC_INPUT, Chain<G, S, F1, F2, GN, true>::NP, team_fn_of<Chain<G, S, F1, F2, GN, true>>()}, \
      {G, S, F1, F2, GN, 0, chain_kernel<Chain<G, S, F1, F2, GN, false>>,                        \
       #G "+" #F1 "+" #F2 "+" #GN "(fast)", S == SRC_INPUT, Chain<G, S, F1, F2, GN, false>::NP,    \
       team_fn_of<Chain<G, S, F1, F2, GN, false>>()},
    MLB_FUSED_LIST(MLB_X_ENTRY)
#undef MLB_X_ENTRY
};

// ------------------------------------------------------------------------------------------
// graph object

enum GraphKind
{
  KIND_FUSED = 0,
  KIND_FDN = 1,
  KIND_GENERIC = 2
};

struct mlb_graph
{
  std::vector<mlb_node> nodes;
  std::vector<int32_t> outs;
  std::vector<int32_t> st_off, co_off;
  mlb_layout layout{};
  int V = 0;
  unsigned flags = 0;
  bool exact = true;
  int kind = KIND_GENERIC;
  std::string kernel_name;

  // device SoA
  uint32_t* d_state = nullptr;
  float* d_coef = nullptr;
  std::vector<float> h_coef;  // host mirror (ring sizing, FDN descriptors)

  // fused chain
  const FusedEntry* fused = nullptr;
  ChainArgs cargs{};

  // FM3 -> FDN8 specialisation
  FdnArgs fargs{};

  // generic interpreter
  std::vector<GNode> gnodes;   // stage programs back to back (imports + nodes)
  GNode* d_gnodes = nullptr;
  int scratch_slot = 0;
  std::vector<GStage> gstages;
  GStage* d_gstages = nullptr;
  int n_stages = 1, n_chan = 0;
  unsigned* d_gsync = nullptr;  // [0] ticket, [1 + s*G + g] blocks finished by stage s of group g
  float* d_chan = nullptr;      // channel planes [T][n_chan][V][64]
  size_t chan_cap = 0;

  int n_slots = 0;

  // FDN delay memory
  int fdn_node = -1;
  float* d_ring = nullptr;
  float* d_carry = nullptr;
  int ring_len = 0;
  long long blocks_done = 0;
  int reserved_sms = 0;
  // delay memory of the section-8(f) functors: member rows [V][64] and rings [V][stride] per node
  bool has_dmem = false;
  float* d_dmem = nullptr;
  size_t dmem_floats = 0;
  std::vector<unsigned> ring_stride;  // per node, 0 = no ring

  // chain scheduler words, all maintained by the kernels themselves (see ChainArgs)
  unsigned* d_sched = nullptr;
  int launch_chunks = 1;  // chunks per group of the most recent launch (all slices of one host call agree)

  // mix bus partials
  float* d_partial = nullptr;
  size_t partial_cap = 0;

  // staging for the host entry point
  float *d_in = nullptr, *d_out = nullptr, *d_mix = nullptr;
  size_t in_cap = 0, out_cap = 0, mix_cap = 0;
  cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  static constexpr int kMaxHostSlices = 64;
  cudaEvent_t ev_up[kMaxHostSlices] = {}, ev_k[kMaxHostSlices] = {};

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  int last_host_slices = 0;  // voice slices of the most recent mlb_graph_process_host call (1 = one launch)
  MapCache maps;
  // asynchronous mix bus (mlb_graph_set_mix_async): mix_reduce_kernel (and the multi-GPU exchange) run on s_mix,
  // beside the next call's chain kernel; the per-group partials are double-buffered by call parity
  int mix_async = 0;
  cudaStream_t s_mix = nullptr;
  cudaEvent_t ev_main = nullptr, ev_mixdone[4] = {};
  unsigned mix_seq = 0;
  bool mix_pending = false;
  float* partial_cur = nullptr;  // the partial buffer of the call being issued
  struct mlb_mixbus* bus = nullptr;  // multi-GPU: the mix bus is all-reduced over peer memory inside mix_reduce_kernel
};

// delay memory of an op: 64-float member rows and IntegerDelay rings per voice (MLB_OP_MEM_TABLE)
static void op_mem(int op, int* n_rows, int* n_rings)
{
  *n_rows = *n_rings = 0;
  switch (op)
  {
#define MLB_X_MEM(NAME, rows, rings) \
  case MLB_OP_##NAME: *n_rows = rows, *n_rings = rings; break;
    MLB_OP_MEM_TABLE(MLB_X_MEM)
#undef MLB_X_MEM
    default: break;
  }
}

static bool is_filter(int op)
{
  switch (op)
  {
    case MLB_OP_LOPASS:
    case MLB_OP_HIPASS:
    case MLB_OP_BANDPASS:
    case MLB_OP_LOSHELF:
    case MLB_OP_HISHELF:
    case MLB_OP_BELL:
    case MLB_OP_ONEPOLE:
    case MLB_OP_DCBLOCKER:
    case MLB_OP_DIFFERENTIATOR:
    case MLB_OP_INTEGRATOR: return true;
  }
  return false;
}
static bool is_vfilter(int op) { return op == MLB_OP_LOPASS_V || op == MLB_OP_LOSHELF_V || op == MLB_OP_HISHELF_V; }
static bool is_gen1(int op)
{
  return op == MLB_OP_SINE || op == MLB_OP_PHASOR || op == MLB_OP_SAW || op == MLB_OP_TICK;
}

// Recognise  out = [gain *] F2( F1( GEN(src) ) )  and fill g->cargs index maps.
static bool has_again_nodes(const mlb_graph* g);
static bool match_fused_chain(mlb_graph* g)
{
  if (has_again_nodes(g)) return false;  // a functor called twice per vector: interpreter only
  if (g->outs.size() != 1) return false;
  const auto& N = g->nodes;
  std::vector<char> used(N.size(), 0);
  int y = g->outs[0];
  int gain_node = -1;
  if (N[y].op == MLB_OP_MULTIPLY)
  {
    const int a = N[y].in[0], b = N[y].in[1];
    used[y] = 1;
    if (N[b].op == MLB_OP_PARAM && N[a].op != MLB_OP_PARAM)
      gain_node = b, y = a;
    else if (N[a].op == MLB_OP_PARAM && N[b].op != MLB_OP_PARAM)
      gain_node = a, y = b;
    else
      return false;
    used[gain_node] = 1;
  }
  int filt[2] = {-1, -1};
  int nf = 0;
  while (is_filter(N[y].op) || is_vfilter(N[y].op))
  {
    if (nf == 2) return false;
    filt[nf++] = y;
    used[y] = 1;
    y = N[y].in[0];
  }
  // filt[] is output-side first: reverse into signal order
  int f1 = -1, f2 = -1;
  if (nf == 1) f1 = filt[0];
  if (nf == 2) f1 = filt[1], f2 = filt[0];
  int gen = -1, src = SRC_NONE, src_node = -1, gen_node = -1;
  if (is_gen1(N[y].op))
  {
    gen = N[y].op, gen_node = y;
    used[y] = 1;
    src_node = N[y].in[0];
  }
  else if (N[y].op == MLB_OP_NOISE)
  {
    gen = MLB_OP_NOISE, gen_node = y;
    used[y] = 1;
  }
  else
    src_node = y;
  if (src_node >= 0)
  {
    used[src_node] = 1;
    if (N[src_node].op == MLB_OP_INPUT)
      src = SRC_INPUT;
    else if (N[src_node].op == MLB_OP_PARAM)
      src = SRC_PARAM;
    else
      return false;
  }
  // a coefficient-row filter is fused only as F1, with every coefficient row an external INPUT plane
  if (f2 >= 0 && is_vfilter(N[f2].op)) return false;
  if (f1 >= 0 && is_vfilter(N[f1].op))
  {
    int nin = 0;
    mlb_op_info(N[f1].op, &nin, nullptr, nullptr);
    for (int k = 1; k < nin; ++k)
    {
      if (N[N[f1].in[k]].op != MLB_OP_INPUT) return false;
      used[N[f1].in[k]] = 1;
    }
  }
  for (char u : used)
    if (!u) return false;  // every node must be on the chain (unused stateful nodes still tick)
  const int f1op = f1 >= 0 ? N[f1].op : -1, f2op = f2 >= 0 ? N[f2].op : -1;
  for (const FusedEntry& e : g_fused)
  {
    if (e.gen == gen && e.src == src && e.f1 == f1op && e.f2 == f2op &&
        e.gain == (gain_node >= 0 ? 1 : 0) && e.exact == (g->exact ? 1 : 0))
    {
      g->fused = &e;
      ChainArgs& c = g->cargs;
      memset(&c, 0, sizeof(c));
      int si = 0, ci = 0;
      auto add_state = [&](int node)
      {
        int nst = 0;
        mlb_op_info(N[node].op, nullptr, &nst, nullptr);
        for (int k = 0; k < nst; ++k) c.st_idx[si++] = g->st_off[node] + k;
      };
      auto add_coef = [&](int node)
      {
        int nco = 0;
        mlb_op_info(N[node].op, nullptr, nullptr, &nco);
        for (int k = 0; k < nco; ++k) c.co_idx[ci++] = g->co_off[node] + k;
      };
      if (gen_node >= 0) add_state(gen_node);
      if (f1 >= 0) add_state(f1);
      if (f2 >= 0) add_state(f2);
      if (src == SRC_PARAM) add_coef(src_node);
      if (f1 >= 0) add_coef(f1);
      if (f2 >= 0) add_coef(f2);
      if (gain_node >= 0) add_coef(gain_node);
      c.in_plane = (src == SRC_INPUT) ? N[src_node].iarg : 0;
      if (f1 >= 0 && is_vfilter(N[f1].op))
      {
        int nin = 0;
        mlb_op_info(N[f1].op, &nin, nullptr, nullptr);
        for (int k = 1; k < nin; ++k) c.cv_plane[k - 1] = N[N[f1].in[k]].iarg;
      }
      g->kernel_name = std::string("fused:") + e.name;
      return true;
    }
  }
  return false;
}

static bool has_again_nodes(const mlb_graph* g)
{
  for (int i = 0; i < (int)g->nodes.size(); ++i)
    if (again_target(g->nodes.data(), i) >= 0) return true;
  return false;
}
static int env_int(const char* name, int dflt);
// ops with two output rows (the second is read through the paired *_R / *_2 op)
static bool is_dual(int op) { return op == MLB_OP_FDN8 || op == MLB_OP_HALFBAND_UP; }
static bool is_second(int op) { return op == MLB_OP_FDN8_R || op == MLB_OP_HALFBAND_UP_2; }

// Relative cost of one node-block in the interpreter (used only to balance pipeline stages).
static int node_cost(int op)
{
  if (op == MLB_OP_PARAM || op == MLB_OP_INPUT || op == MLB_OP_FDN8_R || op == MLB_OP_HALFBAND_UP_2) return 0;
  if (op >= MLB_OP_MAP_FIRST && op < MLB_OP_MAP_END) return 2;
  switch (op)
  {
    case MLB_OP_ONEPOLE:
    case MLB_OP_DIFFERENTIATOR:
    case MLB_OP_INTEGRATOR:
    case MLB_OP_DCBLOCKER:
    case MLB_OP_RMS:
    case MLB_OP_ALLPASS1: return 8;
    case MLB_OP_FDN8: return 200;
    case MLB_OP_ALLPASS_PB:
    case MLB_OP_PITCHBEND_DELAY: return 40;
    case MLB_OP_ALLPASS_INT:
    case MLB_OP_ALLPASS_FRAC:
    case MLB_OP_FRACTIONAL_DELAY:
    case MLB_OP_FRACTIONAL_DELAY_VAR:
    case MLB_OP_INTEGER_DELAY_VAR: return 20;
    case MLB_OP_FEEDBACK_READ:
    case MLB_OP_FEEDBACK_WRITE:
    case MLB_OP_INTEGER_DELAY:
    case MLB_OP_GLIDE:
    case MLB_OP_INTERPOLATOR1: return 4;
    default: return 16;  // SVF family, generators, envelopes
  }
}

// Build the interpreter program.  The node list is cut into S contiguous *stages*
// (balanced by node_cost); stage s of voice group g is run by its own one-warp CTA for all T
// blocks, so a graph with few voices still fills the machine: stage s works on block t while
// stage s+1 works on block t-1 (a software pipeline across CTAs).  A row that crosses a stage
// boundary travels through a *channel* plane in global memory [T][n_chan][V][64]; per-(stage,
// group) progress words order the producer's stores before the consumer's loads.  Inside a stage
// rows live in shared-memory slots allocated by liveness, as before.  S = 1 is the plain
// single-CTA-per-group interpreter.
static int build_generic(mlb_graph* g)
{
  const auto& N = g->nodes;
  const int n = (int)N.size();
  const int n_groups = (g->V + 31) / 32;
  std::vector<int> out_plane(n, -1);
  for (size_t c = 0; c < g->outs.size(); ++c)
  {
    if (out_plane[g->outs[c]] >= 0)
      return fail(MLB_ERR_UNSUPPORTED, "a node may appear only once in outs");
    out_plane[g->outs[c]] = (int)c;
  }
  // ---- choose the number of stages ----
  long long total_cost = 0;
  int n_real = 0;
  for (int i = 0; i < n; ++i)
  {
    total_cost += node_cost(N[i].op);
    if (node_cost(N[i].op) > 0) ++n_real;
  }
  int S = 1;
  {
    // one warp per scheduler: a second warp on a scheduler halves the pace of both (a lone warp issues every
    // ~2.4 cycles, profiles/dep_latency_r2.txt), and every extra stage costs a block of pipeline fill.  Measured on
    // config 5 (32 groups, T = 16): 8 stages 0.930 ms, 16: 0.768, 24: 0.791, 37: 0.852, 48: 0.914, 64: 1.014
    const int target_warps = g_sm_count * 4;
    S = (target_warps + n_groups - 1) / n_groups;
    S = std::min(S, std::max(1, n_real / 3));
    S = std::min(S, 64);
    S = env_int("MLB_STAGES", S);
    S = std::min(std::max(S, 1), std::max(1, n_real));
    if (g->flags & MLB_GRAPH_SINGLE_STAGE) S = 1;
  }
  // ---- cut points: stage_of[i], contiguous, balanced by cost ----
  // A feedback loop (FEEDBACK_READ ... its FEEDBACK_WRITE) stays inside one stage: the reader of
  // block t+1 would otherwise wait for a CTA with a LARGER ticket, which need not be resident.
  std::vector<char> no_cut(n, 0);
  for (int q = 0; q < n; ++q)
  {
    if (N[q].op == MLB_OP_FEEDBACK_WRITE)
      for (int i = N[q].iarg; i < q; ++i) no_cut[i] = 1;
    if (is_second(N[q].op))  // FDN8_R / HALFBAND_UP_2 alias the second row of their producer: same stage
      for (int i = N[q].in[0]; i < q; ++i) no_cut[i] = 1;
    // MLB_AGAIN: the calls of one functor share its state words, which a node loads and stores once per vector --
    // they must run in node order inside one CTA
    if (again_target(N.data(), q) >= 0)
      for (int i = again_target(N.data(), q); i < q; ++i) no_cut[i] = 1;
  }
  std::vector<int> stage_of(n, 0);
  {
    long long acc = 0;
    int s = 0;
    for (int i = 0; i < n; ++i)
    {
      stage_of[i] = s;
      acc += node_cost(N[i].op);
      // close the stage once it holds its share of the cost (and a real node)
      if (s < S - 1 && !no_cut[i] && acc * S >= total_cost * (s + 1) && node_cost(N[i].op) > 0) ++s;
    }
    S = stage_of[n - 1] + 1;
  }
  // ---- channels: one per node whose row is read in a later stage (INPUT rows are re-read from `in`) ----
  std::vector<int> chan_of(n, -1);
  int n_chan = 0;
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < MLB_MAX_INS; ++k)
    {
      const int src = N[i].in[k];
      if (src < 0 || stage_of[src] == stage_of[i]) continue;
      if (N[src].op == MLB_OP_PARAM || N[src].op == MLB_OP_INPUT) continue;
      if (chan_of[src] < 0) chan_of[src] = n_chan++;
    }
  // ---- per-stage programs ----
  g->gnodes.clear();
  g->gstages.assign(S, GStage{});
  int max_slots = 1;
  for (int s = 0; s < S; ++s)
  {
    GStage& st = g->gstages[s];
    st.node_begin = (int)g->gnodes.size();
    std::vector<int> members;
    for (int i = 0; i < n; ++i)
      if (stage_of[i] == s) members.push_back(i);
    // rows produced outside this stage and read inside it
    std::vector<int> ext;
    for (int i : members)
      for (int k = 0; k < MLB_MAX_INS; ++k)
      {
        const int src = N[i].in[k];
        if (src < 0 || stage_of[src] == s || N[src].op == MLB_OP_PARAM) continue;
        if (std::find(ext.begin(), ext.end(), src) == ext.end()) ext.push_back(src);
      }
    // producer stages this stage must wait for (block t), and feedback partners
    auto add_wait = [&](int ws)
    {
      if (ws == s) return true;
      for (int q = 0; q < st.n_wait; ++q)
        if (st.wait_stage[q] == ws) return true;
      if (st.n_wait == GStage::kMaxWait) return false;
      st.wait_stage[st.n_wait++] = ws;
      return true;
    };
    auto add_fbwait = [&](int ws)
    {
      if (ws == s) return true;
      for (int q = 0; q < st.n_fbwait; ++q)
        if (st.fbwait_stage[q] == ws) return true;
      if (st.n_fbwait == GStage::kMaxWait) return false;
      st.fbwait_stage[st.n_fbwait++] = ws;
      return true;
    };
    bool ok = true;
    for (int src : ext)
      if (N[src].op != MLB_OP_INPUT) ok = ok && add_wait(stage_of[src]);
    for (int i : members)
    {
      // a FEEDBACK_READ at block t needs its writer's block t-1; the writer must not overwrite the
      // row before the reader of the same block has taken it
      if (N[i].op == MLB_OP_FEEDBACK_READ)
        for (int q = 0; q < n; ++q)
          if (N[q].op == MLB_OP_FEEDBACK_WRITE && N[q].iarg == i) ok = ok && add_fbwait(stage_of[q]);
      if (N[i].op == MLB_OP_FEEDBACK_WRITE) ok = ok && add_wait(stage_of[N[i].iarg]);
    }
    if (!ok) return fail(MLB_ERR_UNSUPPORTED, "stage %d of the interpreter program depends on too many stages", s);

    // liveness inside the stage: local index -> last local reader
    const int n_ext = (int)ext.size(), n_loc = n_ext + (int)members.size();
    auto local_of = [&](int node)
    {
      for (int e = 0; e < n_ext; ++e)
        if (ext[e] == node) return e;
      for (size_t m = 0; m < members.size(); ++m)
        if (members[m] == node) return n_ext + (int)m;
      return -1;
    };
    std::vector<int> last_use(n_loc, -1), slot(n_loc, -1), slot2(n_loc, -1);
    for (size_t m = 0; m < members.size(); ++m)
      for (int k = 0; k < MLB_MAX_INS; ++k)
      {
        const int src = N[members[m]].in[k];
        if (src < 0 || N[src].op == MLB_OP_PARAM) continue;
        last_use[local_of(src)] = n_ext + (int)m;
      }
    std::vector<int> free_slots;
    int n_slots = 0;
    auto alloc = [&]()
    {
      if (!free_slots.empty())
      {
        int q = free_slots.back();
        free_slots.pop_back();
        return q;
      }
      return n_slots++;
    };
    for (int li = 0; li < n_loc; ++li)
    {
      GNode gn{};
      const bool is_ext = li < n_ext;
      const int node = is_ext ? ext[li] : members[li - n_ext];
      gn.src_node = is_ext ? -1 : node;
      gn.out_slot = gn.out_slot2 = -1;
      gn.out_plane = -1;
      gn.chan_out = -1;
      for (int k = 0; k < MLB_MAX_INS; ++k) gn.in_kind[k] = OPERAND_NONE, gn.in_ref[k] = 0;
      if (is_ext)
      {
        // import: re-read an external input plane, or load the producer's channel plane
        if (N[node].op == MLB_OP_INPUT)
          gn.op = MLB_OP_INPUT, gn.iarg = N[node].iarg;
        else
          gn.op = MLB_OP_IMPORT_ROW, gn.iarg = chan_of[node];
        slot[li] = alloc();
      }
      else
      {
        gn.op = N[node].op;
        gn.st_off = g->st_off[node];
        gn.co_off = g->co_off[node];
        gn.iarg = N[node].iarg;
        gn.out_plane = out_plane[node];
        gn.chan_out = chan_of[node];
        for (int k = 0; k < MLB_MAX_INS; ++k)
        {
          const int src = N[node].in[k];
          if (src < 0) continue;
          if (N[src].op == MLB_OP_PARAM)
            gn.in_kind[k] = OPERAND_PARAM, gn.in_ref[k] = g->co_off[src];
          else
            gn.in_kind[k] = OPERAND_SLOT, gn.in_ref[k] = slot[local_of(src)];
        }
        // allocate the output slot BEFORE freeing inputs: nodes never run in place
        if (is_second(N[node].op))
          slot[li] = slot2[local_of(N[node].in[0])];
        else if (N[node].op != MLB_OP_PARAM)
          slot[li] = alloc();
        if (is_dual(N[node].op)) slot2[li] = alloc();
      }
      gn.out_slot = slot[li];
      gn.out_slot2 = slot2[li];
      g->gnodes.push_back(gn);
      // release rows whose last reader is this node
      for (int j = 0; j < li; ++j)
      {
        const int nj = j < n_ext ? ext[j] : members[j - n_ext];
        if (slot[j] < 0 || is_second(N[nj].op)) continue;
        if (last_use[j] == li) free_slots.push_back(slot[j]);
        if (is_dual(N[nj].op) && j >= n_ext)
        {
          // second row is read only through FDN8_R nodes; free it when their last reader ran
          int lu = -1;
          bool has_r = false;
          for (int q = j + 1; q < n_loc; ++q)
          {
            const int nq = q < n_ext ? ext[q] : members[q - n_ext];
            if (q >= n_ext && is_second(N[nq].op) && N[nq].in[0] == nj)
            {
              has_r = true;
              lu = std::max(lu, std::max(last_use[q], q));
            }
          }
          if (has_r && lu == li) free_slots.push_back(slot2[j]);
        }
      }
      // a dual-output node whose second row nobody in this stage reads gives that row back at once
      if (!is_ext && is_dual(N[node].op))
      {
        bool has_r = false;
        for (int q = li + 1; q < n_loc; ++q)
        {
          const int nq = q < n_ext ? ext[q] : members[q - n_ext];
          if (q >= n_ext && is_second(N[nq].op) && N[nq].in[0] == node) has_r = true;
        }
        if (!has_r) free_slots.push_back(slot2[li]);
      }
      // a row nobody in this stage reads can be recycled right after it was written out
      if (slot[li] >= 0 && last_use[li] < 0 && !(li >= n_ext && is_second(N[node].op)))
        free_slots.push_back(slot[li]);
    }
    st.node_end = (int)g->gnodes.size();
    max_slots = std::max(max_slots, n_slots);
  }
  g->n_slots = std::max(1, max_slots);
  g->scratch_slot = g->n_slots;
  for (int i = 0; i < n; ++i)
    if (N[i].op == MLB_OP_PITCHBEND_DELAY || N[i].op == MLB_OP_ALLPASS_PB || N[i].op == MLB_OP_ALLPASS_INT ||
        N[i].op == MLB_OP_ALLPASS_FRAC)
    {
      g->n_slots += 3;  // delay input + two tap streams (cp.async destinations)
      break;
    }
  g->n_stages = S;
  g->n_chan = n_chan;
  const size_t smem = (size_t)g->n_slots * kSlotBytes;
  if (smem > g_smem_optin)
    return fail(MLB_ERR_UNSUPPORTED, "graph needs %d live rows (%zu B shared memory > %zu)",
                g->n_slots, smem, g_smem_optin);
  g->kernel_name = std::string(g->exact ? "generic" : "generic(fast)") + "[" + std::to_string(S) + " stages, " +
                   std::to_string(g->n_slots) + " rows]";
  return MLB_OK;
}

static int size_functor_memory(mlb_graph* g);

// The interpreter's host-side plan for a graph, WITHOUT a device: which pipeline stage every node goes to and how many
// shared-memory row slots the program needs (assuming a B200: 148 SMs, 227 KB of shared memory per CTA, unless a device
// has been initialised).  The same build_generic() that mlb_graph_create runs; lets the planner's rules -- a feedback
// loop, a paired second row, a functor and its MLB_AGAIN calls stay in one stage; the row pool fits -- be tested on a CPU.
extern "C" int mlb_graph_plan(const mlb_node* nodes, int n_nodes, const int32_t* outs, int n_out, int n_voices,
                              unsigned flags, int32_t* stage_of, int32_t* n_stages, int32_t* n_row_slots)
{
  if (n_voices <= 0) return fail(MLB_ERR_INVALID, "n_voices must be positive");
  if (n_out < 0 || (n_out > 0 && !outs)) return fail(MLB_ERR_INVALID, "bad outs");
  mlb_layout lay;
  std::vector<int32_t> so(std::max(1, n_nodes)), co(std::max(1, n_nodes));
  int rc = mlb_graph_layout(nodes, n_nodes, &lay, so.data(), co.data());
  if (rc != MLB_OK) return rc;
  for (int c = 0; c < n_out; ++c)
    if (outs[c] < 0 || outs[c] >= n_nodes) return fail(MLB_ERR_INVALID, "outs[%d] out of range", c);
  const int sm_saved = g_sm_count;
  const size_t smem_saved = g_smem_optin;
  if (g_sm_count == 0) g_sm_count = 148;
  if (g_smem_optin == 0) g_smem_optin = 232448;
  mlb_graph g;  // host fields only: nothing here touches the CUDA runtime
  g.nodes.assign(nodes, nodes + n_nodes);
  g.outs.assign(outs, outs + n_out);
  g.st_off = so, g.co_off = co, g.layout = lay;
  g.V = n_voices, g.flags = flags, g.exact = !(flags & MLB_GRAPH_FAST);
  rc = build_generic(&g);
  g_sm_count = sm_saved, g_smem_optin = smem_saved;
  if (rc != MLB_OK) return rc;
  if (stage_of)
  {
    for (int i = 0; i < n_nodes; ++i) stage_of[i] = -1;  // PARAM nodes and the like belong to no stage
    for (int s = 0; s < (int)g.gstages.size(); ++s)
      for (int k = g.gstages[s].node_begin; k < g.gstages[s].node_end; ++k)
        if (g.gnodes[k].src_node >= 0) stage_of[g.gnodes[k].src_node] = s;
  }
  if (n_stages) *n_stages = g.n_stages;
  if (n_row_slots) *n_row_slots = g.n_slots;
  return MLB_OK;
}

extern "C" int mlb_graph_create(const mlb_node* nodes, int n_nodes, const int32_t* outs, int n_out,
                                int n_voices, unsigned flags, mlb_graph** out_graph)
{
  if (!out_graph) return fail(MLB_ERR_INVALID, "out_graph is null");
  *out_graph = nullptr;
  if (n_voices <= 0) return fail(MLB_ERR_INVALID, "n_voices must be positive");
  if (n_out < 0 || (n_out > 0 && !outs)) return fail(MLB_ERR_INVALID, "bad outs");
  mlb_layout lay;
  std::vector<int32_t> so(std::max(1, n_nodes)), co(std::max(1, n_nodes));
  int rc = mlb_graph_layout(nodes, n_nodes, &lay, so.data(), co.data());
  if (rc != MLB_OK) return rc;
  for (int c = 0; c < n_out; ++c)
    if (outs[c] < 0 || outs[c] >= n_nodes) return fail(MLB_ERR_INVALID, "outs[%d] out of range", c);
  rc = ensure_init();
  if (rc != MLB_OK) return rc;

  mlb_graph* g = new mlb_graph;
  ++g_live_handles;
  g->nodes.assign(nodes, nodes + n_nodes);
  g->outs.assign(outs, outs + n_out);
  g->st_off = so;
  g->co_off = co;
  g->layout = lay;
  g->V = n_voices;
  g->flags = flags;
  g->exact = !(flags & MLB_GRAPH_FAST);
  for (int i = 0; i < n_nodes; ++i)
  {
    if (nodes[i].op == MLB_OP_FDN8) g->fdn_node = i;
    int rows, rings;
    op_mem(nodes[i].op, &rows, &rings);
    if (rows || rings) g->has_dmem = true;
  }
  g->ring_stride.assign(n_nodes, 0u);

  auto cleanup = [&](int code)
  {
    mlb_graph_destroy(g);
    return code;
  };
  const size_t V = (size_t)n_voices;
  if (cudaMalloc(&g->d_state, std::max<size_t>(1, lay.n_state_words) * V * 4) != cudaSuccess ||
      cudaMalloc(&g->d_coef, std::max<size_t>(1, lay.n_coef_words) * V * 4) != cudaSuccess)
    return cleanup(fail(MLB_ERR_ALLOC, "cudaMalloc of voice state failed: %s",
                        cudaGetErrorString(cudaGetLastError())));
  {
    const size_t n_groups = (V + 31) / 32;
    // [0] unit counter, [1 + g] finished chunks of group g, [1 + G] finished units, [2 + G] progress base
    if (cudaMalloc(&g->d_sched, (n_groups + 3) * 4) != cudaSuccess)
      return cleanup(fail(MLB_ERR_ALLOC, "cudaMalloc of scheduler words failed"));
    cudaMemset(g->d_sched, 0, (n_groups + 3) * 4);
  }
  cudaMemset(g->d_state, 0, std::max<size_t>(1, lay.n_state_words) * V * 4);
  // state of freshly constructed functors: zeros, except the "idle" markers
  for (int i = 0; i < n_nodes; ++i)
  {
    int word = -1;
    uint32_t val = 0;
    if (nodes[i].op == MLB_OP_ADSR) word = 7, val = 4u;                   // segment{off}, F:694
    if (nodes[i].op == MLB_OP_GLIDE) word = 2, val = 0xFFFFFFFFu;         // mVectorsRemaining{-1}, G:440
    if (nodes[i].op == MLB_OP_SAMPLE_GLIDE) word = 3, val = 0xFFFFFFFFu;  // mSamplesRemaining{-1}, G:524
    if (nodes[i].op == MLB_OP_TEMPO_LOCK) word = 0, val = 0xBF800000u;    // _omega{-1.f}, F:1481
    if (word < 0) continue;
    std::vector<uint32_t> fill(V, val);
    cudaMemcpy(g->d_state + (size_t)(so[i] + word) * V, fill.data(), V * 4, cudaMemcpyHostToDevice);
  }
  cudaMemset(g->d_coef, 0, std::max<size_t>(1, lay.n_coef_words) * V * 4);
  g->h_coef.assign((size_t)lay.n_coef_words * V, 0.f);
  cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&g->s_h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&g->s_d2h, cudaStreamNonBlocking);
  for (int i = 0; i < mlb_graph::kMaxHostSlices; ++i)
  {
    cudaEventCreateWithFlags(&g->ev_up[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&g->ev_k[i], cudaEventDisableTiming);
  }
  cudaEventCreate(&g->ev0);
  cudaEventCreate(&g->ev1);

  bool matched = false;
  if (!(flags & MLB_GRAPH_FORCE_GENERIC))
  {
    matched = match_fused_chain(g);
    if (matched) g->kind = KIND_FUSED;
    if (!matched && !has_again_nodes(g) && match_fm3_fdn8(g->nodes, g->outs, g->st_off, g->co_off, &g->fargs))
    {
      matched = true;
      g->kind = KIND_FDN;
      g->kernel_name = g->exact ? "fused:fm3_fdn8" : "fused:fm3_fdn8(fast)";
    }
  }
  if (!matched)
  {
    g->kind = KIND_GENERIC;
    rc = build_generic(g);
    if (rc != MLB_OK) return cleanup(rc);
    const size_t sync_words = 1 + (size_t)g->n_stages * ((V + 31) / 32);
    if (cudaMalloc(&g->d_gnodes, sizeof(GNode) * g->gnodes.size()) != cudaSuccess ||
        cudaMalloc(&g->d_gstages, sizeof(GStage) * g->gstages.size()) != cudaSuccess ||
        cudaMalloc(&g->d_gsync, sync_words * 4) != cudaSuccess)
      return cleanup(fail(MLB_ERR_ALLOC, "cudaMalloc of graph program failed"));
    cudaMemcpy(g->d_gnodes, g->gnodes.data(), sizeof(GNode) * g->gnodes.size(),
               cudaMemcpyHostToDevice);
    cudaMemcpy(g->d_gstages, g->gstages.data(), sizeof(GStage) * g->gstages.size(), cudaMemcpyHostToDevice);
    rc = size_functor_memory(g);
    if (rc != MLB_OK) return cleanup(rc);
  }
  *out_graph = g;
  return MLB_OK;
}

extern "C" int mlb_graph_destroy(mlb_graph* g)
{
  if (!g) return MLB_OK;
  cudaFree(g->d_state);
  cudaFree(g->d_coef);
  cudaFree(g->d_gnodes);
  cudaFree(g->d_gstages);
  cudaFree(g->d_gsync);
  cudaFree(g->d_chan);
  cudaFree(g->d_ring);
  cudaFree(g->d_carry);
  cudaFree(g->d_dmem);
  cudaFree(g->d_partial);
  cudaFree(g->d_sched);
  cudaFree(g->d_in);
  cudaFree(g->d_out);
  cudaFree(g->d_mix);
  if (g->ev0) cudaEventDestroy(g->ev0);
  if (g->ev1) cudaEventDestroy(g->ev1);
  for (int i = 0; i < mlb_graph::kMaxHostSlices; ++i)
  {
    if (g->ev_up[i]) cudaEventDestroy(g->ev_up[i]);
    if (g->ev_k[i]) cudaEventDestroy(g->ev_k[i]);
  }
  if (g->s_mix) cudaStreamDestroy(g->s_mix);
  if (g->ev_main) cudaEventDestroy(g->ev_main);
  for (cudaEvent_t e : g->ev_mixdone)
    if (e) cudaEventDestroy(e);
  if (g->s_h2d) cudaStreamDestroy(g->s_h2d);
  if (g->s_d2h) cudaStreamDestroy(g->s_d2h);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  --g_live_handles;
  return MLB_OK;
}

extern "C" int mlb_graph_set_input_planes(mlb_graph* g, int n_planes)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  int need = 0;
  for (const mlb_node& n : g->nodes)
    if (n.op == MLB_OP_INPUT) need = std::max(need, n.iarg + 1);
  if (n_planes < need || n_planes > 64)
    return fail(MLB_ERR_INVALID, "n_planes must be in [%d, 64] for this graph", need);
  g->layout.n_inputs = n_planes;
  return MLB_OK;
}
extern "C" int mlb_graph_reserve_sms(mlb_graph* g, int n_sms)
{
  if (!g || n_sms < 0) return fail(MLB_ERR_INVALID, "bad argument");
  g->reserved_sms = n_sms;
  return MLB_OK;
}
extern "C" int mlb_graph_layout_of(const mlb_graph* g, mlb_layout* layout)
{
  if (!g || !layout) return fail(MLB_ERR_INVALID, "null");
  *layout = g->layout;
  return MLB_OK;
}
extern "C" const char* mlb_graph_kernel_name(const mlb_graph* g) { return g ? g->kernel_name.c_str() : ""; }

// (re)allocate FDN delay memory when the coefficient upload changes the longest delay
static int size_delay_memory(mlb_graph* g)
{
  if (g->fdn_node < 0) return MLB_OK;
  const size_t V = (size_t)g->V;
  const int base = g->co_off[g->fdn_node] + 24;
  int max_len = 1;
  for (int l = 0; l < 8; ++l)
    for (size_t v = 0; v < V; ++v)
    {
      const float f = g->h_coef[(size_t)(base + l) * V + v];
      if (!(f >= 1.0f) || f > 16777216.0f || f != std::floor(f))
        return fail(MLB_ERR_INVALID, "FDN8 delay length coef must be an integer >= 1 (voice %zu line %d: %g)",
                    v, l, (double)f);
      max_len = std::max(max_len, (int)f);
    }
  // IntegerDelay::setMaxDelayInSamples, F:822-830: size = 1 << bitsToContain(dMax + 64)
  int ring = 1;
  while (ring < max_len + MLB_BLOCK) ring <<= 1;
  if (ring != g->ring_len)
  {
    cudaFree(g->d_ring);
    cudaFree(g->d_carry);
    g->d_ring = g->d_carry = nullptr;
    const size_t ring_bytes = V * 8 * (size_t)ring * 4, carry_bytes = V * 8 * MLB_BLOCK * 4;
    if (cudaMalloc(&g->d_ring, ring_bytes) != cudaSuccess || cudaMalloc(&g->d_carry, carry_bytes) != cudaSuccess)
      return fail(MLB_ERR_ALLOC, "cudaMalloc of %zu B delay memory failed", ring_bytes + carry_bytes);
    g->ring_len = ring;
    return mlb_graph_clear_delays(g);
  }
  return MLB_OK;
}

// Lay out (and, when a ring grows or shrinks, reallocate and clear) the delay memory of the
// section-8(f) functors.  Ring of voice v of a node: 1 << bitsToContain(floor(maxDelay_v') + 64)
// samples (IntegerDelay::setMaxDelayInSamples, F:822-830; maxDelay' = maxDelay - 64 inside an
// Allpass<>, F:1125-1128); every voice of a node gets the stride of the node's longest ring and
// uses its own mask inside it.
static int size_functor_memory(mlb_graph* g)
{
  if (!g->has_dmem) return MLB_OK;
  const size_t V = (size_t)g->V;
  const int n = (int)g->nodes.size();
  std::vector<unsigned> stride(n, 0u);
  std::vector<unsigned long long> row_off(n, 0ull), ring_off(n, 0ull);
  unsigned long long total = 0;
  for (int i = 0; i < n; ++i)
  {
    int rows, rings, nco = 0;
    const int op = g->nodes[i].op;
    op_mem(op, &rows, &rings);
    mlb_op_info(op, nullptr, nullptr, &nco);
    if (again_target(g->nodes.data(), i) >= 0)  // MLB_AGAIN: the member row of the functor's first call (no rings, by rule)
    {
      row_off[i] = row_off[again_target(g->nodes.data(), i)];
      continue;
    }
    if (rows)
    {
      row_off[i] = total;
      total += (unsigned long long)V * MLB_BLOCK;
    }
    if (rings)
    {
      const bool in_allpass = (op == MLB_OP_ALLPASS_INT || op == MLB_OP_ALLPASS_FRAC || op == MLB_OP_ALLPASS_PB);
      const float* md = g->h_coef.data() + (size_t)(g->co_off[i] + nco - 1) * V;  // last coef word = maxDelay
      int dmax = 0;
      for (size_t v = 0; v < V; ++v)
      {
        const float m = md[v] - (in_allpass ? (float)MLB_BLOCK : 0.f);
        if (!(m >= 0.f) || m > 16777216.0f)
        {
          if (md[v] == 0.f) continue;  // coefficients not set yet: smallest ring
          return fail(MLB_ERR_INVALID, "node %d (%s): maxDelay coef of voice %zu is %g (need %s <= maxDelay <= 2^24)", i,
                      mlb_op_name(op), v, (double)md[v], in_allpass ? "64" : "0");
        }
        dmax = std::max(dmax, (int)std::floor(m));
      }
      unsigned ring = 1;
      while (ring < (unsigned)(dmax + MLB_BLOCK)) ring <<= 1;
      stride[i] = ring;
      ring_off[i] = total;
      total += (unsigned long long)V * ring;
    }
  }
  for (int i = 0; i < n; ++i)  // FEEDBACK_WRITE stores into its FEEDBACK_READ's row
    if (g->nodes[i].op == MLB_OP_FEEDBACK_WRITE) row_off[i] = row_off[g->nodes[i].iarg];
  const bool relayout = (total != g->dmem_floats) || (stride != g->ring_stride);
  if (!relayout) return MLB_OK;
  cudaFree(g->d_dmem);
  g->d_dmem = nullptr;
  g->dmem_floats = 0;
  if (cudaMalloc(&g->d_dmem, std::max<unsigned long long>(total, 1) * 4) != cudaSuccess)
    return fail(MLB_ERR_ALLOC, "cudaMalloc of %llu B delay memory failed", total * 4ull);
  g->dmem_floats = (size_t)total;
  g->ring_stride = stride;
  for (GNode& gn : g->gnodes)
  {
    if (gn.src_node < 0) continue;
    gn.ring_stride = stride[gn.src_node];
    gn.row_off = row_off[gn.src_node];
    gn.ring_off = ring_off[gn.src_node];
  }
  CU_CHECK(cudaMemcpy(g->d_gnodes, g->gnodes.data(), sizeof(GNode) * g->gnodes.size(), cudaMemcpyHostToDevice));
  return mlb_graph_clear_delays(g);
}

extern "C" int mlb_graph_set_coefs(mlb_graph* g, const float* coef_host)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  if (g->layout.n_coef_words == 0) return MLB_OK;
  if (!coef_host) return fail(MLB_ERR_INVALID, "null coefs");
  const size_t bytes = (size_t)g->layout.n_coef_words * g->V * 4;
  // the sizing passes read the host mirror: hand them the new values, put the old ones back when they are
  // rejected so that the mirror never disagrees with the device copy.  NOTE: a change of any ring size
  // re-lays-out the delay memory and clears ALL delay state of the graph (rings, member rows, write index).
  std::vector<float> old(g->h_coef);
  memcpy(g->h_coef.data(), coef_host, bytes);
  int rc = size_delay_memory(g);
  if (rc == MLB_OK) rc = size_functor_memory(g);
  if (rc != MLB_OK)
  {
    g->h_coef.swap(old);
    return rc;
  }
  CU_CHECK(cudaMemcpy(g->d_coef, coef_host, bytes, cudaMemcpyHostToDevice));
  return MLB_OK;
}
extern "C" int mlb_graph_set_state(mlb_graph* g, const uint32_t* state_host)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  if (g->layout.n_state_words == 0) return MLB_OK;
  if (!state_host) return fail(MLB_ERR_INVALID, "null state");
  CU_CHECK(cudaMemcpy(g->d_state, state_host, (size_t)g->layout.n_state_words * g->V * 4,
                      cudaMemcpyHostToDevice));
  return MLB_OK;
}
extern "C" int mlb_graph_get_state(mlb_graph* g, uint32_t* state_host)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  if (g->layout.n_state_words == 0) return MLB_OK;
  if (!state_host) return fail(MLB_ERR_INVALID, "null state");
  CU_CHECK(cudaDeviceSynchronize());
  CU_CHECK(cudaMemcpy(state_host, g->d_state, (size_t)g->layout.n_state_words * g->V * 4,
                      cudaMemcpyDeviceToHost));
  return MLB_OK;
}
extern "C" int mlb_graph_clear_delays(mlb_graph* g)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  g->blocks_done = 0;
  if (g->d_ring)
  {
    CU_CHECK(cudaMemset(g->d_ring, 0, (size_t)g->V * 8 * g->ring_len * 4));
    CU_CHECK(cudaMemset(g->d_carry, 0, (size_t)g->V * 8 * MLB_BLOCK * 4));
  }
  if (g->d_dmem) CU_CHECK(cudaMemset(g->d_dmem, 0, g->dmem_floats * 4));
  return MLB_OK;
}
extern "C" size_t mlb_graph_delay_bytes(const mlb_graph* g)
{
  if (!g) return 0;
  size_t bytes = g->dmem_floats * 4;
  if (g->d_ring) bytes += (size_t)g->V * 8 * ((size_t)g->ring_len + MLB_BLOCK) * 4;
  return bytes;
}

// ------------------------------------------------------------------------------------------
// multi-GPU mix bus over peer memory (one process per GPU; buffers shared through CUDA IPC)

struct mlb_mixbus
{
  int rank = 0, world = 1;
  size_t n_floats = 0;   // floats per slot
  int n_planes_cap = 0;  // flag words per (parity, rank)
  void* base = nullptr;  // one allocation: xchg [2][world][n_floats] f32, then flags [2][world][n_planes_cap] u32
  void* peer_base[kMaxBusRanks] = {};
  bool connected = false;
  unsigned seq = 0;
  int async = 0;                  // completion on a side stream (mlb_mixbus_set_async)
  cudaStream_t side = nullptr;
  static constexpr int kStage = 4;  // calls whose exchange may still be pending on the side stream
  cudaEvent_t ev_posted = nullptr, ev_done[kStage] = {};
  bool done_pending = false;
  size_t xchg_bytes() const { return (size_t)2 * world * n_floats * 4; }
  size_t flag_bytes() const { return (size_t)2 * world * n_planes_cap * 4; }
  size_t stage_bytes() const { return (size_t)kStage * n_floats * 4; }
  size_t total_bytes() const { return xchg_bytes() + 2 * flag_bytes() + stage_bytes(); }  // slots, flags, acks, staging
};

extern "C" int mlb_mixbus_create(int rank, int world, size_t max_floats, mlb_mixbus** out)
{
  if (!out) return fail(MLB_ERR_INVALID, "out is null");
  *out = nullptr;
  if (world < 1 || world > kMaxBusRanks || rank < 0 || rank >= world || max_floats == 0 || (max_floats % MLB_BLOCK))
    return fail(MLB_ERR_INVALID, "mlb_mixbus_create: need 0 <= rank < world <= %d and max_floats a multiple of 64", kMaxBusRanks);
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  mlb_mixbus* b = new mlb_mixbus;
  b->rank = rank, b->world = world, b->n_floats = max_floats, b->n_planes_cap = (int)(max_floats / MLB_BLOCK);
  if (cudaMalloc(&b->base, b->total_bytes()) != cudaSuccess)
  {
    delete b;
    return fail(MLB_ERR_ALLOC, "cudaMalloc of %zu B exchange buffer failed", b->total_bytes());
  }
  cudaMemset(b->base, 0, b->total_bytes());
  cudaDeviceSynchronize();
  b->peer_base[rank] = b->base;
  b->connected = (world == 1);
  cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking);
  // The exchange kernel runs BESIDE the next call's chain kernel, whose CTAs need the SM configured for the
  // maximum shared-memory carve-out: a resident CTA that prefers another L1 / shared split keeps the chain CTA
  // off its SM until it has finished (measured: 16 us per step).  Ask for the same split.
  cudaFuncSetAttribute((const void*)mixbus_exchange_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                       (int)cudaSharedmemCarveoutMaxShared);
  cudaEventCreateWithFlags(&b->ev_posted, cudaEventDisableTiming);
  for (cudaEvent_t& e : b->ev_done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  ++g_live_handles;
  *out = b;
  return MLB_OK;
}
extern "C" int mlb_mixbus_handle(mlb_mixbus* b, void* out64)
{
  if (!b || !out64) return fail(MLB_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  CU_CHECK(cudaIpcGetMemHandle(&h, b->base));
  memcpy(out64, &h, 64);
  return MLB_OK;
}
extern "C" int mlb_mixbus_connect(mlb_mixbus* b, const void* handles)
{
  if (!b || !handles) return fail(MLB_ERR_INVALID, "null argument");
  for (int r = 0; r < b->world; ++r)
  {
    if (r == b->rank || b->peer_base[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
    CU_CHECK(cudaIpcOpenMemHandle(&b->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  b->connected = true;
  return MLB_OK;
}
extern "C" int mlb_mixbus_destroy(mlb_mixbus* b)
{
  if (!b) return MLB_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < b->world; ++r)
    if (r != b->rank && b->peer_base[r]) cudaIpcCloseMemHandle(b->peer_base[r]);
  cudaFree(b->base);
  if (b->side) cudaStreamDestroy(b->side);
  if (b->ev_posted) cudaEventDestroy(b->ev_posted);
  for (cudaEvent_t e : b->ev_done)
    if (e) cudaEventDestroy(e);
  delete b;
  --g_live_handles;
  return MLB_OK;
}
extern "C" int mlb_mixbus_set_async(mlb_mixbus* b, int on)
{
  if (!b) return fail(MLB_ERR_INVALID, "null bus");
  if (b->seq != 0 && (on != 0) != (b->async != 0))
    return fail(MLB_ERR_INVALID, "mlb_mixbus_set_async: choose the mode before the first process call");
  b->async = on ? 1 : 0;
  return MLB_OK;
}
extern "C" int mlb_graph_mix_wait(mlb_graph* g, void* stream)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  mlb_mixbus* b = g->bus;
  if (g->mix_async && g->mix_pending)
    CU_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, g->ev_mixdone[g->mix_seq & 3u], 0));
  if (b && b->async && b->done_pending)
    CU_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, b->ev_done[b->seq % mlb_mixbus::kStage], 0));
  return MLB_OK;
}
extern "C" int mlb_graph_attach_mixbus(mlb_graph* g, mlb_mixbus* b)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  if (b && !b->connected) return fail(MLB_ERR_INVALID, "mix bus is not connected to its peers yet");
  g->bus = b;
  return MLB_OK;
}
// kernel arguments of the next call's exchange (advances the call counter)
static int bus_args(mlb_graph* g, int n_planes, MixBusArgs* a, cudaStream_t stream)
{
  memset(a, 0, sizeof(*a));
  a->world = 1;
  mlb_mixbus* b = g->bus;
  if (!b || b->world <= 1) return MLB_OK;
  if ((size_t)n_planes * MLB_BLOCK > b->n_floats)
    return fail(MLB_ERR_INVALID, "mix bus holds %zu floats per rank, this call needs %zu", b->n_floats,
                (size_t)n_planes * MLB_BLOCK);
  a->rank = b->rank, a->world = b->world;
  a->seq = ++b->seq;
  a->n_floats = (int)b->n_floats, a->n_planes_cap = b->n_planes_cap;
  for (int r = 0; r < b->world; ++r)
  {
    a->xchg[r] = (float*)b->peer_base[r];
    a->flags[r] = (unsigned*)((char*)b->peer_base[r] + b->xchg_bytes());
    a->acks[r] = (unsigned*)((char*)b->peer_base[r] + b->xchg_bytes() + b->flag_bytes());
  }
  a->async = b->async;
  if (b->async)
  {
    // staging block of this call; its previous user (four calls ago) must have been exchanged
    a->stage = (float*)((char*)b->base + b->xchg_bytes() + 2 * b->flag_bytes()) + (size_t)(a->seq % mlb_mixbus::kStage) * b->n_floats;
    if (a->seq > (unsigned)mlb_mixbus::kStage)
      CU_CHECK(cudaStreamWaitEvent(stream, b->ev_done[a->seq % mlb_mixbus::kStage], 0));
  }
  return MLB_OK;
}
// async mode: the exchange on the bus's side stream, ordered after the local sums on `stream`
static int bus_complete(mlb_graph* g, const MixBusArgs& ba, int n_planes, float* mix_dev, cudaStream_t stream)
{
  mlb_mixbus* b = g->bus;
  if (!b || b->world <= 1 || !b->async) return MLB_OK;
  static const int dbg = env_int("MLB_BUS_DEBUG", 0);  // timing experiments only: 1 = events but no kernel, 2 = nothing
  if (dbg == 2) return MLB_OK;
  CU_CHECK(cudaEventRecord(b->ev_posted, stream));
  CU_CHECK(cudaStreamWaitEvent(b->side, b->ev_posted, 0));
  if (dbg != 1)
  {
    mixbus_exchange_kernel<<<(n_planes + kExchangePlanesPerCta - 1) / kExchangePlanesPerCta,
                             64 * kExchangePlanesPerCta, 0, b->side>>>(mix_dev, ba, n_planes);
    ++g_launches;
    CU_CHECK(cudaGetLastError());
  }
  CU_CHECK(cudaEventRecord(b->ev_done[ba.seq % mlb_mixbus::kStage], b->side));
  b->done_pending = true;
  return MLB_OK;
}

static int env_int(const char* name, int dflt)
{
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

static int ensure_buf(float** p, size_t* cap, size_t bytes)
{
  if (bytes <= *cap) return MLB_OK;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  if (cudaMalloc(p, bytes) != cudaSuccess) return fail(MLB_ERR_ALLOC, "cudaMalloc of %zu B staging failed", bytes);
  *cap = bytes;
  return MLB_OK;
}

static int ensure_partial(mlb_graph* g, int T, int n_groups)
{
  // per-group partials followed by the per-chunk scratch of mix_reduce_kernel
  const size_t n_chunks = ((size_t)n_groups + kMixChunkGroups - 1) / kMixChunkGroups;
  const size_t one = (size_t)T * std::max<size_t>(1, g->outs.size()) * ((size_t)n_groups + n_chunks) * MLB_BLOCK * 4;
  const size_t need = one * (g->mix_async ? 2 : 1);  // async: two buffers, alternating by call parity
  if (need > g->partial_cap)
  {
    if (g->mix_pending) cudaStreamSynchronize(g->s_mix);  // nothing may still read the old buffers
    cudaFree(g->d_partial);
    g->d_partial = nullptr;
    g->partial_cap = 0;
    if (cudaMalloc(&g->d_partial, need) != cudaSuccess)
      return fail(MLB_ERR_ALLOC, "cudaMalloc of %zu B mix partials failed", need);
    g->partial_cap = need;
  }
  g->partial_cur = g->d_partial;
  return MLB_OK;
}

// Launch the fused chain kernel for the voice slice [va, vb) (va a multiple of 32) of the bank.
// in_dev / out_dev are the FULL planes [T][n][V][64]; the tensor maps, SoA pointers, mix
// partials and progress words are offset to the slice.
static int launch_chain_slice(mlb_graph* g, const float* in_dev, float* out_dev, bool want_mix, int T,
                              cudaStream_t stream, int va, int vb, int force_chunks = 0)
{
  const FusedEntry& e = *g->fused;
  const int V = g->V, Vs = vb - va;
  const int groups_total = (V + 31) / 32, n_groups = (Vs + 31) / 32;
  const int n_in = g->layout.n_inputs;
  int rc;
  ChainArgs a = g->cargs;
  a.state = g->d_state + va;
  a.coef = g->d_coef + va;
  a.mix_partial = want_mix ? g->partial_cur + (size_t)(va / 32) * MLB_BLOCK : nullptr;
  a.V = Vs;
  a.T = T;
  a.v_stride = V;
  a.n_in_planes = std::max(1, n_in);
  a.n_out_planes = 1;
  a.out_plane = 0;
  a.n_groups = n_groups;
  a.groups_stride = groups_total;
  a.write_out = out_dev ? 1 : 0;
  // Small banks: a TEAM of two warps per voice group (generator | filter pipeline, chain_team_kernel):
  // with fewer than ~4 groups per SM a lone warp per group is latency-bound, and the chain is the
  // only axis left to split.  One 64-thread CTA per group, static assignment, no scheduler words.
  // Up to 4 teams per SM (5-stage rings, 41 KB each); beyond that one warp per group.
  // (teams at 5 and 6 per SM measured no better than one lone warp per group: 0.204 ms at 20 480 and 24 576 voices)
  const int teams_per_sm = (n_groups + g_sm_count - 1) / g_sm_count;
  if (e.team_fn && teams_per_sm <= 4 && env_int("MLB_CHAIN_TEAM", 1) != 0)
  {
    // 3 S named barriers + barrier 0 <= 16  ->  S <= 5
    const int S = std::min(std::max(env_int("MLB_TEAM_STAGES", 5), 4), 5);
    a.stages = S;
    a.chunk_blocks = T, a.n_chunks = 1;
    a.sched = g->d_sched, a.progress = g->d_sched + 1 + va / 32;
    a.done = g->d_sched + 1 + groups_total, a.base_word = g->d_sched + 2 + groups_total;
    g->launch_chunks = 1;
    const size_t smem = (size_t)S * kBlockBytes + (size_t)S * 8;
    CUtensorMap in_map, out_map;
    memset(&in_map, 0, sizeof(in_map));
    memset(&out_map, 0, sizeof(out_map));
    if (e.has_in)
    {
      rc = cached_block_map(&g->maps, &in_map, in_dev + (size_t)va * MLB_BLOCK, Vs, (long long)T * a.n_in_planes,
                            (long long)V * MLB_BLOCK);
      if (rc != MLB_OK) return rc;
    }
    if (out_dev)
    {
      rc = cached_block_map(&g->maps, &out_map, out_dev + (size_t)va * MLB_BLOCK, Vs, (long long)T,
                            (long long)V * MLB_BLOCK);
      if (rc != MLB_OK) return rc;
    }
    rc = ensure_func_smem((const void*)e.team_fn, smem);
    if (rc != MLB_OK) return rc;
    static unsigned long long* d_prof = nullptr;
    const bool prof = env_int("MLB_TEAM_PROF", 0) != 0;
    if (prof && !d_prof) cudaMalloc(&d_prof, 64);
    a.prof = prof ? d_prof : nullptr;
    a.n_sms = g_sm_count;
    const int team_threads = teams_per_sm <= 2 ? 128 : 96;  // see chain_team_kernel: role rotation only pays up to 2 teams per SM
    e.team_fn<<<n_groups, team_threads, smem, stream>>>(in_map, out_map, a);
    ++g_launches;
    CU_CHECK(cudaGetLastError());
    if (prof)
    {
      unsigned long long h[5];
      cudaStreamSynchronize(stream);
      cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
      const double n = (double)T * MLB_BLOCK;
      int occ = -1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)e.team_fn, 96, smem);
      fprintf(stderr, "team launch: %d groups, S = %d, %zu B smem, occupancy %d CTAs/SM\n", n_groups, S, smem, occ);
      fprintf(stderr, "team prof (cycles per sample, CTA 0): G wait %.1f compute %.1f | F wait %.1f compute %.1f\n",
              h[0] / n, h[1] / n, h[2] / n, h[3] / n);
    }
    return MLB_OK;
  }
  // Launch shape (DESIGN.md "occupancy").  The grid is persistent: one CTA of W warps per SM,
  // each warp owns a ring of S 8-KB blocks and pulls (group, chunk) work units from an atomic
  // queue.  W is a multiple of 4 so the four SM sub-partitions carry equal warp counts; small
  // banks simply get one warp per group.
  const size_t budget = (size_t)g_smem_optin - 64;
  const size_t stage_bytes = (size_t)std::max(1, e.n_planes) * kBlockBytes;  // a ring stage holds n_planes blocks
  const int w_cap = std::min(kChainMaxWarps, (int)(budget / (2 * (stage_bytes + 8))));  // warps whose 2-stage rings fit
  // Three regimes (n_groups = 32-voice groups of this launch):
  //   * up to 8 warps per SM cover every group (n_groups <= 8 * SMs): W = ceil(n_groups / SMs) warps per
  //     CTA, every warp owns ONE group for the whole launch (n_chunks = 1): no state hopping, no progress waits,
  //     and the spare shared memory deepens each warp's ring.  (Cutting such a launch into time chunks adds no
  //     parallelism -- a group's chunks are serial anyway -- it only idles the warps beyond n_groups.)
  //   * more groups: W = 12 (3 per SM sub-partition, 2-stage rings, 197 KB: measured best, profiles/), work
  //     units = (group, chunk of the launch's blocks) so that SM-to-SM speed differences and the tail stay
  //     balanced (static one-group-per-warp measured 0.415 ms against 0.361 ms at 2 048 groups, round 1).
  int W, S;
  const bool one_warp_per_group = n_groups <= g_sm_count * std::min(w_cap, 8);
  if (one_warp_per_group)
    W = std::max(1, (n_groups + g_sm_count - 1) / g_sm_count);
  else
    W = std::min(12, w_cap);
  W = env_int("MLB_CHAIN_WARPS", W);
  W = std::min(std::max(W, 1), kChainMaxWarps);
  S = 2;
  if (e.n_planes > 0)
  {
    S = (int)(budget / ((size_t)W * (stage_bytes + 8)));
    S = std::min(std::max(S, 2), 6);
    S = env_int("MLB_CHAIN_STAGES", S);
    S = std::min(std::max(S, 2), 12);
  }
  a.stages = S;
  // work units: (group, chunk of blocks).  Aim at >= 6 units per resident warp so the tail of
  // the dynamic schedule costs < 1/6 of a unit column; chunks never shorter than 8 blocks.
  {
    const int n_warps_resident = g_sm_count * W;
    int n_chunks = (int)((6LL * n_warps_resident + n_groups - 1) / n_groups);
    n_chunks = std::min(std::max(n_chunks, 1), std::max(1, T / 8));
    if (one_warp_per_group && n_groups <= n_warps_resident) n_chunks = 1;
    n_chunks = env_int("MLB_CHAIN_CHUNKS", n_chunks);
    if (force_chunks > 0) n_chunks = force_chunks;  // all slices of one call must agree
    n_chunks = std::min(std::max(n_chunks, 1), T);
    a.chunk_blocks = (T + n_chunks - 1) / n_chunks;
    a.n_chunks = (T + a.chunk_blocks - 1) / a.chunk_blocks;
    a.sched = g->d_sched;
    a.progress = g->d_sched + 1 + va / 32;
    a.done = g->d_sched + 1 + groups_total;
    a.base_word = g->d_sched + 2 + groups_total;
    g->launch_chunks = a.n_chunks;
  }
  const size_t smem = (size_t)W * S * stage_bytes + (size_t)W * S * 8;
  if (smem > g_smem_optin) return fail(MLB_ERR_INVALID, "chain launch needs %zu B shared memory", smem);
  const int ctas_per_sm = std::max<size_t>(1, (size_t)(227 * 1024) / (smem + 1024));
  // the grid is persistent and fills every SM's shared memory; a caller that overlaps a collective
  // (the mix-bus all-reduce of a multi-GPU run) keeps a few SMs free for its kernels
  const int sms = std::max(1, g_sm_count - std::max(0, g->reserved_sms));
  int n_ctas = std::min((n_groups + W - 1) / W, sms * ctas_per_sm);
  n_ctas = std::min(std::max(env_int("MLB_CHAIN_CTAS", n_ctas), 1), (n_groups + W - 1) / W);
  CUtensorMap in_map, out_map;
  memset(&in_map, 0, sizeof(in_map));
  memset(&out_map, 0, sizeof(out_map));
  if (e.n_planes > 0)
  {
    rc = cached_block_map(&g->maps, &in_map, in_dev + (size_t)va * MLB_BLOCK, Vs, (long long)T * a.n_in_planes,
                          (long long)V * MLB_BLOCK);
    if (rc != MLB_OK) return rc;
  }
  if (out_dev)
  {
    rc = cached_block_map(&g->maps, &out_map, out_dev + (size_t)va * MLB_BLOCK, Vs, (long long)T,
                          (long long)V * MLB_BLOCK);
    if (rc != MLB_OK) return rc;
  }
  rc = ensure_func_smem((const void*)e.fn, smem);
  if (rc != MLB_OK) return rc;
  e.fn<<<n_ctas, W * 32, smem, stream>>>(in_map, out_map, a);
  ++g_launches;
  CU_CHECK(cudaGetLastError());
  return MLB_OK;
}

extern "C" int mlb_graph_process_device(mlb_graph* g, const float* in_dev, float* out_dev,
                                        float* mix_dev, int n_blocks, void* stream_v)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  if (n_blocks <= 0) return fail(MLB_ERR_INVALID, "n_blocks must be positive");
  if (g->layout.n_inputs > 0 && !in_dev) return fail(MLB_ERR_INVALID, "graph has INPUT nodes but in is null");
  if (g->outs.empty() && (out_dev || mix_dev)) return fail(MLB_ERR_INVALID, "graph has no outputs");
  if (g->fdn_node >= 0 && !g->d_ring)
    return fail(MLB_ERR_INVALID, "FDN8 graph: call mlb_graph_set_coefs before processing");
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int V = g->V, T = n_blocks;
  const int n_groups = (V + 31) / 32;
  const int n_out = (int)g->outs.size(), n_in = g->layout.n_inputs;
  if (mix_dev)
  {
    rc = ensure_partial(g, T, n_groups);
    if (rc != MLB_OK) return rc;
    if (g->mix_async)
    {
      // this call's partial buffer; its previous user (two calls ago) must have been reduced
      ++g->mix_seq;
      g->partial_cur = g->d_partial + (size_t)(g->mix_seq & 1u) * (g->partial_cap / 8);
      if (g->mix_seq > 2u) CU_CHECK(cudaStreamWaitEvent(stream, g->ev_mixdone[(g->mix_seq - 2u) & 3u], 0));
    }
  }

  // the timing events are skipped while `stream` is being captured into a CUDA graph
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cap);
  const bool capturing = (cap != cudaStreamCaptureStatusNone);
  if (capturing && (g->fdn_node >= 0 || g->has_dmem || (g->bus && g->bus->world > 1) || (g->mix_async && mix_dev)))
    return fail(MLB_ERR_UNSUPPORTED, "CUDA-graph capture of a process call is supported for graphs without delay "
                                     "memory and without a mix bus (their call counters are kernel arguments)");
  if (!capturing) cudaEventRecord(g->ev0, stream);
  if (g->kind == KIND_FUSED)
  {
    rc = launch_chain_slice(g, in_dev, out_dev, mix_dev != nullptr, T, stream, 0, V);
    if (rc != MLB_OK) return rc;
  }
  else if (g->kind == KIND_FDN)
  {
    // outputs are needed for the mix bus even when the caller does not want them
    float* planes = out_dev;
    if (!planes)
    {
      rc = ensure_buf(&g->d_out, &g->out_cap, (size_t)T * 2 * V * MLB_BLOCK * 4);
      if (rc != MLB_OK) return rc;
      planes = g->d_out;
    }
    rc = launch_fm3_fdn8(g->fargs, g->exact, g->d_state, g->d_coef, g->d_ring, g->d_carry, g->ring_len,
                         g->blocks_done, in_dev, planes, V, T, std::max(1, n_in), stream);
    if (rc != MLB_OK) return fail(rc, "fdn8 launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ++g_launches;
    if (mix_dev)
    {
      mix_partial_from_planes_kernel<<<dim3(n_groups, T * 2), MLB_BLOCK, 0, stream>>>(planes, g->partial_cur, V,
                                                                                      n_groups);
      ++g_launches;
    }
  }
  else
  {
    GenericArgs a;
    memset(&a, 0, sizeof(a));
    a.nodes = g->d_gnodes;
    a.n_nodes = (int)g->gnodes.size();
    a.state = g->d_state;
    a.coef = g->d_coef;
    a.in = in_dev;
    a.out = out_dev;
    a.mix_partial = mix_dev ? g->partial_cur : nullptr;
    a.V = V, a.T = T, a.n_in = std::max(1, n_in), a.n_out = std::max(1, n_out);
    a.n_groups = n_groups, a.n_slots = g->n_slots;
    a.fdn_ring = g->d_ring, a.fdn_carry = g->d_carry, a.fdn_ring_len = g->ring_len;
    a.blocks_done = g->blocks_done;
    a.dmem = g->d_dmem;
    a.scratch_slot = g->scratch_slot;
    a.stages = g->d_gstages;
    a.n_stages = g->n_stages;
    a.n_chan = g->n_chan;
    a.sync = g->d_gsync;
    if (g->n_chan > 0)
    {
      rc = ensure_buf(&g->d_chan, &g->chan_cap, (size_t)T * g->n_chan * V * MLB_BLOCK * 4);
      if (rc != MLB_OK) return rc;
    }
    a.chan = g->d_chan;
    const int grid = n_groups * g->n_stages;
    CU_CHECK(cudaMemsetAsync(g->d_gsync, 0, (1 + (size_t)grid) * 4, stream));
    const size_t smem = (size_t)g->n_slots * kSlotBytes;
    if (g->exact)
    {
      CU_CHECK(cudaFuncSetAttribute((const void*)generic_graph_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      generic_graph_kernel<true><<<grid, 32, smem, stream>>>(a);
    }
    else
    {
      CU_CHECK(cudaFuncSetAttribute((const void*)generic_graph_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      generic_graph_kernel<false><<<grid, 32, smem, stream>>>(a);
    }
    ++g_launches;
  }
  if (!capturing)
  {
    cudaEventRecord(g->ev1, stream);
    g->timed = true;
  }
  CU_CHECK(cudaGetLastError());
  if (g->fdn_node >= 0 || g->has_dmem) g->blocks_done += T;
  if (mix_dev)
  {
    // asynchronous mix bus: the reduction leaves the caller's stream here and runs beside the next call's kernel
    cudaStream_t ms = stream;
    if (g->mix_async)
    {
      CU_CHECK(cudaEventRecord(g->ev_main, stream));
      CU_CHECK(cudaStreamWaitEvent(g->s_mix, g->ev_main, 0));
      ms = g->s_mix;
    }
    float* scratch = g->partial_cur + (size_t)T * std::max(1, n_out) * n_groups * MLB_BLOCK;
    MixBusArgs ba;
    rc = bus_args(g, T * std::max(1, n_out), &ba, ms);
    if (rc != MLB_OK) return rc;
    // (asynchronous mode, measured on config A: 1024-thread CTAs do not fit beside a resident chain CTA and run in
    // the gaps between chain kernels: 0.3820 -> 0.3775 ms per step; 256-thread CTAs that DO run beside it only slow
    // the HBM-bound chain kernel down by what they read: 0.3803 ms.  The reduction costs bandwidth, not latency.)
    mix_reduce_kernel<<<T * std::max(1, n_out), dim3(MLB_BLOCK, 16), 0, ms>>>(g->partial_cur, scratch, mix_dev,
                                                                               n_groups, ba);
    ++g_launches;
    CU_CHECK(cudaGetLastError());
    if (g->mix_async)
    {
      CU_CHECK(cudaEventRecord(g->ev_mixdone[g->mix_seq & 3u], ms));
      g->mix_pending = true;
    }
    rc = bus_complete(g, ba, T * std::max(1, n_out), mix_dev, ms);
    if (rc != MLB_OK) return rc;
  }
  return MLB_OK;
}

extern "C" int mlb_graph_set_mix_async(mlb_graph* g, int on)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  if (g->mix_pending) CU_CHECK(cudaStreamSynchronize(g->s_mix));
  g->mix_pending = false;
  if (on && !g->s_mix)
  {
    CU_CHECK(cudaStreamCreateWithFlags(&g->s_mix, cudaStreamNonBlocking));
    CU_CHECK(cudaEventCreateWithFlags(&g->ev_main, cudaEventDisableTiming));
    for (cudaEvent_t& e : g->ev_mixdone) CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  g->mix_async = on ? 1 : 0;
  return MLB_OK;
}

extern "C" int mlb_graph_process_host(mlb_graph* g, const float* in_host, float* out_host,
                                      float* mix_host, int n_blocks)
{
  if (!g) return fail(MLB_ERR_INVALID, "null graph");
  if (n_blocks <= 0) return fail(MLB_ERR_INVALID, "n_blocks must be positive");
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  const size_t V = (size_t)g->V, T = (size_t)n_blocks;
  const size_t n_in = (size_t)g->layout.n_inputs, n_out = g->outs.size();
  const size_t in_bytes = T * n_in * V * MLB_BLOCK * 4, out_bytes = T * n_out * V * MLB_BLOCK * 4;
  const size_t mix_bytes = T * n_out * MLB_BLOCK * 4;
  if (n_in && !in_host) return fail(MLB_ERR_INVALID, "graph has INPUT nodes but in is null");
  if (n_in && (rc = ensure_buf(&g->d_in, &g->in_cap, in_bytes)) != MLB_OK) return rc;
  if (out_host && (rc = ensure_buf(&g->d_out, &g->out_cap, out_bytes)) != MLB_OK) return rc;
  if (mix_host && (rc = ensure_buf(&g->d_mix, &g->mix_cap, mix_bytes)) != MLB_OK) return rc;
  cudaStream_t s = g->stream;
  // Large fused banks: pipeline over voice slices so that the H2D copy of slice p+1, the kernel
  // of slice p and the D2H copy of slice p-1 overlap (three streams; PCIe is full duplex).
  // A slice is a strided window [T*n][va:vb][64] of the host planes -> cudaMemcpy2DAsync.
  const int n_slices_env = env_int("MLB_HOST_SLICES", 16);
  // Only planes that really cross PCIe count: a mix-only call (out_host == NULL) moves 256 B per
  // block and must stay ONE launch -- 16 slice launches of a 4096-voice bank are latency-bound.
  const size_t pcie_bytes = in_bytes + (out_host ? out_bytes : 0);
  const size_t slice_min_bytes = (size_t)env_int("MLB_HOST_SLICE_MIN_MB", 32) << 20;
  g->last_host_slices = 1;
  if (g->kind == KIND_FUSED && n_slices_env > 1 && g->V >= 4096 && pcie_bytes >= slice_min_bytes)
  {
    const int n_slices = std::min(n_slices_env, (int)mlb_graph::kMaxHostSlices);
    int per = (g->V + n_slices - 1) / n_slices;
    per = (per + 31) / 32 * 32;
    const size_t pitch = V * MLB_BLOCK * 4;
    if (mix_host && (rc = ensure_partial(g, n_blocks, (g->V + 31) / 32)) != MLB_OK) return rc;
    if (mix_host && (rc = mlb_graph_mix_wait(g, s)) != MLB_OK) return rc;  // an asynchronous reduce may still read the partials
    int chunks = 0;
    cudaEventRecord(g->ev0, s);
    g->last_host_slices = 0;
    for (int p = 0, va = 0; va < g->V; ++p, va += per)
    {
      ++g->last_host_slices;
      const int vb = std::min(g->V, va + per);
      const size_t width = (size_t)(vb - va) * MLB_BLOCK * 4;
      if (n_in)
      {
        CU_CHECK(cudaMemcpy2DAsync(g->d_in + (size_t)va * MLB_BLOCK, pitch, in_host + (size_t)va * MLB_BLOCK,
                                   pitch, width, T * n_in, cudaMemcpyHostToDevice, g->s_h2d));
        CU_CHECK(cudaEventRecord(g->ev_up[p], g->s_h2d));
        CU_CHECK(cudaStreamWaitEvent(s, g->ev_up[p], 0));
      }
      rc = launch_chain_slice(g, n_in ? g->d_in : nullptr, out_host ? g->d_out : nullptr, mix_host != nullptr,
                              n_blocks, s, va, vb, chunks);
      if (rc != MLB_OK) return rc;
      chunks = g->launch_chunks;
      if (out_host)
      {
        CU_CHECK(cudaEventRecord(g->ev_k[p], s));
        CU_CHECK(cudaStreamWaitEvent(g->s_d2h, g->ev_k[p], 0));
        CU_CHECK(cudaMemcpy2DAsync(out_host + (size_t)va * MLB_BLOCK, pitch, g->d_out + (size_t)va * MLB_BLOCK,
                                   pitch, width, T * n_out, cudaMemcpyDeviceToHost, g->s_d2h));
      }
    }
    cudaEventRecord(g->ev1, s);
    g->timed = true;
    if (mix_host)
    {
      const int n_groups = (g->V + 31) / 32;
      float* scratch = g->partial_cur + T * std::max<size_t>(1, n_out) * n_groups * MLB_BLOCK;
      MixBusArgs ba;
      rc = bus_args(g, (int)(T * std::max<size_t>(1, n_out)), &ba, s);
      if (rc != MLB_OK) return rc;
      mix_reduce_kernel<<<(int)(T * std::max<size_t>(1, n_out)), dim3(MLB_BLOCK, 16), 0, s>>>(g->partial_cur, scratch,
                                                                                            g->d_mix, n_groups, ba);
      rc = bus_complete(g, ba, (int)(T * std::max<size_t>(1, n_out)), g->d_mix, s);
      if (rc != MLB_OK) return rc;
      rc = mlb_graph_mix_wait(g, s);  // the host entry point returns finished results
      if (rc != MLB_OK) return rc;
      ++g_launches;
      CU_CHECK(cudaGetLastError());
      CU_CHECK(cudaMemcpyAsync(mix_host, g->d_mix, mix_bytes, cudaMemcpyDeviceToHost, s));
    }
    CU_CHECK(cudaStreamSynchronize(s));
    CU_CHECK(cudaStreamSynchronize(g->s_d2h));
    return MLB_OK;
  }
  if (n_in) CU_CHECK(cudaMemcpyAsync(g->d_in, in_host, in_bytes, cudaMemcpyHostToDevice, s));
  rc = mlb_graph_process_device(g, n_in ? g->d_in : nullptr, out_host ? g->d_out : nullptr,
                                mix_host ? g->d_mix : nullptr, n_blocks, s);
  if (rc != MLB_OK) return rc;
  if (out_host) CU_CHECK(cudaMemcpyAsync(out_host, g->d_out, out_bytes, cudaMemcpyDeviceToHost, s));
  if (mix_host)
  {
    rc = mlb_graph_mix_wait(g, s);  // async mix bus: the completion runs on a side stream
    if (rc != MLB_OK) return rc;
    CU_CHECK(cudaMemcpyAsync(mix_host, g->d_mix, mix_bytes, cudaMemcpyDeviceToHost, s));
  }
  CU_CHECK(cudaStreamSynchronize(s));
  return MLB_OK;
}

extern "C" int mlb_graph_last_host_slices(const mlb_graph* g) { return g ? g->last_host_slices : -1; }

extern "C" int mlb_graph_last_kernel_ms(mlb_graph* g, float* ms)
{
  if (!g || !ms) return fail(MLB_ERR_INVALID, "null");
  if (!g->timed) return fail(MLB_ERR_INVALID, "no launch recorded yet");
  CU_CHECK(cudaEventSynchronize(g->ev1));
  CU_CHECK(cudaEventElapsedTime(ms, g->ev0, g->ev1));
  return MLB_OK;
}

// ------------------------------------------------------------------------------------------
// stateless elementwise ops

extern "C" int mlb_map_device(int op, const float* x1, const float* x2, const float* x3, float* y,
                              size_t n_rows, void* stream)
{
  int nin, nst, nco;
  if (mlb_op_info(op, &nin, &nst, &nco) != MLB_OK || nst != 0 || nco != 0 || nin < 1 ||
      op == MLB_OP_FDN8_R)
    return fail(MLB_ERR_INVALID, "op %d is not a stateless elementwise op", op);
  if (!x1 || !y || (nin >= 2 && !x2) || (nin >= 3 && !x3)) return fail(MLB_ERR_INVALID, "null operand");
  if (n_rows == 0) return MLB_OK;
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  const size_t n4 = n_rows * (MLB_BLOCK / 4);
  const int threads = 256;
  const size_t want = (n4 + threads - 1) / threads;
  const int blocks = (int)std::min<size_t>(want, (size_t)g_sm_count * 16);
  MapKernelFn fn = map_kernel_for(op);
  if (!fn) return fail(MLB_ERR_INVALID, "op %d has no map kernel", op);
  fn<<<blocks, threads, 0, (cudaStream_t)stream>>>((const float4*)x1, nin >= 2 ? (const float4*)x2 : nullptr,
                                                   nin >= 3 ? (const float4*)x3 : nullptr, (float4*)y, n4);
  ++g_launches;
  CU_CHECK(cudaGetLastError());
  return MLB_OK;
}

// Device staging of mlb_map_host: four buffers (three operands + result) that only ever grow, so that the
// host value-type operators (`a + b` on two mlb::DSPVectors) do not allocate after their first use.
#include <mutex>
namespace
{
struct MapPool
{
  std::mutex mu;
  float* d[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap = 0;
  cudaStream_t stream = nullptr;
  long long allocations = 0;
};
MapPool g_map_pool;
}  // namespace
extern "C" long long mlb_map_host_allocations(void) { return g_map_pool.allocations; }

extern "C" int mlb_map_host(int op, const float* x1, const float* x2, const float* x3, float* y,
                            size_t n_rows)
{
  int nin, nst, nco;
  if (mlb_op_info(op, &nin, &nst, &nco) != MLB_OK) return fail(MLB_ERR_INVALID, "unknown op %d", op);
  if (n_rows == 0) return MLB_OK;
  if (!x1 || !y || (nin >= 2 && !x2) || (nin >= 3 && !x3)) return fail(MLB_ERR_INVALID, "null operand");
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  const size_t bytes = n_rows * MLB_BLOCK * 4;
  MapPool& P = g_map_pool;
  std::lock_guard<std::mutex> lock(P.mu);
  if (!P.stream) CU_CHECK(cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking));
  if (bytes > P.cap)
  {
    const size_t want = std::max(bytes, std::max<size_t>(P.cap * 2, (size_t)64 << 10));
    for (float*& p : P.d)
    {
      cudaFree(p);
      p = nullptr;
    }
    P.cap = 0;
    for (float*& p : P.d)
      if (cudaMalloc(&p, want) != cudaSuccess) return fail(MLB_ERR_ALLOC, "cudaMalloc of %zu B failed", want);
    P.cap = want;
    ++P.allocations;
  }
  const float* h[3] = {x1, x2, x3};
  for (int k = 0; k < nin && k < 3; ++k)
    CU_CHECK(cudaMemcpyAsync(P.d[k], h[k], bytes, cudaMemcpyHostToDevice, P.stream));
  rc = mlb_map_device(op, P.d[0], nin >= 2 ? P.d[1] : nullptr, nin >= 3 ? P.d[2] : nullptr, P.d[3], n_rows, P.stream);
  if (rc != MLB_OK) return rc;
  CU_CHECK(cudaMemcpyAsync(y, P.d[3], bytes, cudaMemcpyDeviceToHost, P.stream));
  CU_CHECK(cudaStreamSynchronize(P.stream));
  return MLB_OK;
}


// ------------------------------------------------------------------------------------------
// EventsToSignals::Voice bank (K7, voice_kernel.cuh)

struct mlb_voices
{
  int V = 0;
  float sr = 0.f;
  float gl_per = 0.f, gl_dy = 0.f, dr_per = 0.f, dr_dy = 0.f, pc_per = 0.f, pc_dy = 0.f;
  unsigned flags = 0;
  uint32_t* d_state = nullptr;
  float* d_coef = nullptr;
  float* d_grows = nullptr;
  int32_t* d_main = nullptr;  // MPE: index of each voice's main voice, or -1
  mlb_voice_events* d_ev = nullptr;
  float* d_out = nullptr;
  size_t ev_cap = 0, out_cap = 0;
  cudaStream_t stream = nullptr;
};

extern "C" int mlb_voices_destroy(mlb_voices* vb)
{
  if (!vb) return MLB_OK;
  cudaFree(vb->d_state);
  cudaFree(vb->d_coef);
  cudaFree(vb->d_grows);
  cudaFree(vb->d_main);
  cudaFree(vb->d_ev);
  cudaFree(vb->d_out);
  if (vb->stream) cudaStreamDestroy(vb->stream);
  delete vb;
  --g_live_handles;
  return MLB_OK;
}

extern "C" int mlb_voices_create(int n_voices, float sample_rate, const int32_t* voice_index,
                                 const float* pitch_glide_seconds, const float* drift_amount,
                                 const float* pitch_bend, unsigned flags, mlb_voices** out)
{
  if (!out) return fail(MLB_ERR_INVALID, "out is null");
  *out = nullptr;
  if (n_voices <= 0 || !(sample_rate > 0.f) || !voice_index || !pitch_glide_seconds || !drift_amount || !pitch_bend)
    return fail(MLB_ERR_INVALID, "bad argument");
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  const size_t V = (size_t)n_voices;
  mlb_voices* vb = new mlb_voices;
  ++g_live_handles;
  vb->V = n_voices;
  vb->sr = sample_rate;
  // the recalc block of the first beginProcess (MLEventsToSignals.cpp:92-113), done once here
  const double sr = (double)sample_rate;
  float c[2];
  mlb_coeffs_glide((float)(sr * 0.02f), c);  // kGlideTimeSeconds
  vb->gl_per = c[0], vb->gl_dy = c[1];
  mlb_coeffs_glide((float)(sr * 8.0f), c);   // kDriftTimeSeconds
  vb->dr_per = c[0], vb->dr_dy = c[1];
  mlb_coeffs_glide((float)(int)(sr * 0.02f), c);  // SmoothedController::process, .cpp:278-280
  vb->pc_per = c[0], vb->pc_dy = c[1];
  vb->flags = flags;
  std::vector<uint32_t> st((size_t)VS_COUNT * V, 0u);
  std::vector<float> co((size_t)VC_COUNT * V, 0.f);
  auto fbits = [](float f)
  {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
  };
  for (size_t v = 0; v < V; ++v)
  {
    const int glide_samples = (int)(sr * pitch_glide_seconds[v]);  // pitchGlideTimeInSamples, :97
    float sg[2];
    mlb_coeffs_sample_glide((float)glide_samples, sg);             // pitchGlide.setGlideTimeInSamples, :101
    st[(size_t)VS_PG_REM * V + v] = 0xFFFFFFFFu;                    // mSamplesRemaining{-1}
    st[(size_t)VS_PG_PER * V + v] = fbits(sg[0]);
    st[(size_t)VS_PG_DY * V + v] = fbits(sg[1]);
    for (int g = 0; g < VG_COUNT; ++g)                              // reset(): setValue(0) on all but the drift glide, :80-84
      st[(size_t)(VS_GL + 4 * g + 2) * V + v] = (g == VG_DRIFT || g == VG_PRESSURE) ? 0xFFFFFFFFu : 0u;
    st[(size_t)VS_SEED * V + v] = (uint32_t)(voice_index[v] * 232);  // :62
    co[(size_t)VC_GLIDE_SAMPLES * V + v] = (float)glide_samples;
    co[(size_t)VC_DRIFT_AMOUNT * V + v] = drift_amount[v];
    co[(size_t)VC_BEND_RANGE * V + v] = pitch_bend[v];
    co[(size_t)VC_VOICE_ROW * V + v] = (float)voice_index[v] - 1;   // :292
  }
  if (cudaMalloc(&vb->d_state, st.size() * 4) != cudaSuccess || cudaMalloc(&vb->d_coef, co.size() * 4) != cudaSuccess ||
      cudaMalloc(&vb->d_grows, (size_t)VG_COUNT * V * MLB_BLOCK * 4) != cudaSuccess)
  {
    mlb_voices_destroy(vb);
    return fail(MLB_ERR_ALLOC, "cudaMalloc of voice bank failed");
  }
  cudaMemcpy(vb->d_state, st.data(), st.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(vb->d_coef, co.data(), co.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(vb->d_grows, 0, (size_t)VG_COUNT * V * MLB_BLOCK * 4);
  cudaStreamCreateWithFlags(&vb->stream, cudaStreamNonBlocking);
  *out = vb;
  return MLB_OK;
}

extern "C" int mlb_voices_set_main_voices(mlb_voices* vb, const int32_t* main_voice)
{
  if (!vb) return fail(MLB_ERR_INVALID, "null voice bank");
  if (!main_voice)
  {
    cudaFree(vb->d_main);
    vb->d_main = nullptr;
    return MLB_OK;
  }
  for (int v = 0; v < vb->V; ++v)
    if (main_voice[v] >= vb->V || (main_voice[v] >= 0 && main_voice[main_voice[v]] >= 0))
      return fail(MLB_ERR_INVALID, "main_voice[%d] = %d is not a main voice of this bank", v, main_voice[v]);
  if (!vb->d_main && cudaMalloc(&vb->d_main, (size_t)vb->V * 4) != cudaSuccess)
    return fail(MLB_ERR_ALLOC, "cudaMalloc of main-voice table failed");
  CU_CHECK(cudaMemcpy(vb->d_main, main_voice, (size_t)vb->V * 4, cudaMemcpyHostToDevice));
  return MLB_OK;
}

extern "C" int mlb_voices_process_device(mlb_voices* vb, const mlb_voice_events* events_dev, float* out_dev,
                                         int n_blocks, unsigned row_mask, void* stream)
{
  if (!vb || !events_dev || (!out_dev && (row_mask & 0xFFu))) return fail(MLB_ERR_INVALID, "null argument");
  if (n_blocks <= 0) return fail(MLB_ERR_INVALID, "n_blocks must be positive");
  VoiceArgs a;
  a.ev = events_dev;
  a.out = out_dev;
  a.state = vb->d_state;
  a.coef = vb->d_coef;
  a.grows = vb->d_grows;
  a.V = vb->V, a.T = n_blocks;
  a.row_mask = row_mask & 0xFFu;
  a.sr = vb->sr;
  a.gl_per = vb->gl_per, a.gl_dy = vb->gl_dy, a.dr_per = vb->dr_per, a.dr_dy = vb->dr_dy;
  a.pc_per = vb->pc_per, a.pc_dy = vb->pc_dy;
  a.midi = (vb->flags & MLB_VOICES_MIDI) ? 1 : 0;
  // warps x {gate, pitch (, elapsed time)} 8 KB row tiles; 7 warps per CTA and two CTAs per SM put the
  // 2 048 warps of a 65 536-voice bank on the 148 SMs in one wave (voice_kernel.cuh)
  const bool want_time = (row_mask & 128u) != 0;
  const int wpc_max = want_time ? kVoiceWarpsPerCtaTime : kVoiceWarpsPerCta;
  const int wpc = env_int("MLB_VOICE_WARPS", wpc_max);
  if (wpc < 1 || wpc > wpc_max) return fail(MLB_ERR_INVALID, "MLB_VOICE_WARPS must be in [1, %d]", wpc_max);
  const size_t smem = (size_t)wpc * (want_time ? 3 : 2) * kVoiceTileFloats * sizeof(float);
  const void* fn = want_time ? (const void*)voice_bank_kernel<true> : (const void*)voice_bank_kernel<false>;
  {
    int rc = ensure_func_smem(fn, smem);
    if (rc != MLB_OK) return rc;
  }
  const int n_warps = (vb->V + 31) / 32;
  if (want_time)
    voice_bank_kernel<true><<<(n_warps + wpc - 1) / wpc, 32 * wpc, smem, (cudaStream_t)stream>>>(a);
  else
    voice_bank_kernel<false><<<(n_warps + wpc - 1) / wpc, 32 * wpc, smem, (cudaStream_t)stream>>>(a);
  ++g_launches;
  CU_CHECK(cudaGetLastError());
  if (vb->d_main && (a.row_mask & 0x79u))  // MPE: add the main voices' pitch / z / x / y / mod rows
  {
    const size_t n4 = (size_t)n_blocks * vb->V * 16;
    const int grid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)g_sm_count * 16);
    voice_mpe_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out_dev, vb->d_main, vb->V, n_blocks, a.row_mask);
    ++g_launches;
    CU_CHECK(cudaGetLastError());
  }
  return MLB_OK;
}

extern "C" int mlb_voices_process_host(mlb_voices* vb, const mlb_voice_events* events_host, float* out_host,
                                       int n_blocks, unsigned row_mask)
{
  if (!vb || !events_host) return fail(MLB_ERR_INVALID, "null argument");
  if (n_blocks <= 0) return fail(MLB_ERR_INVALID, "n_blocks must be positive");
  row_mask &= 0xFFu;
  if (row_mask && !out_host) return fail(MLB_ERR_INVALID, "out is null");
  const size_t V = (size_t)vb->V, T = (size_t)n_blocks;
  const size_t ev_bytes = T * V * sizeof(mlb_voice_events), out_bytes = T * MLB_VOICE_ROWS * V * MLB_BLOCK * 4;
  int rc = ensure_buf(reinterpret_cast<float**>(&vb->d_ev), &vb->ev_cap, ev_bytes);
  if (rc != MLB_OK) return rc;
  rc = ensure_buf(&vb->d_out, &vb->out_cap, out_bytes);
  if (rc != MLB_OK) return rc;
  cudaStream_t s = vb->stream;
  CU_CHECK(cudaMemcpyAsync(vb->d_ev, events_host, ev_bytes, cudaMemcpyHostToDevice, s));
  rc = mlb_voices_process_device(vb, vb->d_ev, vb->d_out, n_blocks, row_mask, s);
  if (rc != MLB_OK) return rc;
  const size_t plane = V * MLB_BLOCK * 4;
  if (row_mask == 0xFFu)
    CU_CHECK(cudaMemcpyAsync(out_host, vb->d_out, out_bytes, cudaMemcpyDeviceToHost, s));
  else
    for (size_t t = 0; t < T; ++t)
      for (int r = 0; r < MLB_VOICE_ROWS; ++r)
        if (row_mask & (1u << r))
          CU_CHECK(cudaMemcpyAsync(reinterpret_cast<char*>(out_host) + (t * MLB_VOICE_ROWS + r) * plane,
                                   reinterpret_cast<char*>(vb->d_out) + (t * MLB_VOICE_ROWS + r) * plane, plane,
                                   cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaStreamSynchronize(s));
  return MLB_OK;
}


// Events -> signals -> chain with nothing but the event records crossing PCIe ("contract E"):
// the Voice bank writes the rows the graph reads (EventsToSignals::processVector,
// MLEventsToSignals.cpp:383-470) into a device buffer that IS the graph's input (the bank's out planes
// have the layout of graph inputs, plane r = row r), and the graph runs on it in the same stream.
// Pipelined over TIME chunks -- blocks are contiguous in every buffer involved, and both the bank and
// the graph carry their state from launch to launch -- so that the H2D copy of chunk c+1's records, the
// two kernels of chunk c and the D2H copy of chunk c-1's output rows overlap.
extern "C" int mlb_synth_process_host(mlb_voices* vb, mlb_graph* g, const mlb_voice_events* events_host,
                                      float* out_host, float* mix_host, int n_blocks)
{
  if (!vb || !g || !events_host) return fail(MLB_ERR_INVALID, "null argument");
  if (n_blocks <= 0) return fail(MLB_ERR_INVALID, "n_blocks must be positive");
  if (vb->V != g->V) return fail(MLB_ERR_INVALID, "voice bank has %d voices, graph has %d", vb->V, g->V);
  unsigned row_mask = 0;
  for (const mlb_node& n : g->nodes)
    if (n.op == MLB_OP_INPUT)
    {
      if (n.iarg < 0 || n.iarg >= MLB_VOICE_ROWS)
        return fail(MLB_ERR_INVALID, "graph INPUT plane %d is not a Voice row (0..%d)", n.iarg, MLB_VOICE_ROWS - 1);
      row_mask |= 1u << n.iarg;
    }
  if (!row_mask) return fail(MLB_ERR_INVALID, "graph reads no Voice row (no INPUT node)");
  if (g->outs.empty() && (out_host || mix_host)) return fail(MLB_ERR_INVALID, "graph has no outputs");
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  const size_t V = (size_t)g->V, T = (size_t)n_blocks, n_out = g->outs.size();
  const size_t plane = V * MLB_BLOCK * 4;
  // rows of one chunk stay under ~1 GB (8 planes per block are addressed, only the masked ones are touched)
  int Tc = env_int("MLB_SYNTH_CHUNK_BLOCKS", 0);
  if (Tc <= 0) Tc = (int)std::max<size_t>(1, ((size_t)1 << 30) / (MLB_VOICE_ROWS * plane));
  Tc = std::min<int>(Tc, n_blocks);
  Tc = std::max<int>(Tc, (n_blocks + mlb_graph::kMaxHostSlices - 1) / mlb_graph::kMaxHostSlices);
  const size_t ev_bytes = T * V * sizeof(mlb_voice_events);
  if ((rc = ensure_buf(reinterpret_cast<float**>(&vb->d_ev), &vb->ev_cap, ev_bytes)) != MLB_OK) return rc;
  if ((rc = ensure_buf(&vb->d_out, &vb->out_cap, (size_t)Tc * MLB_VOICE_ROWS * plane)) != MLB_OK) return rc;
  if (out_host && (rc = ensure_buf(&g->d_out, &g->out_cap, T * n_out * plane)) != MLB_OK) return rc;
  if (mix_host && (rc = ensure_buf(&g->d_mix, &g->mix_cap, T * n_out * MLB_BLOCK * 4)) != MLB_OK) return rc;
  const int saved_inputs = g->layout.n_inputs;
  g->layout.n_inputs = MLB_VOICE_ROWS;
  cudaStream_t s = g->stream;
  rc = MLB_OK;
  int p = 0;
  for (size_t t0 = 0; t0 < T && rc == MLB_OK; t0 += (size_t)Tc, ++p)
  {
    const int tc = (int)std::min<size_t>((size_t)Tc, T - t0);
    cudaError_t ce = cudaMemcpyAsync(vb->d_ev + t0 * V, events_host + t0 * V, (size_t)tc * V * sizeof(mlb_voice_events),
                                     cudaMemcpyHostToDevice, g->s_h2d);
    if (ce == cudaSuccess) ce = cudaEventRecord(g->ev_up[p], g->s_h2d);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s, g->ev_up[p], 0);
    if (ce != cudaSuccess)
    {
      rc = fail(MLB_ERR_CUDA, "events H2D: %s", cudaGetErrorString(ce));
      break;
    }
    rc = mlb_voices_process_device(vb, vb->d_ev + t0 * V, vb->d_out, tc, row_mask, s);
    if (rc != MLB_OK) break;
    rc = mlb_graph_process_device(g, vb->d_out, out_host ? g->d_out + t0 * n_out * V * MLB_BLOCK : nullptr,
                                  mix_host ? g->d_mix + t0 * n_out * MLB_BLOCK : nullptr, tc, s);
    if (rc != MLB_OK) break;
    if (out_host)
    {
      ce = cudaEventRecord(g->ev_k[p], s);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(g->s_d2h, g->ev_k[p], 0);
      if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(out_host + t0 * n_out * V * MLB_BLOCK, g->d_out + t0 * n_out * V * MLB_BLOCK,
                             (size_t)tc * n_out * plane, cudaMemcpyDeviceToHost, g->s_d2h);
      if (ce != cudaSuccess) rc = fail(MLB_ERR_CUDA, "rows D2H: %s", cudaGetErrorString(ce));
    }
  }
  g->layout.n_inputs = saved_inputs;
  g->last_host_slices = p;
  if (rc == MLB_OK && mix_host)
  {
    rc = mlb_graph_mix_wait(g, s);
    if (rc == MLB_OK && cudaMemcpyAsync(mix_host, g->d_mix, T * n_out * MLB_BLOCK * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess)
      rc = fail(MLB_ERR_CUDA, "mix D2H failed");
  }
  // drain all three streams whatever happened: nothing may still touch the caller's buffers on return
  cudaStreamSynchronize(g->s_h2d);
  cudaError_t e1 = cudaStreamSynchronize(s), e2 = cudaStreamSynchronize(g->s_d2h);
  if (rc == MLB_OK && (e1 != cudaSuccess || e2 != cudaSuccess))
    rc = fail(MLB_ERR_CUDA, "synth pipeline: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  return rc;
}

// ------------------------------------------------------------------------------------------
// Upsampler / Downsampler banks (K8, resample_kernel.cuh)

struct mlb_resampler
{
  int dir = 0, oct = 0, V = 0;
  unsigned counter = 0;
  uint32_t* d_state = nullptr;
  float* d_buf = nullptr;
  float *d_in = nullptr, *d_out = nullptr;
  size_t in_cap = 0, out_cap = 0;
  cudaStream_t stream = nullptr;
};

extern "C" int mlb_resampler_destroy(mlb_resampler* r)
{
  if (!r) return MLB_OK;
  cudaFree(r->d_state);
  cudaFree(r->d_buf);
  cudaFree(r->d_in);
  cudaFree(r->d_out);
  if (r->stream) cudaStreamDestroy(r->stream);
  delete r;
  --g_live_handles;
  return MLB_OK;
}
extern "C" int mlb_resampler_clear(mlb_resampler* r)
{
  if (!r) return fail(MLB_ERR_INVALID, "null resampler");
  r->counter = 0;
  CU_CHECK(cudaMemset(r->d_state, 0, (size_t)r->oct * 9 * r->V * 4));
  if (r->d_buf) CU_CHECK(cudaMemset(r->d_buf, 0, (size_t)(2 * r->oct + 1) * r->V * MLB_BLOCK * 4));
  return MLB_OK;
}
extern "C" int mlb_resampler_create(int direction, int octaves, int n_voices, mlb_resampler** out)
{
  if (!out) return fail(MLB_ERR_INVALID, "out is null");
  *out = nullptr;
  if ((direction != MLB_RESAMPLE_UP && direction != MLB_RESAMPLE_DOWN) || octaves < 1 || octaves > 4 || n_voices <= 0)
    return fail(MLB_ERR_INVALID, "bad argument (direction %d, octaves %d, voices %d)", direction, octaves, n_voices);
  int rc = ensure_init();
  if (rc != MLB_OK) return rc;
  mlb_resampler* r = new mlb_resampler;
  ++g_live_handles;
  r->dir = direction, r->oct = octaves, r->V = n_voices;
  bool ok = cudaMalloc(&r->d_state, (size_t)octaves * 9 * n_voices * 4) == cudaSuccess;
  if (ok && direction == MLB_RESAMPLE_DOWN)
    ok = cudaMalloc(&r->d_buf, (size_t)(2 * octaves + 1) * n_voices * MLB_BLOCK * 4) == cudaSuccess;
  if (!ok)
  {
    mlb_resampler_destroy(r);
    return fail(MLB_ERR_ALLOC, "cudaMalloc of resampler state failed");
  }
  cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
  rc = mlb_resampler_clear(r);
  if (rc != MLB_OK)
  {
    mlb_resampler_destroy(r);
    return rc;
  }
  *out = r;
  return MLB_OK;
}
static int resampler_out_blocks(const mlb_resampler* r, int n_in)
{
  if (r->dir == MLB_RESAMPLE_UP) return n_in << r->oct;
  return (int)(((long long)r->counter + n_in) >> r->oct);
}
extern "C" int mlb_resampler_process_device(mlb_resampler* r, const float* in_dev, float* out_dev, int n_blocks_in,
                                            int* n_blocks_out, void* stream)
{
  if (!r || !in_dev) return fail(MLB_ERR_INVALID, "null argument");
  if (n_blocks_in <= 0) return fail(MLB_ERR_INVALID, "n_blocks_in must be positive");
  const int n_out = resampler_out_blocks(r, n_blocks_in);
  if (n_out > 0 && !out_dev) return fail(MLB_ERR_INVALID, "out is null");
  ResampleArgs a;
  a.in = in_dev, a.out = out_dev, a.state = r->d_state, a.buf = r->d_buf;
  a.V = r->V, a.T = n_blocks_in, a.oct = r->oct, a.counter = r->counter;
  const int grid = (r->V + 127) / 128;
  if (r->dir == MLB_RESAMPLE_UP)
    upsample_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  else
  {
    downsample_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    r->counter = (r->counter + (unsigned)n_blocks_in) & ((1u << r->oct) - 1u);
  }
  ++g_launches;
  CU_CHECK(cudaGetLastError());
  if (n_blocks_out) *n_blocks_out = n_out;
  return MLB_OK;
}
extern "C" int mlb_resampler_process_host(mlb_resampler* r, const float* in_host, float* out_host, int n_blocks_in,
                                          int* n_blocks_out)
{
  if (!r || !in_host) return fail(MLB_ERR_INVALID, "null argument");
  if (n_blocks_in <= 0) return fail(MLB_ERR_INVALID, "n_blocks_in must be positive");
  const size_t row = (size_t)r->V * MLB_BLOCK * 4;
  const int n_out = resampler_out_blocks(r, n_blocks_in);
  if (n_out > 0 && !out_host) return fail(MLB_ERR_INVALID, "out is null");
  int rc = ensure_buf(&r->d_in, &r->in_cap, (size_t)n_blocks_in * row);
  if (rc != MLB_OK) return rc;
  rc = ensure_buf(&r->d_out, &r->out_cap, std::max<size_t>(1, (size_t)n_out) * row);
  if (rc != MLB_OK) return rc;
  cudaStream_t s = r->stream;
  CU_CHECK(cudaMemcpyAsync(r->d_in, in_host, (size_t)n_blocks_in * row, cudaMemcpyHostToDevice, s));
  rc = mlb_resampler_process_device(r, r->d_in, r->d_out, n_blocks_in, n_blocks_out, s);
  if (rc != MLB_OK) return rc;
  if (n_out > 0) CU_CHECK(cudaMemcpyAsync(out_host, r->d_out, (size_t)n_out * row, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaStreamSynchronize(s));
  return MLB_OK;
}


// ------------------------------------------------------------------------------------------
// C face of mlb::VoiceRouter (host only)
#include "../../include/mlb200_events.hpp"

struct mlb_router
{
  mlb::VoiceRouter router;
  mlb_router(int polyphony, int protocol)
      : router(polyphony, protocol == 1 ? mlb::VoiceRouter::kMPE : mlb::VoiceRouter::kMIDI)
  {
  }
};
extern "C" mlb_router* mlb_router_create(int polyphony, int protocol)
{
  if (polyphony < 1 || (protocol != 0 && protocol != 1)) return nullptr;
  return new mlb_router(polyphony, protocol);
}
extern "C" void mlb_router_destroy(mlb_router* r) { delete r; }
extern "C" void mlb_router_set_unison(mlb_router* r, int on)
{
  if (r) r->router.setUnison(on != 0);
}
extern "C" void mlb_router_add_event(mlb_router* r, const mlb_event* e)
{
  if (!r || !e) return;
  mlb::Event ev;
  ev.type = e->type, ev.channel = e->channel, ev.sourceIdx = e->source_idx, ev.time = e->time;
  ev.value1 = e->value1, ev.value2 = e->value2;
  r->router.addEvent(ev);
}
extern "C" void mlb_router_clear_events(mlb_router* r)
{
  if (r) r->router.clearEvents();
}
extern "C" int mlb_router_unsupported_count(const mlb_router* r) { return r ? r->router.unsupportedEvents() : -1; }
extern "C" void mlb_router_set_mod_cc(mlb_router* r, int cc)
{
  if (r) r->router.setModCC(cc);
}
extern "C" int mlb_router_record_count(const mlb_router* r) { return r ? r->router.recordCount() : -1; }
extern "C" int mlb_router_process_vector(mlb_router* r, int start_time, mlb_voice_events* records)
{
  if (!r || !records) return -1;
  return r->router.processVector(start_time, records);
}
