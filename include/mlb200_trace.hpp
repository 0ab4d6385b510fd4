// mlb200_trace.hpp -- the reference's functor / operator spelling on the host, as a TRACING layer.
//
// The reference's hot path is written by its users like this (examples/audio-and-midi/sine.cpp:21-35,
// reverb.cpp:68-123):
//
//     void process(AudioContext* ctx, void* state) {
//       auto s = static_cast<MyState*>(state);
//       ctx->outputs[0] = s->lp(s->osc(220.f / kSampleRate)) * kOutputGain;
//     }
//
// i.e. functor objects held in a state struct, called with DSPVector values, chained with operators
// (MLDSPGens.h:373-381, MLDSPFilters.h:51-153, MLDSPOps.h:337-352, MLSignalProcessBuffer.h:18,
// MLAudioContext.h:23-101).  In namespace mlb::tr the same spellings exist, but a DSPVector is SYMBOLIC:
// running the process function records the DAG of functors it applies instead of computing samples.
// The recorded graph is then compiled into one fused GPU kernel launch per call for V independent
// instances (mlb_graph_*), which is what "source/procs' runtime graph evaluator becomes a launcher"
// means in this project.  So a reference process function is ported by changing
//
//     using namespace ml;      ->      using namespace mlb::tr;
//
// and handing it to mlb::tr::TracedProcessor instead of ml::AudioTask / ml::SignalProcessBuffer.
//
// How state carried between calls is found: the process function is traced TWICE.  A DSPVector member that is
// read before it is assigned (reverb.cpp:34,114-119: mvFeedbackL) still holds, in the second pass, the symbol
// assigned at the end of the first pass; consuming such a stale symbol yields a one-block feedback edge
// (FEEDBACK_READ now, FEEDBACK_WRITE of the same expression at the end of the block).  Functor state
// (phases, filter memories, delay lines) lives on the device per instance; the host functor objects hold only
// what the caller sets before tracing (coeffs, mGain, setMaxDelayInSamples, clear(), setSeed ...).
//
// Scalars: a `float` handed to a functor or mixed with a DSPVector becomes a per-instance PARAM (the
// reference's implicit float -> DSPVector broadcast, MLDSPOps.h:157); `tr::param(name, value)` names one so
// that it can be set per instance afterwards.  Coefficient design (makeCoeffs) runs on the host through
// mlb_coeffs_* = the reference's own libm calls.
//
// A functor object called more than once in a vector -- by Upsample2xFunction<N> (MLDSPFunctional.h:114-160), which runs
// its process function twice, or in an oversampled loop between an Upsampler and a Downsampler -- ticks once per call, as
// the reference's objects do: the further calls record MLB_AGAIN nodes (mlb200.h), same functor, same state.  Functors
// with a delay ring cannot be called again (the graph is refused when trace() lays it out).
//
// Header only; link against libmlb200.so.  No sample arithmetic happens on the CPU here.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "mlb200.hpp"
#include "mlb200_host.hpp"

namespace mlb
{
namespace tr
{
// kFloatsPerDSPVector is the global constant of mlb200.hpp (global in the reference too, MLDSPMath.h:8-9)

// ------------------------------------------------------------------------------------------------------
// the recorder

struct TNode
{
  mlb_node node;
  std::vector<float> coef;      // initial value of every coef word (same for all instances until overridden)
  std::vector<uint32_t> state;  // initial value of every state word
  int opIndex = -1;             // position among the user-visible operations of one pass
  std::string name;             // named PARAMs
};

class Recorder
{
 public:
  std::vector<TNode> nodes;
  std::vector<int> outs;
  int epoch = 0;
  int opCounter = 0;
  std::map<int, int> opIndexToNode;    // this pass
  std::map<int, int> feedbackReaders;  // stale opIndex -> FEEDBACK_READ node of this pass
  bool active = false;
  // a functor object called AGAIN in the same vector (allowed inside the second run of a process function by
  // Upsample2xFunction): its next node becomes a further call of its first one (MLB_AGAIN in mlb200.h)
  int repeatDepth = 0;
  int forbidFunctors = 0;       // > 0 inside a Downsample2xFunction's process function (see there)
  int pendingAgain = -1;        // node of the functor's first call in this pass
  int* pendingOwner = nullptr;  // where a functor's first call wants its node index written

  static Recorder*& current()
  {
    static thread_local Recorder* r = nullptr;
    return r;
  }
  static Recorder& get()
  {
    Recorder* r = current();
    if (!r || !r->active) throw Error(MLB_ERR_INVALID, "mlb::tr: DSPVector operation outside TracedProcessor::trace()");
    return *r;
  }
  void beginPass()
  {
    ++epoch;
    nodes.clear();
    outs.clear();
    opCounter = 0;
    opIndexToNode.clear();
    feedbackReaders.clear();
    repeatDepth = 0, forbidFunctors = 0, pendingAgain = -1, pendingOwner = nullptr;
    active = true;
  }
  int addNode(int op, std::initializer_list<int> ins, int iarg = 0, bool userOp = true)
  {
    int nin = 0, nst = 0, nco = 0;
    if (mlb_op_info(op, &nin, &nst, &nco) != MLB_OK) throw Error(MLB_ERR_INVALID, "mlb::tr: unknown op");
    TNode t;
    t.node.op = op;
    for (int k = 0; k < MLB_MAX_INS; ++k) t.node.in[k] = -1;
    int k = 0;
    for (int i : ins) t.node.in[k++] = i;
    if (k != nin && op != MLB_OP_FEEDBACK_WRITE) throw Error(MLB_ERR_INVALID, "mlb::tr: operand count");
    t.node.iarg = iarg;
    t.coef.assign((size_t)nco, 0.f);
    t.state.assign((size_t)nst, 0u);
    if (userOp && (nst > 0 || nco > 0) && (pendingOwner || pendingAgain >= 0))  // the node of the functor that just said once()
    {
      if (pendingAgain >= 0)
      {
        if (nodes[(size_t)pendingAgain].node.op != op)
          throw Error(MLB_ERR_UNSUPPORTED, "mlb::tr: a functor called again in a vector recorded a different operation");
        t.node.iarg = MLB_AGAIN(pendingAgain);
      }
      else
        *pendingOwner = (int)nodes.size();
      pendingAgain = -1, pendingOwner = nullptr;
    }
    if (userOp)
    {
      t.opIndex = opCounter++;
      opIndexToNode[t.opIndex] = (int)nodes.size();
    }
    nodes.push_back(t);
    return (int)nodes.size() - 1;
  }
};

// ------------------------------------------------------------------------------------------------------
// DSPVector: a symbolic row

class DSPVector
{
 public:
  enum Kind
  {
    kConst,
    kNode
  };
  Kind kind = kConst;
  float k = 0.f;  // kConst: the broadcast value (default-constructed = zero filled, MLDSPOps.h:153)
  int epoch = 0, node = -1, opIndex = -1;

  DSPVector() = default;
  DSPVector(float v) : kind(kConst), k(v) {}  // implicit, like the reference's DSPVectorArray(float)
  DSPVector(double v) : kind(kConst), k((float)v) {}
  DSPVector(int v) : kind(kConst), k((float)v) {}
  DSPVector(size_t v) : kind(kConst), k((float)v) {}
  static DSPVector ofNode(int n)
  {
    Recorder& r = Recorder::get();
    DSPVector d;
    d.kind = kNode, d.epoch = r.epoch, d.node = n, d.opIndex = r.nodes[(size_t)n].opIndex;
    return d;
  }
  // node index of this row in the pass being recorded (constants become PARAMs, stale symbols feedback reads)
  int resolve() const
  {
    Recorder& r = Recorder::get();
    if (kind == kConst)
    {
      const int n = r.addNode(MLB_OP_PARAM, {}, 0, false);
      r.nodes[(size_t)n].coef[0] = k;
      return n;
    }
    if (epoch == r.epoch) return node;
    // a symbol of the previous pass: state the caller keeps between process calls
    auto it = r.feedbackReaders.find(opIndex);
    if (it != r.feedbackReaders.end()) return it->second;
    const int n = r.addNode(MLB_OP_FEEDBACK_READ, {}, 0, false);
    r.feedbackReaders[opIndex] = n;
    return n;
  }
};

inline DSPVector op1(int op, const DSPVector& a) { return DSPVector::ofNode(Recorder::get().addNode(op, {a.resolve()})); }
inline DSPVector op2(int op, const DSPVector& a, const DSPVector& b)
{
  const int x = a.resolve(), y = b.resolve();
  return DSPVector::ofNode(Recorder::get().addNode(op, {x, y}));
}
inline DSPVector op3(int op, const DSPVector& a, const DSPVector& b, const DSPVector& c)
{
  const int x = a.resolve(), y = b.resolve(), z = c.resolve();
  return DSPVector::ofNode(Recorder::get().addNode(op, {x, y, z}));
}

// a named per-instance scalar, broadcast like DSPVector(float); set per instance with TracedProcessor::setParam
inline DSPVector param(const char* name, float value)
{
  Recorder& r = Recorder::get();
  const int n = r.addNode(MLB_OP_PARAM, {}, 0, false);
  r.nodes[(size_t)n].coef[0] = value;
  r.nodes[(size_t)n].name = name;
  DSPVector d;
  d.kind = DSPVector::kNode, d.epoch = r.epoch, d.node = n, d.opIndex = -1;
  return d;
}

// operators, MLDSPOps.h:337-352, 370-388 (compound forms)
inline DSPVector operator+(const DSPVector& a, const DSPVector& b) { return op2(MLB_OP_ADD, a, b); }
inline DSPVector operator-(const DSPVector& a, const DSPVector& b) { return op2(MLB_OP_SUBTRACT, a, b); }
inline DSPVector operator*(const DSPVector& a, const DSPVector& b) { return op2(MLB_OP_MULTIPLY, a, b); }
inline DSPVector operator/(const DSPVector& a, const DSPVector& b) { return op2(MLB_OP_DIVIDE, a, b); }
inline DSPVector& operator+=(DSPVector& a, const DSPVector& b) { return a = a + b; }
inline DSPVector& operator-=(DSPVector& a, const DSPVector& b) { return a = a - b; }
inline DSPVector& operator*=(DSPVector& a, const DSPVector& b) { return a = a * b; }
inline DSPVector& operator/=(DSPVector& a, const DSPVector& b) { return a = a / b; }
inline DSPVector operator-(const DSPVector& a) { return DSPVector(0.f) - a; }  // MLDSPOps.h:354-358

#define MLB_TR_OP1(NAME, OP) \
  inline DSPVector NAME(const DSPVector& x) { return op1(OP, x); }
MLB_TR_OP1(sqrt, MLB_OP_SQRT)
MLB_TR_OP1(sqrtApprox, MLB_OP_SQRT_APPROX)
MLB_TR_OP1(abs, MLB_OP_ABS)
MLB_TR_OP1(sign, MLB_OP_SIGN)
MLB_TR_OP1(signBit, MLB_OP_SIGNBIT)
MLB_TR_OP1(sin, MLB_OP_SIN)
MLB_TR_OP1(cos, MLB_OP_COS)
MLB_TR_OP1(log, MLB_OP_LOG)
MLB_TR_OP1(exp, MLB_OP_EXP)
MLB_TR_OP1(log2, MLB_OP_LOG2)
MLB_TR_OP1(exp2, MLB_OP_EXP2)
MLB_TR_OP1(sinApprox, MLB_OP_SIN_APPROX)
MLB_TR_OP1(cosApprox, MLB_OP_COS_APPROX)
MLB_TR_OP1(expApprox, MLB_OP_EXP_APPROX)
MLB_TR_OP1(logApprox, MLB_OP_LOG_APPROX)
MLB_TR_OP1(log2Approx, MLB_OP_LOG2_APPROX)
MLB_TR_OP1(exp2Approx, MLB_OP_EXP2_APPROX)
MLB_TR_OP1(fractionalPart, MLB_OP_FRACTIONAL_PART)
#undef MLB_TR_OP1
#define MLB_TR_OP2(NAME, OP) \
  inline DSPVector NAME(const DSPVector& a, const DSPVector& b) { return op2(OP, a, b); }
MLB_TR_OP2(add, MLB_OP_ADD)
MLB_TR_OP2(subtract, MLB_OP_SUBTRACT)
MLB_TR_OP2(multiply, MLB_OP_MULTIPLY)
MLB_TR_OP2(divide, MLB_OP_DIVIDE)
MLB_TR_OP2(divideApprox, MLB_OP_DIVIDE_APPROX)
MLB_TR_OP2(pow, MLB_OP_POW)
MLB_TR_OP2(powApprox, MLB_OP_POW_APPROX)
MLB_TR_OP2(min, MLB_OP_MIN)
MLB_TR_OP2(max, MLB_OP_MAX)
#undef MLB_TR_OP2
inline DSPVector lerp(const DSPVector& a, const DSPVector& b, const DSPVector& m) { return op3(MLB_OP_LERP, a, b, m); }
inline DSPVector inverseLerp(const DSPVector& a, const DSPVector& b, const DSPVector& m) { return op3(MLB_OP_INVERSE_LERP, a, b, m); }
inline DSPVector clamp(const DSPVector& x, const DSPVector& lo, const DSPVector& hi) { return op3(MLB_OP_CLAMP, x, lo, hi); }
// interpolateDSPVectorLinear(start, end), MLDSPOps.h:986-990
inline DSPVector interpolateDSPVectorLinear(const DSPVector& start, const DSPVector& end) { return op2(MLB_OP_RAMP, start, end); }

// DSPVectorInt: a row of int32 (MLDSPOps.h:370-498).  On the device a row is 64 words either way; the type only says
// how the words are read, as in the reference (whose int and float vectors share their storage layout).
class DSPVectorInt
{
 public:
  DSPVector bits;  // the symbolic row
  DSPVectorInt() : bits(intAsFloat(0)) {}
  explicit DSPVectorInt(int32_t k) : bits(intAsFloat(k)) {}
  static float intAsFloat(int32_t k)
  {
    float f;
    std::memcpy(&f, &k, 4);
    return f;
  }
  static DSPVectorInt of(const DSPVector& d)
  {
    DSPVectorInt y;
    y.bits = d;
    return y;
  }
};
// conversions, MLDSPOps.h:780-820
inline DSPVectorInt roundFloatToInt(const DSPVector& x) { return DSPVectorInt::of(op1(MLB_OP_ROUND_F2I, x)); }
inline DSPVectorInt truncateFloatToInt(const DSPVector& x) { return DSPVectorInt::of(op1(MLB_OP_TRUNC_F2I, x)); }
inline DSPVector intToFloat(const DSPVectorInt& x) { return op1(MLB_OP_INT_TO_FLOAT, x.bits); }
inline DSPVector unsignedIntToFloat(const DSPVectorInt& x) { return op1(MLB_OP_UNSIGNED_TO_FLOAT, x.bits); }
// int arithmetic, MLDSPOps.h:713-714
inline DSPVectorInt addInt32(const DSPVectorInt& a, const DSPVectorInt& b) { return DSPVectorInt::of(op2(MLB_OP_ADD_INT32, a.bits, b.bits)); }
inline DSPVectorInt subtractInt32(const DSPVectorInt& a, const DSPVectorInt& b) { return DSPVectorInt::of(op2(MLB_OP_SUBTRACT_INT32, a.bits, b.bits)); }
// comparisons give all-ones / all-zeros masks, MLDSPOps.h:851-856
#define MLB_TR_CMP(NAME, OP) \
  inline DSPVectorInt NAME(const DSPVector& a, const DSPVector& b) { return DSPVectorInt::of(op2(OP, a, b)); }
MLB_TR_CMP(equal, MLB_OP_EQUAL)
MLB_TR_CMP(notEqual, MLB_OP_NOT_EQUAL)
MLB_TR_CMP(greaterThan, MLB_OP_GREATER_THAN)
MLB_TR_CMP(greaterThanOrEqual, MLB_OP_GREATER_EQUAL)
MLB_TR_CMP(lessThan, MLB_OP_LESS_THAN)
MLB_TR_CMP(lessThanOrEqual, MLB_OP_LESS_EQUAL)
#undef MLB_TR_CMP
// bitwise select(resultIfTrue, resultIfFalse, mask), float and int forms, MLDSPOps.h:886,917
inline DSPVector select(const DSPVector& a, const DSPVector& b, const DSPVectorInt& m) { return op3(MLB_OP_SELECT, a, b, m.bits); }
inline DSPVectorInt select(const DSPVectorInt& a, const DSPVectorInt& b, const DSPVectorInt& m)
{
  return DSPVectorInt::of(op3(MLB_OP_SELECT, a.bits, b.bits, m.bits));
}
// within(x, lo, hi): is x in [lo, hi) -- a mask in a float row, MLDSPOps.h:748
inline DSPVector within(const DSPVector& x, const DSPVector& lo, const DSPVector& hi) { return op3(MLB_OP_WITHIN, x, lo, hi); }

// DSPVectorArray<ROWS>: ROWS symbolic rows (MLDSPOps.h:94-353); rowwise use only
template <size_t ROWS>
class DSPVectorArray
{
  std::array<DSPVector, ROWS> rows_;

 public:
  DSPVectorArray() = default;
  DSPVectorArray(float k) { rows_.fill(DSPVector(k)); }
  DSPVector& row(int j) { return rows_[(size_t)j]; }
  const DSPVector& constRow(int j) const { return rows_[(size_t)j]; }
};
template <size_t A, size_t B>
inline DSPVectorArray<A + B> concatRows(const DSPVectorArray<A>& a, const DSPVectorArray<B>& b)
{
  DSPVectorArray<A + B> y;
  for (size_t i = 0; i < A; ++i) y.row((int)i) = a.constRow((int)i);
  for (size_t i = 0; i < B; ++i) y.row((int)(A + i)) = b.constRow((int)i);
  return y;
}
inline DSPVectorArray<2> concatRows(const DSPVector& a, const DSPVector& b)
{
  DSPVectorArray<2> y;
  y.row(0) = a, y.row(1) = b;
  return y;
}
// ---- DSPVectorArray<ROWS> as a value: rowwise arithmetic, MLDSPOps.h:337-358 ----
#define MLB_TR_ARRAY_OP(SYM, OP)                                                                              \
  template <size_t ROWS>                                                                                      \
  inline DSPVectorArray<ROWS> operator SYM(const DSPVectorArray<ROWS>& a, const DSPVectorArray<ROWS>& b)      \
  {                                                                                                           \
    DSPVectorArray<ROWS> y;                                                                                   \
    for (size_t j = 0; j < ROWS; ++j) y.row((int)j) = op2(OP, a.constRow((int)j), b.constRow((int)j));        \
    return y;                                                                                                 \
  }                                                                                                           \
  template <size_t ROWS>                                                                                      \
  inline DSPVectorArray<ROWS>& operator SYM##=(DSPVectorArray<ROWS>& a, const DSPVectorArray<ROWS>& b)        \
  {                                                                                                           \
    return a = a SYM b;                                                                                       \
  }
MLB_TR_ARRAY_OP(+, MLB_OP_ADD)
MLB_TR_ARRAY_OP(-, MLB_OP_SUBTRACT)
MLB_TR_ARRAY_OP(*, MLB_OP_MULTIPLY)
MLB_TR_ARRAY_OP(/, MLB_OP_DIVIDE)
#undef MLB_TR_ARRAY_OP

// the "1" forms, MLDSPOps.h:678-687: the second operand is ONE row, used for every row of the first
#define MLB_TR_ARRAY_OP1(NAME, OP)                                                               \
  template <size_t ROWS>                                                                         \
  inline DSPVectorArray<ROWS> NAME(const DSPVectorArray<ROWS>& a, const DSPVector& b)            \
  {                                                                                              \
    DSPVectorArray<ROWS> y;                                                                      \
    for (size_t j = 0; j < ROWS; ++j) y.row((int)j) = op2(OP, a.constRow((int)j), b);            \
    return y;                                                                                    \
  }
MLB_TR_ARRAY_OP1(add1, MLB_OP_ADD)
MLB_TR_ARRAY_OP1(subtract1, MLB_OP_SUBTRACT)
MLB_TR_ARRAY_OP1(multiply1, MLB_OP_MULTIPLY)
MLB_TR_ARRAY_OP1(divide1, MLB_OP_DIVIDE)
MLB_TR_ARRAY_OP1(divideApprox1, MLB_OP_DIVIDE_APPROX)
MLB_TR_ARRAY_OP1(pow1, MLB_OP_POW)
MLB_TR_ARRAY_OP1(powApprox1, MLB_OP_POW_APPROX)
MLB_TR_ARRAY_OP1(min1, MLB_OP_MIN)
MLB_TR_ARRAY_OP1(max1, MLB_OP_MAX)
#undef MLB_TR_ARRAY_OP1
// lerp of two arrays with one scalar mixture, MLDSPOps.h:753-775
template <size_t ROWS>
inline DSPVectorArray<ROWS> lerp(const DSPVectorArray<ROWS>& a, const DSPVectorArray<ROWS>& b, float m)
{
  DSPVectorArray<ROWS> y;
  for (size_t j = 0; j < ROWS; ++j) y.row((int)j) = lerp(a.constRow((int)j), b.constRow((int)j), DSPVector(m));
  return y;
}

// ---- row operations, MLDSPOps.h:1056-1359: pure rearrangements of the symbolic rows (no nodes), except addRows ----
// output row j = input row src(j), or a zero row where src(j) < 0
template <size_t OUT, size_t N, class F>
inline DSPVectorArray<OUT> mapRows(const DSPVectorArray<N>& x, F src)
{
  DSPVectorArray<OUT> y;
  for (size_t j = 0; j < OUT; ++j)
  {
    const int k = src((int)j);
    y.row((int)j) = (k >= 0 && k < (int)N) ? x.constRow(k) : DSPVector(0.f);
  }
  return y;
}
template <size_t ROWS, size_t N>
inline DSPVectorArray<ROWS * N> repeatRows(const DSPVectorArray<N>& x)  // the N rows over and over
{
  return mapRows<ROWS * N>(x, [](int j) { return j % (int)N; });
}
template <size_t ROWS>
inline DSPVectorArray<ROWS> repeatRows(const DSPVector& x)  // (a DSPVector is a one-row array in the reference)
{
  DSPVectorArray<ROWS> y;
  for (size_t j = 0; j < ROWS; ++j) y.row((int)j) = x;
  return y;
}
template <size_t ROWS, size_t N>
inline DSPVectorArray<ROWS> stretchRows(const DSPVectorArray<N>& x)  // nearest input row for every output row
{
  return mapRows<ROWS>(x, [](int j) { return (int)roundf((j * ((float)N - 1.f)) / ((float)ROWS - 1.f)); });
}
template <size_t ROWS, size_t N>
inline DSPVectorArray<ROWS> zeroPadRows(const DSPVectorArray<N>& x)
{
  return mapRows<ROWS>(x, [](int j) { return j < (int)N ? j : -1; });
}
template <size_t ROWS>
inline DSPVectorArray<ROWS> shiftRows(const DSPVectorArray<ROWS>& x, int rowsToShift)  // zeros come in from outside
{
  return mapRows<ROWS>(x, [=](int j) { const int k = j - rowsToShift; return (k >= 0 && k < (int)ROWS) ? k : -1; });
}
template <size_t ROWS>
inline DSPVectorArray<ROWS> rotateRows(const DSPVectorArray<ROWS>& x, int rowsToRotate)
{
  return mapRows<ROWS>(x, [=](int j) { const int m = (int)ROWS; return (((j - rowsToRotate) % m) + m) % m; });
}
template <size_t A, size_t B>
inline DSPVectorArray<A + B> shuffleRows(const DSPVectorArray<A>& a, const DSPVectorArray<B>& b)  // a0 b0 a1 b1 ..., then the rest
{
  DSPVectorArray<A + B> y;
  size_t ja = 0, jb = 0, jy = 0;
  while (ja < A || jb < B)
  {
    if (ja < A) y.row((int)jy++) = a.constRow((int)ja++);
    if (jb < B) y.row((int)jy++) = b.constRow((int)jb++);
  }
  return y;
}
template <size_t ROWS>
inline DSPVectorArray<(ROWS + 1) / 2> evenRows(const DSPVectorArray<ROWS>& x)
{
  return mapRows<(ROWS + 1) / 2>(x, [](int j) { return 2 * j; });
}
template <size_t ROWS>
inline DSPVectorArray<ROWS / 2> oddRows(const DSPVectorArray<ROWS>& x)
{
  return mapRows<ROWS / 2>(x, [](int j) { return 2 * j + 1; });
}
template <size_t A, size_t B, size_t ROWS>
inline DSPVectorArray<B - A> separateRows(const DSPVectorArray<ROWS>& x)  // rows [A, B)
{
  static_assert(B <= ROWS && A < ROWS, "separateRows: range out of bounds");
  return mapRows<B - A>(x, [](int j) { return j + (int)A; });
}
template <size_t ROWS>
inline DSPVector addRows(const DSPVectorArray<ROWS>& x)  // from a zero row, left to right (MLDSPOps.h:1349-1359)
{
  DSPVector y(0.f);
  for (size_t j = 0; j < ROWS; ++j) y = add(y, x.constRow((int)j));
  return y;
}
template <size_t ROWS>
inline DSPVectorArray<ROWS> rowIndex()  // row j filled with j
{
  DSPVectorArray<ROWS> y;
  for (size_t j = 0; j < ROWS; ++j) y.row((int)j) = DSPVector((float)j);
  return y;
}
// columnIndex() = 0, 1, ... 63: exactly interpolateDSPVectorLinear(-1, 63) (interval 1, first value -1 + 1);
// rangeOpen / rangeClosed as the reference composes them from it (MLDSPOps.h:965-982)
inline DSPVector columnIndex() { return interpolateDSPVectorLinear(-1.f, 63.f); }
template <size_t ROWS>
inline DSPVectorArray<ROWS> columnIndex()
{
  return repeatRows<ROWS>(columnIndex());
}
inline DSPVector rangeOpen(float start, float end)
{
  const float interval = (end - start) / (float)kFloatsPerDSPVector;
  return columnIndex() * DSPVector(interval) + DSPVector(start);
}
inline DSPVector rangeClosed(float start, float end)
{
  const float interval = (end - start) / ((float)kFloatsPerDSPVector - 1.f);
  return columnIndex() * DSPVector(interval) + DSPVector(start);
}

// interpolateCoeffsLinear, MLDSPFilters.h:32-44
template <size_t N>
inline DSPVectorArray<N> interpolateCoeffsLinear(const std::array<float, N> c0, const std::array<float, N> c1)
{
  DSPVectorArray<N> vy;
  for (size_t i = 0; i < N; ++i) vy.row((int)i) = interpolateDSPVectorLinear(c0[i], c1[i]);
  return vy;
}

// ------------------------------------------------------------------------------------------------------
// functors.  Each object records ONE node per pass, the first time it is called.

class Functor
{
 protected:
  int lastEpoch_ = 0;
  int firstNode_ = -1;  // the node this object recorded when it was first called in the current pass
  void once()
  {
    Recorder& r = Recorder::get();
    if (r.forbidFunctors > 0)
      throw Error(MLB_ERR_UNSUPPORTED,
                  "mlb::tr: the process function of a Downsample2xFunction must be stateless (its functors would have to "
                  "tick on every second vector only)");
    if (lastEpoch_ == r.epoch)
    {
      // called again in the same vector: the reference object would simply tick once more (a process function run twice
      // by Upsample2xFunction, MLDSPFunctional.h:138-140; an oversampled loop between an Upsampler and a Downsampler).
      // Recorded as a further call of the first node (MLB_AGAIN); a functor with a delay ring cannot be -- trace()
      // reports that when it lays the graph out.
      if (firstNode_ < 0) throw Error(MLB_ERR_UNSUPPORTED, "mlb::tr: this object cannot be called again in the same vector");
      r.pendingAgain = firstNode_, r.pendingOwner = nullptr;
      return;
    }
    lastEpoch_ = r.epoch;
    firstNode_ = -1;
    r.pendingOwner = &firstNode_, r.pendingAgain = -1;
  }
  static TNode& nodeOf(const DSPVector& d) { return Recorder::get().nodes[(size_t)d.node]; }
};

inline float dBToGain(float dB) { return mlb_db_to_gain(dB); }  // MLDSPFilters.h:30

// ---- generators, MLDSPGens.h ----
class NoiseGen : public Functor
{
  uint32_t seed_ = 0;

 public:
  void setSeed(uint32_t s) { seed_ = s; }
  void reset() { seed_ = 0; }
  DSPVector operator()()
  {
    once();
    DSPVector y = DSPVector::ofNode(Recorder::get().addNode(MLB_OP_NOISE, {}));
    nodeOf(y).state[0] = seed_;
    return y;
  }
};
template <int OP>
class PhaseGen : public Functor
{
 protected:
  uint32_t omega32_ = 0;

 public:
  DSPVector operator()(const DSPVector& cyclesPerSample)
  {
    once();
    DSPVector y = op1(OP, cyclesPerSample);
    nodeOf(y).state[0] = omega32_;
    return y;
  }
};
class PhasorGen : public PhaseGen<MLB_OP_PHASOR>
{
 public:
  void clear(uint32_t omega = 0) { omega32_ = omega; }  // MLDSPGens.h:182
};
class SineGen : public PhaseGen<MLB_OP_SINE>
{
 public:
  void clear() { omega32_ = 0xC0000000u; }  // kZeroPhase = -(2 << 29), MLDSPGens.h:375-379
};
class SawGen : public PhaseGen<MLB_OP_SAW>
{
 public:
  void clear() { omega32_ = 0; }
};
class PulseGen : public Functor
{
  uint32_t omega32_ = 0;

 public:
  void clear() { omega32_ = 0; }
  DSPVector operator()(const DSPVector& freq, const DSPVector& width)
  {
    once();
    DSPVector y = op2(MLB_OP_PULSE, freq, width);
    nodeOf(y).state[0] = omega32_;
    return y;
  }
};
class TickGen : public Functor
{
 public:
  DSPVector operator()(const DSPVector& cyclesPerSample)
  {
    once();
    return op1(MLB_OP_TICK, cyclesPerSample);
  }
};
class ImpulseGen : public Functor
{
 public:
  DSPVector operator()(const DSPVector& cyclesPerSample)
  {
    once();
    return op1(MLB_OP_IMPULSE, cyclesPerSample);
  }
};
class OneShotGen : public Functor
{
  bool triggered_ = false;

 public:
  void trigger() { triggered_ = true; }  // MLDSPGens.h:229-233: {mOmega32 0, mGate 1, mOmegaPrev 0}
  DSPVector operator()(const DSPVector& cyclesPerSample)
  {
    once();
    DSPVector y = op1(MLB_OP_ONESHOT, cyclesPerSample);
    if (triggered_) nodeOf(y).state[1] = 1u;
    return y;
  }
};

// ---- SVF family, MLDSPFilters.h:51-442 ----
template <int OP, size_t NC>
class FixedFilter : public Functor
{
 public:
  typedef std::array<float, NC> Coeffs;
  Coeffs coeffs{};
  void clear() {}
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(OP, vx);
    for (size_t i = 0; i < NC; ++i) nodeOf(y).coef[i] = coeffs[i];
    return y;
  }
};
class Lopass : public FixedFilter<MLB_OP_LOPASS, 3>
{
 public:
  typedef DSPVectorArray<3> coeffsVec;
  static Coeffs makeCoeffs(float omega, float k)
  {
    Coeffs c;
    mlb_coeffs_lopass(omega, k, c.data());
    return c;
  }
  using FixedFilter<MLB_OP_LOPASS, 3>::operator();
  // operator()(vx, omega, k), MLDSPFilters.h:136-152: the coefficients are designed per sample.  On the device
  // that is LOPASS_MOD (CUDA sinf instead of glibc's: the stated-tolerance variant); for
  // bit-exactness design the rows on the host (mlb_coeffs_lopass_vec) and use operator()(vx, coeffsVec).
  DSPVector operator()(const DSPVector& vx, const DSPVector& omega, const DSPVector& k)
  {
    once();
    return op3(MLB_OP_LOPASS_MOD, vx, omega, k);
  }
  // coefficient ROWS g0, g1, g2 (what makeCoeffsVec returns), MLDSPFilters.h:141-149
  DSPVector operator()(const DSPVector& vx, const coeffsVec& vc)
  {
    once();
    const int x = vx.resolve(), a = vc.constRow(0).resolve(), b = vc.constRow(1).resolve(), c = vc.constRow(2).resolve();
    return DSPVector::ofNode(Recorder::get().addNode(MLB_OP_LOPASS_V, {x, a, b, c}));
  }
};
class Hipass : public FixedFilter<MLB_OP_HIPASS, 4>
{
 public:
  static Coeffs makeCoeffs(float omega, float k)
  {
    Coeffs c;
    mlb_coeffs_hipass(omega, k, c.data());
    return c;
  }
};
class Bandpass : public FixedFilter<MLB_OP_BANDPASS, 3>
{
 public:
  static Coeffs makeCoeffs(float omega, float k)
  {
    Coeffs c;
    mlb_coeffs_bandpass(omega, k, c.data());
    return c;
  }
};
class LoShelf : public FixedFilter<MLB_OP_LOSHELF, 5>
{
 public:
  typedef std::array<float, 3> params;  // omega, k, A
  typedef DSPVectorArray<5> _vcoeffs;
  static Coeffs makeCoeffs(params p)
  {
    Coeffs c;
    mlb_coeffs_loshelf(p[0], p[1], p[2], c.data());
    return c;
  }
  static _vcoeffs vcoeffs(const params p0, const params p1) { return interpolateCoeffsLinear<5>(makeCoeffs(p0), makeCoeffs(p1)); }
  using FixedFilter<MLB_OP_LOSHELF, 5>::operator();
  DSPVector operator()(const DSPVector& vx, const _vcoeffs& vc)  // MLDSPFilters.h:304-319
  {
    once();
    int in[6] = {vx.resolve(), 0, 0, 0, 0, 0};
    for (int i = 0; i < 5; ++i) in[1 + i] = vc.constRow(i).resolve();
    return DSPVector::ofNode(Recorder::get().addNode(MLB_OP_LOSHELF_V, {in[0], in[1], in[2], in[3], in[4], in[5]}));
  }
};
class HiShelf : public FixedFilter<MLB_OP_HISHELF, 6>
{
 public:
  typedef std::array<float, 3> params;
  typedef DSPVectorArray<6> _vcoeffs;
  static Coeffs makeCoeffs(params p)
  {
    Coeffs c;
    mlb_coeffs_hishelf(p[0], p[1], p[2], c.data());
    return c;
  }
  static _vcoeffs vcoeffs(const params p0, const params p1) { return interpolateCoeffsLinear<6>(makeCoeffs(p0), makeCoeffs(p1)); }
  using FixedFilter<MLB_OP_HISHELF, 6>::operator();
  DSPVector operator()(const DSPVector& vx, const _vcoeffs& vc)  // MLDSPFilters.h:385-400
  {
    once();
    int in[7] = {vx.resolve(), 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) in[1 + i] = vc.constRow(i).resolve();
    return DSPVector::ofNode(Recorder::get().addNode(MLB_OP_HISHELF_V, {in[0], in[1], in[2], in[3], in[4], in[5], in[6]}));
  }
};
class Bell : public FixedFilter<MLB_OP_BELL, 4>
{
 public:
  static Coeffs makeCoeffs(float omega, float k, float A)
  {
    Coeffs c;
    mlb_coeffs_bell(omega, k, A, c.data());
    return c;
  }
};
class OnePole : public FixedFilter<MLB_OP_ONEPOLE, 2>
{
 public:
  static Coeffs makeCoeffs(float omega)
  {
    Coeffs c;
    mlb_coeffs_onepole(omega, c.data());
    return c;
  }
  static Coeffs passthru() { return {1.f, 0.f}; }
};
class DCBlocker : public FixedFilter<MLB_OP_DCBLOCKER, 1>
{
 public:
  static Coeffs makeCoeffs(float omega) { return {mlb_coeffs_dcblocker(omega)}; }
};
class Differentiator : public FixedFilter<MLB_OP_DIFFERENTIATOR, 0>
{
};
class Integrator : public Functor
{
 public:
  float mLeak = 0.f;  // MLDSPFilters.h:544
  void setLeak(float k) { mLeak = k; }
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_INTEGRATOR, vx);
    nodeOf(y).coef[0] = mLeak;
    return y;
  }
};
class Peak : public Functor
{
 public:
  typedef std::array<float, 2> Coeffs;
  Coeffs coeffs{};
  int peakHoldSamples = 44100;  // MLDSPFilters.h:574
  static Coeffs makeCoeffs(float omega)
  {
    Coeffs c;
    mlb_coeffs_peak(omega, c.data());
    return c;
  }
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_PEAK, vx);
    nodeOf(y).coef[0] = coeffs[0], nodeOf(y).coef[1] = coeffs[1], nodeOf(y).coef[2] = (float)peakHoldSamples;
    return y;
  }
};
class RMS : public FixedFilter<MLB_OP_RMS, 2>
{
 public:
  static Coeffs makeCoeffs(float omega)
  {
    Coeffs c;
    mlb_coeffs_rms(omega, c.data());
    return c;
  }
};
class ADSR : public FixedFilter<MLB_OP_ADSR, 4>
{
 public:
  static Coeffs calcCoeffs(float a, float d, float s, float r, float sr)
  {
    Coeffs c;
    mlb_coeffs_adsr(a, d, s, r, sr, c.data());
    return c;
  }
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_ADSR, vx);
    for (size_t i = 0; i < 4; ++i) nodeOf(y).coef[i] = coeffs[i];
    nodeOf(y).state[7] = 4u;  // segment{off}, MLDSPFilters.h:694
    return y;
  }
};
class Allpass1 : public FixedFilter<MLB_OP_ALLPASS1, 1>
{
 public:
  Allpass1() = default;
  Allpass1(float a) { coeffs = {a}; }  // the reference's only constructor, MLDSPFilters.h:927
  static Coeffs makeCoeffs(float d) { return {mlb_coeffs_allpass1(d)}; }
};

// ---- smoothers, MLDSPGens.h:412-590 ----
class LinearGlide : public Functor
{
  float vectorsPerGlide_ = 32.f, dyPerVector_ = 1.f / 32.f;  // MLDSPGens.h:437-439

 public:
  void setGlideTimeInSamples(float t)
  {
    float c[2];
    mlb_coeffs_glide(t, c);
    vectorsPerGlide_ = c[0], dyPerVector_ = c[1];
  }
  DSPVector operator()(const DSPVector& f)  // the reference takes a float: sample 0 of the operand is used
  {
    once();
    DSPVector y = op1(MLB_OP_GLIDE, f);
    nodeOf(y).coef[0] = vectorsPerGlide_, nodeOf(y).coef[1] = dyPerVector_;
    nodeOf(y).state[2] = 0xFFFFFFFFu;  // mVectorsRemaining{-1}, MLDSPGens.h:440
    return y;
  }
};
class SampleAccurateLinearGlide : public Functor
{
  float samplesPerGlide_ = 32.f, dyPerSample_ = 1.f / 32.f;

 public:
  void setGlideTimeInSamples(float t)
  {
    float c[2];
    mlb_coeffs_sample_glide(t, c);
    samplesPerGlide_ = c[0], dyPerSample_ = c[1];
  }
  DSPVector operator()(const DSPVector& target)  // nextSample(target[n]) for every sample
  {
    once();
    DSPVector y = op1(MLB_OP_SAMPLE_GLIDE, target);
    nodeOf(y).coef[0] = samplesPerGlide_, nodeOf(y).coef[1] = dyPerSample_;
    nodeOf(y).state[3] = 0xFFFFFFFFu;
    return y;
  }
};
class Interpolator1 : public Functor
{
 public:
  DSPVector operator()(const DSPVector& f)
  {
    once();
    return op1(MLB_OP_INTERPOLATOR1, f);
  }
};

// ---- delays, MLDSPFilters.h:803-1155 ----
class IntegerDelay : public Functor
{
  float maxDelay_ = 0.f, delay_ = 0.f;

 public:
  IntegerDelay() = default;
  IntegerDelay(int d) { setMaxDelayInSamples((float)d), setDelayInSamples(d); }
  void setMaxDelayInSamples(float d) { maxDelay_ = d; }
  void setDelayInSamples(int d) { delay_ = (float)d; }
  void clear() {}
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_INTEGER_DELAY, vx);
    nodeOf(y).coef[0] = delay_, nodeOf(y).coef[1] = maxDelay_;
    return y;
  }
  DSPVector operator()(const DSPVector& vx, const DSPVector& vDelay)
  {
    once();
    DSPVector y = op2(MLB_OP_INTEGER_DELAY_VAR, vx, vDelay);
    nodeOf(y).coef[0] = maxDelay_;
    return y;
  }
};
class FractionalDelay : public Functor
{
  float maxDelay_ = 0.f, delay_ = 0.f;

 public:
  void setMaxDelayInSamples(float d) { maxDelay_ = d; }
  void setDelayInSamples(float d) { delay_ = d; }
  void clear() {}
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_FRACTIONAL_DELAY, vx);
    nodeOf(y).coef[0] = delay_, nodeOf(y).coef[1] = maxDelay_;
    return y;
  }
  DSPVector operator()(const DSPVector& vx, const DSPVector& vDelay)
  {
    once();
    DSPVector y = op2(MLB_OP_FRACTIONAL_DELAY_VAR, vx, vDelay);
    nodeOf(y).coef[0] = maxDelay_;
    return y;
  }
};
class PitchbendableDelay : public Functor
{
  float maxDelay_ = 0.f;

 public:
  void setMaxDelayInSamples(float d) { maxDelay_ = d; }
  float maxDelay() const { return maxDelay_; }
  void clear() {}
  DSPVector operator()(const DSPVector& vx, const DSPVector& vDelay)
  {
    once();
    DSPVector y = op2(MLB_OP_PITCHBEND_DELAY, vx, vDelay);
    nodeOf(y).coef[0] = maxDelay_;
    return y;
  }
};
// Allpass<DELAY_TYPE>, MLDSPFilters.h:1111-1155
template <typename DELAY_TYPE>
class Allpass;
template <>
class Allpass<PitchbendableDelay> : public Functor
{
  float maxDelay_ = 0.f;

 public:
  float mGain = 0.f;
  void setMaxDelayInSamples(float d) { maxDelay_ = d; }
  void clear() {}
  DSPVector operator()(const DSPVector& vx, const DSPVector& vDelay)
  {
    once();
    DSPVector y = op2(MLB_OP_ALLPASS_PB, vx, vDelay);
    nodeOf(y).coef[0] = mGain, nodeOf(y).coef[1] = maxDelay_;
    return y;
  }
};
template <>
class Allpass<IntegerDelay> : public Functor
{
  float maxDelay_ = 0.f, delay_ = 0.f;

 public:
  float mGain = 0.f;
  void setMaxDelayInSamples(float d) { maxDelay_ = d; }
  void setDelayInSamples(float d) { delay_ = d; }
  void clear() {}
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_ALLPASS_INT, vx);
    nodeOf(y).coef[0] = mGain, nodeOf(y).coef[1] = delay_, nodeOf(y).coef[2] = maxDelay_;
    return y;
  }
};
template <>
class Allpass<FractionalDelay> : public Functor
{
  float maxDelay_ = 0.f, delay_ = 0.f;

 public:
  float mGain = 0.f;
  void setMaxDelayInSamples(float d) { maxDelay_ = d; }
  void setDelayInSamples(float d) { delay_ = d; }
  void clear() {}
  DSPVector operator()(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_ALLPASS_FRAC, vx);
    nodeOf(y).coef[0] = mGain, nodeOf(y).coef[1] = delay_, nodeOf(y).coef[2] = maxDelay_;
    return y;
  }
};

// FDN<SIZE>, MLDSPFilters.h:1162-1239.  Any SIZE is recorded as what it is made of -- SIZE IntegerDelays fed by the
// vectors kept from the previous call (they become one-block feedback edges like any carried DSPVector member), the
// stereo sums, the Householder step, SIZE OnePoles, the gains -- in the reference's operation order, so with its bits
// (tests/test_oracle_port_vs_ref.py::test_fdn_of_any_size_written_out_with_its_parts).  FDN<8> (below) is ONE node
// with its own kernel.  The delays are sized for exactly their length (the reference never sizes them: SURVEY D7).
template <int SIZE>
class FDN
{
  std::array<IntegerDelay, (size_t)SIZE> delays_;
  std::array<OnePole, (size_t)SIZE> filters_;
  std::array<DSPVector, (size_t)SIZE> lines_;  // mDelayInputVectors: read before they are assigned

 public:
  std::array<float, (size_t)SIZE> mFeedbackGains{};
  void setDelaysInSamples(std::array<float, (size_t)SIZE> times)
  {
    for (int n = 0; n < SIZE; ++n)
    {
      int len = (int)(times[(size_t)n] - (float)kFloatsPerDSPVector);  // one vector of latency is in the feedback edge
      if (len < 1) len = 1;
      delays_[(size_t)n].setMaxDelayInSamples((float)len);
      delays_[(size_t)n].setDelayInSamples(len);
    }
  }
  void setFilterCutoffs(std::array<float, (size_t)SIZE> omegas)
  {
    for (int n = 0; n < SIZE; ++n) filters_[(size_t)n].coeffs = OnePole::makeCoeffs(omegas[(size_t)n]);
  }
  DSPVectorArray<2> operator()(const DSPVector& x)
  {
    for (int n = 0; n < SIZE; ++n) lines_[(size_t)n] = delays_[(size_t)n](lines_[(size_t)n]);
    DSPVector sumR, sumL;  // zero filled
    for (int n = 0; n < (SIZE & ~1); ++n)
    {
      if (n & 1)
        sumL += lines_[(size_t)n];
      else
        sumR += lines_[(size_t)n];
    }
    DSPVector sumOfDelays;
    for (int n = 0; n < SIZE; ++n) sumOfDelays += lines_[(size_t)n];
    sumOfDelays *= DSPVector(2.0f / SIZE);  // the unit-gain Householder matrix: identity minus 2 / SIZE
    for (int n = 0; n < SIZE; ++n)
    {
      DSPVector& line = lines_[(size_t)n];
      line -= sumOfDelays;
      line = filters_[(size_t)n](line) * DSPVector(mFeedbackGains[(size_t)n]);
      line += x;
    }
    return concatRows(sumL, sumR);
  }
};
template <>
class FDN<8> : public Functor
{
  static constexpr int SIZE = 8;
  std::array<float, 8> times_{}, cutoffs_{};

 public:
  std::array<float, 8> mFeedbackGains{{0, 0, 0, 0, 0, 0, 0, 0}};
  void setDelaysInSamples(std::array<float, 8> times) { times_ = times; }
  void setFilterCutoffs(std::array<float, 8> omegas) { cutoffs_ = omegas; }
  DSPVectorArray<2> operator()(const DSPVector& x)
  {
    once();
    DSPVector l = op1(MLB_OP_FDN8, x);
    mlb_coeffs_fdn8(times_.data(), cutoffs_.data(), mFeedbackGains.data(), nodeOf(l).coef.data());
    DSPVector r = DSPVector::ofNode(Recorder::get().addNode(MLB_OP_FDN8_R, {l.node}));
    return concatRows(l, r);
  }
};

// Bank<T, ROWS>, MLDSPFunctional.h:321-360: ROWS processors inside ONE traced instance.  (The batch axis --
// thousands of instances of the whole process function -- is TracedProcessor's run-time `instances`.)
template <class T, size_t ROWS>
class Bank
{
  std::array<T, ROWS> procs_;

 public:
  T& operator[](size_t n) { return procs_[n]; }
  DSPVectorArray<ROWS> operator()()
  {
    DSPVectorArray<ROWS> y;
    for (size_t j = 0; j < ROWS; ++j) y.row((int)j) = procs_[j]();
    return y;
  }
  template <typename... Args>
  DSPVectorArray<ROWS> operator()(const Args&... args)
  {
    DSPVectorArray<ROWS> y;
    for (size_t j = 0; j < ROWS; ++j) y.row((int)j) = procs_[j](args.constRow((int)j)...);
    return y;
  }
};

// HalfBandFilter, MLDSPFilters.h:1245-1310, used directly: one object works in one direction per vector --
// upsampleFirstHalf(x) then upsampleSecondHalf(x) of the same x (two rows at twice the rate), or downsample(x1, x2).
class HalfBandFilter : public Functor
{
  int upNode_ = -1, upEpoch_ = 0;

 public:
  void clear() {}
  DSPVector upsampleFirstHalf(const DSPVector& vx)
  {
    once();
    DSPVector y = op1(MLB_OP_HALFBAND_UP, vx);
    upNode_ = y.node, upEpoch_ = Recorder::get().epoch;
    return y;
  }
  DSPVector upsampleSecondHalf(const DSPVector& /*the same vx*/)
  {
    Recorder& r = Recorder::get();
    if (upEpoch_ != r.epoch || upNode_ < 0)
      throw Error(MLB_ERR_UNSUPPORTED, "mlb::tr: HalfBandFilter::upsampleSecondHalf needs upsampleFirstHalf of the same vector first");
    return DSPVector::ofNode(r.addNode(MLB_OP_HALFBAND_UP_2, {upNode_}));
  }
  DSPVector downsample(const DSPVector& vx1, const DSPVector& vx2)
  {
    once();
    return op2(MLB_OP_HALFBAND_DOWN, vx1, vx2);
  }
};

// TempoLock, MLDSPFilters.h:1478-1579: a phasor locked to the input phasor at the ratio dydx
class TempoLock : public Functor
{
 public:
  void clear() {}
  DSPVector operator()(const DSPVector& x, float dydx, float isr)
  {
    once();
    DSPVector y = op2(MLB_OP_TEMPO_LOCK, x, DSPVector(dydx));
    nodeOf(y).coef[0] = isr;
    nodeOf(y).state[0] = 0xBF800000u;  // _omega{-1.f}: stopped, MLDSPFilters.h:1481
    return y;
  }
};

// Upsample2xFunction<IN_ROWS>, MLDSPFunctional.h:114-160: the input rows upsampled by two half-band filters, the
// process function run on BOTH halves, the two results downsampled to one row.  The functors inside fn are called
// twice per vector, as in the reference (the second call records MLB_AGAIN nodes: same state, ticked again);
// functors with a delay ring cannot be (the graph is then rejected when it is compiled).
template <int IN_ROWS>
class Upsample2xFunction : public Functor
{
  using inputType = DSPVectorArray<(size_t)IN_ROWS>;
  using ProcessFn = std::function<DSPVector(const inputType)>;

 public:
  DSPVector operator()(ProcessFn fn, const inputType& vx)
  {
    once();  // the wrapper's own half-band filters are one-per-vector objects like any functor
    Recorder& r = Recorder::get();
    r.pendingOwner = nullptr, r.pendingAgain = -1;  // (its nodes are recorded here, not through a functor member)
    inputType first, second;
    for (int j = 0; j < IN_ROWS; ++j)
    {
      DSPVector up = op1(MLB_OP_HALFBAND_UP, vx.constRow(j));  // mUppers[j].upsampleFirstHalf
      first.row(j) = up;
      second.row(j) = DSPVector::ofNode(r.addNode(MLB_OP_HALFBAND_UP_2, {up.node}));  // ... upsampleSecondHalf
    }
    const DSPVector y1 = fn(first);
    ++r.repeatDepth;
    DSPVector y2;
    try
    {
      y2 = fn(second);
    }
    catch (...)
    {
      --r.repeatDepth;
      throw;
    }
    --r.repeatDepth;
    return op2(MLB_OP_HALFBAND_DOWN, y1, y2);  // mDowners[0].downsample
  }
  // the one-row form: upper(fn, x) with DSPVector in and out (dspOpsExample.cpp:100-102)
  template <int R = IN_ROWS, typename = typename std::enable_if<R == 1>::type>
  DSPVector operator()(std::function<DSPVector(const DSPVector)> fn, const DSPVector& x)
  {
    inputType in;
    in.row(0) = x;
    return (*this)([&](const inputType a) { return fn(a.constRow(0)); }, in);
  }
};

// Upsampler(octaves) / Downsampler(octaves), MLDSPFilters.h:1316-1473, used inside ONE process function: write a vector,
// read 2^octaves vectors at the higher rate, process them, write them to the Downsampler, read one vector back -- an
// oversampled loop.  Stage j of either cascade owns one half-band filter and runs it 2^j (up) / 2^(octaves-1-j) (down)
// times per vector: the first call records a HALFBAND node, the others MLB_AGAIN calls of it.  (The block-rate-changing
// use -- several process calls per write -- is the mlb_resampler_* bank.)
class Upsampler
{
  int octaves_;
  std::vector<int> firstNode_, firstEpoch_;  // per stage: the filter's first call in the current pass
  std::vector<DSPVector> rows_;
  size_t readIdx_ = 0;

 public:
  explicit Upsampler(int octavesUp) : octaves_(octavesUp), firstNode_((size_t)octavesUp, -1), firstEpoch_((size_t)octavesUp, 0) {}
  void clear() {}
  void write(const DSPVector& x)
  {
    Recorder& r = Recorder::get();
    rows_.assign(1, x);
    for (int j = 0; j < octaves_; ++j)
    {
      std::vector<DSPVector> next;
      for (const DSPVector& src : rows_)
      {
        const int in = src.resolve();
        const bool again = firstEpoch_[(size_t)j] == r.epoch;
        const int up = r.addNode(MLB_OP_HALFBAND_UP, {in}, again ? MLB_AGAIN(firstNode_[(size_t)j]) : 0);
        if (!again) firstNode_[(size_t)j] = up, firstEpoch_[(size_t)j] = r.epoch;
        next.push_back(DSPVector::ofNode(up));
        next.push_back(DSPVector::ofNode(r.addNode(MLB_OP_HALFBAND_UP_2, {up})));
      }
      rows_.swap(next);
    }
    readIdx_ = 0;
  }
  DSPVector read()  // after a write, 1 << octaves reads are available
  {
    if (readIdx_ >= rows_.size()) throw Error(MLB_ERR_INVALID, "mlb::tr: Upsampler::read past the vectors of the last write");
    return rows_[readIdx_++];
  }
};
class Downsampler
{
  int octaves_;
  std::vector<int> firstNode_, firstEpoch_;
  std::vector<DSPVector> buffers_;  // 2 per octave + the output, as in the reference
  uint32_t counter_ = 0;
  int epoch_ = 0;

 public:
  explicit Downsampler(int octavesDown)
      : octaves_(octavesDown), firstNode_((size_t)octavesDown, -1), firstEpoch_((size_t)octavesDown, 0),
        buffers_((size_t)(2 * octavesDown + 1))
  {
  }
  void clear() {}
  bool write(const DSPVector& v)  // true when a new output vector is ready (every 2^octaves writes)
  {
    Recorder& r = Recorder::get();
    if (epoch_ != r.epoch)
    {
      if (counter_ != 0)
        throw Error(MLB_ERR_UNSUPPORTED, "mlb::tr: a Downsampler must be written 2^octaves times in every vector");
      epoch_ = r.epoch;
    }
    if (octaves_ == 0)
    {
      buffers_[0] = v;
      return true;
    }
    buffers_[counter_ & 1u] = v;
    uint32_t mask = 1;
    for (int h = 0; h < octaves_; ++h)
    {
      if (!(counter_ & mask)) break;  // an octave runs when its bit and all lesser bits are 1
      mask <<= 1;
      const bool b1 = (counter_ & mask) != 0;
      const int x1 = buffers_[(size_t)(2 * h)].resolve(), x2 = buffers_[(size_t)(2 * h + 1)].resolve();
      const bool again = firstEpoch_[(size_t)h] == r.epoch;
      const int down = r.addNode(MLB_OP_HALFBAND_DOWN, {x1, x2}, again ? MLB_AGAIN(firstNode_[(size_t)h]) : 0);
      if (!again) firstNode_[(size_t)h] = down, firstEpoch_[(size_t)h] = r.epoch;
      buffers_[(size_t)(2 * h + 2 + (b1 ? 1 : 0))] = DSPVector::ofNode(down);
    }
    counter_ = (counter_ + 1u) & ((1u << octaves_) - 1u);
    return counter_ == 0;
  }
  DSPVector read() const { return buffers_.back(); }
};

// Downsample2xFunction<IN_ROWS>, MLDSPFunctional.h:166-223: on every second vector the input (this vector and the
// buffered previous one) is downsampled, the process function runs on that half-rate vector and its result is upsampled
// into this vector and the next.  Recorded as DOWN2X_IN per input row, fn's nodes, DOWN2X_OUT (mlb200.h).  fn must be
// stateless: its nodes run on every vector here, its functors would have to tick on every second one only.
template <int IN_ROWS>
class Downsample2xFunction : public Functor
{
  using inputType = DSPVectorArray<(size_t)IN_ROWS>;
  using ProcessFn = std::function<DSPVector(const inputType)>;

 public:
  DSPVector operator()(ProcessFn fn, const inputType& vx)
  {
    once();
    Recorder& r = Recorder::get();
    r.pendingOwner = nullptr, r.pendingAgain = -1;
    inputType half;
    for (int j = 0; j < IN_ROWS; ++j) half.row(j) = op1(MLB_OP_DOWN2X_IN, vx.constRow(j));
    ++r.forbidFunctors;
    DSPVector y;
    try
    {
      y = fn(half);
    }
    catch (...)
    {
      --r.forbidFunctors;
      throw;
    }
    --r.forbidFunctors;
    return op1(MLB_OP_DOWN2X_OUT, y);
  }
  template <int R = IN_ROWS, typename = typename std::enable_if<R == 1>::type>
  DSPVector operator()(std::function<DSPVector(const DSPVector)> fn, const DSPVector& x)
  {
    inputType in;
    in.row(0) = x;
    return (*this)([&](const inputType a) { return fn(a.constRow(0)); }, in);
  }
};

// ---- scalar host helpers the examples use (MLDSPProjections.h:15-23,105-123,176-195) ----
struct Interval
{
  float x1, x2;
};
using Projection = std::function<float(float)>;
namespace projections
{
inline Projection log(Interval m)
{
  const float a = m.x1, b = m.x2;
  if (b - a == 0.f) return [=](float) { return a; };
  if (a == 0.f) return [=](float) { return 0.f; };
  return [=](float x) { return a * (powf((b / a), x) - 1) / (b - a); };
}
inline Projection intervalMap(const Interval a, const Interval b, Projection c)
{
  return [=](float x)
  {
    const float scaleA = 1 / (a.x2 - a.x1);
    const float offsetA = (-a.x1) / (a.x2 - a.x1);
    const float scaleB = (b.x2 - b.x1);
    const float offsetB = b.x1;
    return c(x * scaleA + offsetA) * scaleB + offsetB;
  };
}
inline Projection unityToLogParam(Interval paramInterval) { return intervalMap({0, 1}, paramInterval, log(paramInterval)); }
}  // namespace projections

// ------------------------------------------------------------------------------------------------------
// AudioContext / SignalProcessFn (source/app/MLAudioContext.h:23-101, MLSignalProcessBuffer.h:18)

class DSPVectorDynamic
{
  std::vector<DSPVector> rows_;

 public:
  explicit DSPVectorDynamic(size_t n = 0) : rows_(n) {}
  size_t size() const { return rows_.size(); }
  DSPVector& operator[](size_t i) { return rows_[i]; }
  const DSPVector& operator[](size_t i) const { return rows_[i]; }
};
using MainInputs = const DSPVectorDynamic&;
using MainOutputs = DSPVectorDynamic&;

class AudioContext
{
  double sampleRate_ = 0;

 public:
  AudioContext(size_t nInputs, size_t nOutputs, int rate = 0) : sampleRate_(rate), inputs(nInputs), outputs(nOutputs) {}
  double getSampleRate() const { return sampleRate_; }
  void setSampleRate(int r) { sampleRate_ = r; }
  DSPVectorDynamic inputs;
  DSPVectorDynamic outputs;
};
using SignalProcessFn = void (*)(AudioContext*, void*);

// ------------------------------------------------------------------------------------------------------
// TracedProcessor: trace a SignalProcessFn, compile it for `instances` independent copies, run it on the GPU.

class TracedProcessor
{
  Recorder rec_;
  std::vector<mlb_node> nodes_;
  std::vector<int32_t> outs_;
  std::vector<int32_t> stOff_, coOff_;
  mlb_layout layout_{};
  std::vector<float> coef_;
  std::vector<uint32_t> state_;
  std::map<std::string, int> named_;
  mlb_graph* g_ = nullptr;
  int instances_ = 0;
  size_t nIn_ = 0, nOut_ = 0;
  BatchedSignalProcessBuffer* buffer_ = nullptr;

  void runPass(AudioContext* ctx, SignalProcessFn fn, void* state)
  {
    rec_.beginPass();
    for (size_t i = 0; i < ctx->inputs.size(); ++i)
      ctx->inputs[i] = DSPVector::ofNode(rec_.addNode(MLB_OP_INPUT, {}, (int)i));
    fn(ctx, state);
  }

 public:
  TracedProcessor() = default;
  ~TracedProcessor()
  {
    if (g_) mlb_graph_destroy(g_);
    delete buffer_;
  }
  TracedProcessor(const TracedProcessor&) = delete;
  TracedProcessor& operator=(const TracedProcessor&) = delete;

  // Record the graph `fn` builds (host only: works without a GPU).
  void trace(AudioContext* ctx, SignalProcessFn fn, void* state)
  {
    Recorder*& cur = Recorder::current();
    Recorder* prev = cur;
    cur = &rec_;
    try
    {
      runPass(ctx, fn, state);  // pass 1: discovers which symbols survive the call
      runPass(ctx, fn, state);  // pass 2: the graph; stale symbols of pass 1 become feedback edges
      // the expression a stale symbol stood for is, in this pass, the node with the same operation index
      for (const auto& fr : rec_.feedbackReaders)
      {
        auto it = rec_.opIndexToNode.find(fr.first);
        if (it == rec_.opIndexToNode.end()) throw Error(MLB_ERR_INVALID, "mlb::tr: process function is not deterministic between calls");
        const int w = rec_.addNode(MLB_OP_FEEDBACK_WRITE, {it->second}, fr.second, false);
        (void)w;
      }
      nOut_ = ctx->outputs.size();
      for (size_t i = 0; i < nOut_; ++i) rec_.outs.push_back(ctx->outputs[i].resolve());
    }
    catch (...)
    {
      rec_.active = false;
      cur = prev;
      throw;
    }
    rec_.active = false;
    cur = prev;
    nIn_ = ctx->inputs.size();
    nodes_.clear();
    for (const TNode& t : rec_.nodes) nodes_.push_back(t.node);
    outs_.assign(rec_.outs.begin(), rec_.outs.end());
    stOff_.assign(nodes_.size(), 0);
    coOff_.assign(nodes_.size(), 0);
    check(mlb_graph_layout(nodes_.data(), (int)nodes_.size(), &layout_, stOff_.data(), coOff_.data()));
    named_.clear();
    for (size_t i = 0; i < rec_.nodes.size(); ++i)
      if (!rec_.nodes[i].name.empty()) named_[rec_.nodes[i].name] = (int)i;
  }

  size_t nodeCount() const { return nodes_.size(); }
  const std::vector<mlb_node>& nodes() const { return nodes_; }
  const std::vector<int32_t>& outs() const { return outs_; }
  size_t inputs() const { return nIn_; }
  size_t outputs() const { return nOut_; }

  static bool isAgain(const mlb_node& n)
  {
    return n.iarg < 0 && n.op != MLB_OP_INPUT && n.op != MLB_OP_PARAM && n.op != MLB_OP_FEEDBACK_WRITE;
  }
  // Write the traced graph as JSON (nodes, outs, initial coefficient and state words): lets a test or a tool
  // rebuild it elsewhere (tests/test_trace.py feeds it to the CPU checkers).
  void dump(FILE* f) const
  {
    std::fprintf(f, "{\"n_in\": %zu, \"nodes\": [", nIn_);
    for (size_t i = 0; i < rec_.nodes.size(); ++i)
    {
      const TNode& t = rec_.nodes[i];
      std::fprintf(f, "%s{\"op\": %d, \"in\": [", i ? ", " : "", t.node.op);
      for (int k = 0; k < MLB_MAX_INS; ++k) std::fprintf(f, "%s%d", k ? ", " : "", t.node.in[k]);
      std::fprintf(f, "], \"iarg\": %d, \"coef\": [", t.node.iarg);
      const bool again = isAgain(t.node);  // a further call of an earlier functor owns no words
      for (size_t k = 0; k < t.coef.size() && !again; ++k)
      {
        uint32_t u;
        std::memcpy(&u, &t.coef[k], 4);
        std::fprintf(f, "%s%u", k ? ", " : "", u);  // bit patterns: exact round trip
      }
      std::fprintf(f, "], \"state\": [");
      for (size_t k = 0; k < t.state.size() && !again; ++k) std::fprintf(f, "%s%u", k ? ", " : "", t.state[k]);
      std::fprintf(f, "], \"name\": \"%s\"}", t.name.c_str());
    }
    std::fprintf(f, "], \"outs\": [");
    for (size_t i = 0; i < outs_.size(); ++i) std::fprintf(f, "%s%d", i ? ", " : "", outs_[i]);
    std::fprintf(f, "]}\n");
  }

  // Create the device bank: `instances` independent copies of the traced process function (needs a GPU).
  void compile(int instances, unsigned flags = MLB_GRAPH_EXACT)
  {
    if (nodes_.empty()) throw Error(MLB_ERR_INVALID, "mlb::tr: trace() first");
    if (g_) mlb_graph_destroy(g_), g_ = nullptr;
    instances_ = instances;
    check(mlb_graph_create(nodes_.data(), (int)nodes_.size(), outs_.data(), (int)outs_.size(), instances, flags, &g_));
    coef_.assign((size_t)layout_.n_coef_words * instances, 0.f);
    state_.assign((size_t)layout_.n_state_words * instances, 0u);
    for (size_t i = 0; i < rec_.nodes.size(); ++i)
    {
      const TNode& t = rec_.nodes[i];
      if (isAgain(t.node)) continue;  // the words belong to the functor's first call
      for (size_t k = 0; k < t.coef.size(); ++k)
        for (int v = 0; v < instances; ++v) coef_[(size_t)(coOff_[i] + (int)k) * instances + v] = t.coef[k];
      for (size_t k = 0; k < t.state.size(); ++k)
        for (int v = 0; v < instances; ++v) state_[(size_t)(stOff_[i] + (int)k) * instances + v] = t.state[k];
    }
    commit();
  }
  void commit()
  {
    check(mlb_graph_set_coefs(g_, coef_.data()));
    check(mlb_graph_set_state(g_, state_.data()));
  }
  int instances() const { return instances_; }
  const char* kernelName() const { return g_ ? mlb_graph_kernel_name(g_) : ""; }
  // per-instance value of a tr::param(name, ...) scalar; call commit() after a batch of changes
  void setParam(const std::string& name, int instance, float value)
  {
    auto it = named_.find(name);
    if (it == named_.end()) throw Error(MLB_ERR_INVALID, "mlb::tr: no such param");
    coef_[(size_t)coOff_[(size_t)it->second] * instances_ + instance] = value;
  }
  // per-instance coefficient / state word of node `node` (indices as in nodes())
  void setCoef(int node, int word, int instance, float value) { coef_[(size_t)(coOff_[(size_t)node] + word) * instances_ + instance] = value; }
  void setStateWord(int node, int word, int instance, uint32_t value) { state_[(size_t)(stOff_[(size_t)node] + word) * instances_ + instance] = value; }

  // n_blocks successive process calls for every instance in ONE kernel launch.
  // in [T][inputs][instances][64], out [T][outputs][instances][64], mix [T][outputs][64] (sum over instances).
  void processBlocks(const float* in, float* out, float* mix, int nBlocks) { check(mlb_graph_process_host(g_, in, out, mix, nBlocks)); }

  // SignalProcessBuffer::process for ONE instance's I/O (source/app/MLSignalProcessBuffer.cpp:36-90): any host
  // buffer size; every vector the callback needs is computed by one launch.  With instances > 1 the outputs are
  // the mix bus (sum over instances, Synth::processVector's accumulate, source/app/MLSynth.h:36-60) and every
  // instance receives the same external inputs.
  void process(const float** externalInputs, float** externalOutputs, int nFrames, int maxFrames = 4096)
  {
    if (!g_) throw Error(MLB_ERR_INVALID, "mlb::tr: compile() first");
    if (!buffer_) buffer_ = new BatchedSignalProcessBuffer(nIn_, nOut_, (size_t)maxFrames);
    std::vector<float>& bin = bcast_;
    buffer_->process(externalInputs, externalOutputs, nFrames,
                     [&](const float* in, float* out, int nVec)
                     {
                       const float* src = in;
                       if (nIn_ && instances_ > 1)
                       {
                         bin.resize((size_t)nVec * nIn_ * instances_ * MLB_BLOCK);
                         for (size_t r = 0; r < (size_t)nVec * nIn_; ++r)
                           for (int v = 0; v < instances_; ++v)
                             std::memcpy(&bin[(r * instances_ + v) * MLB_BLOCK], in + r * MLB_BLOCK, MLB_BLOCK * 4);
                         src = bin.data();
                       }
                       if (instances_ > 1)
                         check(mlb_graph_process_host(g_, nIn_ ? src : nullptr, nullptr, out, nVec));
                       else
                         check(mlb_graph_process_host(g_, nIn_ ? src : nullptr, out, nullptr, nVec));
                     });
  }

 private:
  std::vector<float> bcast_;
};

}  // namespace tr
}  // namespace mlb
