// tests/cpp/test_hostapi.cpp -- the reference's own hot-path assertions, restated against the
// C++ host mirror (include/mlb200.hpp) of the B200 engine.  Mirrors:
//   Tests/dspGensTest.cpp:15-31   (SineGen one cycle ends at 0)
//   Tests/dspOpsTest.cpp:77-106   (precision of sin/cos/log/exp, precise < 2e-6, approx < 2e-4)
//   Tests/dspOpsTest.cpp:148-165  (lerp, fractionalPart)
//   Tests/dspOpsTest.cpp:273-293  (Bank<SineGen, 5> runs)
// plus the SURVEY 8c pinned value of 0.5 * Lopass{0.1,1.0}(SineGen.clear()(440/48000)).
// Exit code 0 = all assertions hold; 77 = no GPU (the library refuses to compute on the CPU).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "mlb200.hpp"
#include "mlb200_host.hpp"

using namespace mlb;

static int g_checks = 0, g_fail = 0;
#define REQUIRE(cond)                                                   \
  do                                                                    \
  {                                                                     \
    ++g_checks;                                                         \
    if (!(cond))                                                        \
    {                                                                   \
      ++g_fail;                                                         \
      std::printf("REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
    }                                                                   \
  } while (0)

static float maxAbsDiff(const DSPVector& a, const DSPVector& b, const DSPVector* domain = nullptr)
{
  float m = 0.f;
  for (size_t i = 0; i < kFloatsPerDSPVector; ++i)
  {
    if (domain && !((*domain)[i] > 0.f)) continue;
    const float d = std::fabs(a[i] - b[i]);
    if (d > m) m = d;
  }
  return m;
}

// host-only parts of the value types (no device needed): the map helpers of MLDSPFunctional.h:23-100,
// as tests/DSPOpsTest.cpp uses them (map with a lambda == the loop it stands for)
static int host_only_checks()
{
  int bad = 0;
  DSPVectorArray<3> x;
  for (size_t n = 0; n < 3 * kFloatsPerDSPVector; ++n) x[n] = 0.25f * (float)n - 7.f;
  int calls = 0;
  const auto counted = map<3>([&]() { return (float)calls++; }, x);
  bad += !(calls == 3 * (int)kFloatsPerDSPVector && counted[5] == 5.f && counted[191] == 191.f);
  const auto sq = map<3>([](float v) { return v * v; }, x);
  for (size_t n = 0; n < 3 * kFloatsPerDSPVector; ++n) bad += !(sq[n] == x[n] * x[n]);
  DSPVectorArrayInt<2> xi;
  for (int n = 0; n < 2 * (int)kFloatsPerDSPVector; ++n) xi[n] = n - 3;
  const auto half = map<2>([](int v) { return 0.5f * (float)v; }, xi);
  bad += !(half[0] == -1.5f && half[127] == 62.f);
  // row forms: a row function that needs no device (assignment and indexing only)
  const auto rev = map<3>(std::function<DSPVector(const DSPVector)>([](const DSPVector r) {
                            DSPVector y;
                            for (size_t i = 0; i < kFloatsPerDSPVector; ++i) y[i] = r[kFloatsPerDSPVector - 1 - i];
                            return y;
                          }),
                          x);
  bad += !(rev[0] == x[63] && rev[64 + 10] == x[64 + 53] && rev[191] == x[128]);
  const auto tagged = map<3>(std::function<DSPVector(const DSPVector, int)>([](const DSPVector r, int j) {
                               DSPVector y(r);
                               y[0] = (float)(100 + j);
                               return y;
                             }),
                             x);
  bad += !(tagged[0] == 100.f && tagged[64] == 101.f && tagged[128] == 102.f && tagged[129] == x[129]);
  const auto idx = map<3>(std::function<DSPVector(const DSPVector, const DSPVector)>(
                              [](const DSPVector, const DSPVector j) { return j; }),
                          x);
  bad += !(idx[7] == 0.f && idx[64 + 7] == 1.f && idx[191] == 2.f);
  std::printf("host-only checks (map helpers): %s\n", bad ? "FAILED" : "ok");
  return bad;
}

int main()
{
  if (host_only_checks()) return 1;
  if (mlb_device_count() < 1)
  {
    // no silent fallback: creating a bank must fail loudly
    Graph g;
    g.output(g.sine(g.input(0)));
    bool threw = false;
    try { DeviceBank b(g, 4); } catch (const Error& e) { threw = (e.code == MLB_ERR_NO_DEVICE); }
    std::printf("no GPU visible: DeviceBank creation %s\n", threw ? "failed loudly (ok)" : "DID NOT FAIL");
    return threw ? 77 : 1;
  }

  // ---- dsp_gens: one cycle of sine should end at 0 ----
  {
    Graph g;
    const int f = g.input(0), s = g.sine(f);
    g.output(s);
    DeviceBank bank(g, 1);
    bank.setStateWord(s, 0, 0, 0xC0000000u);  // SineGen::clear()
    bank.commit();
    DSPVector v1 = bank(DSPVector(1.f / kFloatsPerDSPVector));
    const float epsilon = std::pow(10.f, -120.f / 20.f);  // dBToAmp(-120)
    REQUIRE(std::fabs(v1[kFloatsPerDSPVector - 1]) < epsilon);
  }

  // ---- pinned reference value: 0.5 * Lopass{makeCoeffs(0.1, 1.0)}(SineGen.clear()(440/48000)) ----
  {
    Graph g;
    const int f = g.input(0), s = g.sine(f), lp = g.lopass(s), k = g.param(), y = g.multiply(lp, k);
    g.output(y);
    DeviceBank bank(g, 1);
    float c[3];
    mlb_coeffs_lopass(0.1f, 1.0f, c);
    bank.setCoeffs(lp, 0, c, 3);
    bank.setParam(k, 0, 0.5f);
    bank.setStateWord(s, 0, 0, 0xC0000000u);
    bank.commit();
    REQUIRE(std::strncmp(bank.kernelName(), "fused:", 6) == 0);
    DSPVector out = bank(DSPVector(440.f / 48000.f));
    REQUIRE(out[0] == -0x1.09e64p-9f);
    REQUIRE(out[1] == -0x1.5cd8fap-7f);
    REQUIRE(out[63] == 0x1.b2539p-3f);
  }

  // ---- dsp_ops precision: precise < 2e-6, approx < 2e-4 vs libm on rangeClosed(-pi, pi) ----
  {
    const float kPi = 3.1415926535897932384626433f;
    DSPVector a = rangeClosed(-kPi, kPi);
    DSPVector nat;
    for (size_t i = 0; i < kFloatsPerDSPVector; ++i) nat[i] = std::sin(a[i]);
    REQUIRE(maxAbsDiff(nat, sin(a)) < 2e-6f);
    REQUIRE(maxAbsDiff(nat, sinApprox(a)) < 2e-4f);
    for (size_t i = 0; i < kFloatsPerDSPVector; ++i) nat[i] = std::cos(a[i]);
    REQUIRE(maxAbsDiff(nat, cos(a)) < 2e-6f);
    REQUIRE(maxAbsDiff(nat, cosApprox(a)) < 2e-4f);
    for (size_t i = 0; i < kFloatsPerDSPVector; ++i) nat[i] = std::exp(a[i]);
    REQUIRE(maxAbsDiff(nat, exp(a)) < 2e-6f);
    REQUIRE(maxAbsDiff(nat, expApprox(a)) < 2e-4f);
    for (size_t i = 0; i < kFloatsPerDSPVector; ++i) nat[i] = a[i] > 0 ? std::log(a[i]) : 0.f;
    REQUIRE(maxAbsDiff(nat, log(a), &a) < 2e-6f);  // x <= 0 drops out of the reference's max() too
    REQUIRE(maxAbsDiff(nat, logApprox(a), &a) < 2e-4f);
  }

  // ---- lerp / convert ----
  {
    DSPVector a;
    for (size_t i = 0; i < kFloatsPerDSPVector; ++i) a[i] = (float)i;
    DSPVector b(0.f);
    DSPVector c = lerp(a, b, DSPVector(0.5f));
    REQUIRE(c[kFloatsPerDSPVector - 1] == (kFloatsPerDSPVector - 1) * 0.5f);
    DSPVector fa = fractionalPart(DSPVector(1.25f)), fb = fractionalPart(DSPVector(-1.25f));
    REQUIRE(fa[kFloatsPerDSPVector - 1] == -fb[kFloatsPerDSPVector - 1]);
    // operators with implicit float -> DSPVector conversion (MLDSPOps.h:157)
    DSPVector d = a * 2.f + 1.f;
    REQUIRE(d[10] == 21.f);
    DSPVectorArray<2> e(3.f);
    REQUIRE((e * e)[100] == 9.f);
  }

  // ---- row operations, the int type, compound assignment, "*1" forms (MLDSPOps.h:312-331,370-498,655-687,
  // 1057-1383; the reference's own section, Tests/dspOpsTest.cpp:187-230, builds these without assertions) ----
  {
    DSPVectorArray<2> a{repeatRows<2>(columnIndex())};
    auto a2{a * 2.f};
    REQUIRE(a2[5] == 10.f && a2[64 + 63] == 126.f);
    DSPVector b{columnIndex()};
    auto e = a * repeatRows<2>(b);
    REQUIRE(e[64 + 7] == 49.f);
    REQUIRE(multiply1(a, b) == e);  // one row applied to every row
    auto aa = repeatRows<4>(a);
    REQUIRE(aa.constRow(7) == b);
    DSPVectorArray<2> g = a;
    g.row(1) = b * 2.f;
    auto h = stretchRows<6>(g);
    REQUIRE(h.constRow(0) == b && h.constRow(2) == b && h.constRow(3) == g.constRow(1) && h.constRow(5) == g.constRow(1));
    auto k = zeroPadRows<6>(columnIndex());
    REQUIRE(k.constRow(0) == b && k.constRow(1) == DSPVector(0.f));
    auto m = rotateRows(k, -1) * 3.f;   // row 0 moves to row 5
    REQUIRE(m.constRow(5) == b * 3.f && m.constRow(0) == DSPVector(0.f));
    auto n = shiftRows(k, 2);
    REQUIRE(n.constRow(2) == b && n.constRow(0) == DSPVector(0.f));
    DSPVectorArray<3> gains = concatRows(DSPVector{0.300f}, DSPVector{0.030f}, DSPVector{0.003f});
    DSPVectorArray<6> gg = repeatRows<2>(gains);
    DSPVectorArray<2> hh = separateRows<4, 6>(gg);
    REQUIRE(hh.constRow(0) == DSPVector(0.030f) && hh.constRow(1) == DSPVector(0.003f));
    REQUIRE(evenRows(gg).constRow(1) == DSPVector(0.003f) && oddRows(gg).constRow(0) == DSPVector(0.030f));
    REQUIRE(shuffleRows(DSPVector(1.f), concatRows(DSPVector(2.f), DSPVector(3.f))).constRow(2) == DSPVector(3.f));
    REQUIRE(addRows(gg)[17] == ((((0.f + 0.3f) + 0.03f) + 0.003f) + 0.3f) + 0.03f + 0.003f);
    DSPVectorArray<2> acc(1.f);
    acc += a;
    acc *= DSPVectorArray<2>(2.f);
    acc -= DSPVectorArray<2>(1.f);
    acc /= DSPVectorArray<2>(0.5f);
    REQUIRE(acc[3] == ((1.f + 3.f) * 2.f - 1.f) / 0.5f);
    // ints: round / truncate, comparisons -> masks, bitwise select, int add
    DSPVector x = columnIndex() * 0.5f - 4.f;
    DSPVectorInt r = roundFloatToInt(x), t = truncateFloatToInt(x);
    REQUIRE(r[1] == -4 && r[3] == -2 && t[1] == -3 && t[9] == 0);  // RN-even: -3.5 -> -4, -2.5 -> -2
    DSPVectorInt mask = greaterThan(x, DSPVector(0.f));
    REQUIRE(mask[8] == 0 && mask[9] == -1);
    DSPVector sel = select(DSPVector(1.f), DSPVector(-1.f), mask);
    REQUIRE(sel[0] == -1.f && sel[63] == 1.f);
    REQUIRE((r + t)[1] == -7 && subtractInt32(r, t)[1] == -1);
    REQUIRE(intToFloat(r)[3] == -2.f);
    REQUIRE(rangeOpen(0.f, 1.f)[32] == 0.5f && interpolateDSPVectorLinear(0.f, 1.f)[63] == 1.f);
    // the staging pool behind these operators stopped allocating long ago
    const long long allocs = mlb_map_host_allocations();
    for (int i = 0; i < 50; ++i) a2 = a2 + a;
    REQUIRE(mlb_map_host_allocations() == allocs);
  }

  // ---- Bank<SineGen, 5>-shaped bank: rows are independent voices ----
  {
    Graph g;
    const int f = g.input(0), s = g.sine(f);
    g.output(s);
    DeviceBank bank(g, 5);
    for (int r = 0; r < 5; ++r) bank.setStateWord(s, 0, r, 0xC0000000u);
    bank.commit();
    DSPVectorArray<5> freqs;
    for (int r = 0; r < 5; ++r) freqs.row(r) = DSPVector(0.01f * (r + 1));
    DSPVectorArray<5> y = bank(freqs);
    DeviceBank one(g, 1);
    one.setStateWord(s, 0, 0, 0xC0000000u);
    one.commit();
    DSPVector y3 = one(DSPVector(0.04f));
    REQUIRE(y.constRow(3) == y3);  // row 3 of the bank == a single voice at the same frequency
    REQUIRE(!(y.constRow(0) == y.constRow(1)));
    bank.readState();
    REQUIRE(bank.stateWord(s, 0, 2) != 0xC0000000u);  // phase advanced
  }

  // ---- the caller of the boundary: arbitrary host callback sizes, ONE launch per callback ----
  // (reference: SignalProcessBuffer::process, source/app/MLSignalProcessBuffer.cpp:36-90, drives the
  //  user's SignalProcessFn once per 64 frames; here the whole callback is one batched call whose
  //  output is the mix bus of a 64-voice bank, like Synth::processVector summing voices)
  {
    const int V = 64;
    Graph g;
    const int f = g.param(), s = g.sine(f), lp = g.lopass(s), k = g.param(), y = g.multiply(lp, k);
    g.output(y);
    auto setup = [&](DeviceBank& bank) {
      for (int v = 0; v < V; ++v)
      {
        float c[3];
        mlb_coeffs_lopass(0.02f + 0.003f * v, 0.5f, c);
        bank.setParam(f, v, (110.f + 3.f * v) / 48000.f);
        bank.setCoeffs(lp, v, c, 3);
        bank.setParam(k, v, 0.01f);
        bank.setStateWord(s, 0, v, 0xC0000000u);
      }
      bank.commit();
    };
    // direct: 40 blocks of mix bus in one call
    DeviceBank direct(g, V);
    setup(direct);
    std::vector<float> want(40 * 64);
    direct.process(nullptr, nullptr, want.data(), 40);
    // through the batched process buffer with awkward callback sizes
    DeviceBank bank(g, V);
    setup(bank);
    BatchedSignalProcessBuffer spb(0, 1, 4096);  // kMaxProcessBlockFrames, source/app/MLAudioTask.h:25
    const long long launches0 = mlb_kernel_launches();
    int callbacks = 0, vectors = 0;
    std::vector<float> got;
    const int sizes[] = {480, 17, 1000, 3, 64, 513, 129, 255};
    for (int frames : sizes)
    {
      std::vector<float> buf(frames);
      float* outs[1] = {buf.data()};
      vectors += spb.process(nullptr, outs, frames, [&](const float*, float* out, int n) {
        bank.process(nullptr, nullptr, out, n);  // out = mix bus [n][1][64]
      });
      ++callbacks;
      got.insert(got.end(), buf.begin(), buf.end());
    }
    REQUIRE(got.size() <= want.size());
    bool same = true;
    for (size_t i = 0; i < got.size(); ++i) same = same && (got[i] == want[i]);
    REQUIRE(same);  // the host sees one continuous stream, whatever the callback sizes
    REQUIRE(vectors == (int)((got.size() + 63) / 64));
    REQUIRE(mlb_kernel_launches() - launches0 <= 2 * callbacks);  // chain + mix-reduce per callback at most
  }

  // ---- functor nodes through the C++ mirror: a 64-sample comb through a feedback edge ----
  {
    // y[t] = x[t] + 0.5 * y[t - 1 block]: an impulse comes back halved every 64 samples
    Graph g;
    const int x = g.input(0), k = g.param(), fb = g.feedbackRead();
    const int y = g.add2(x, g.multiply(fb, k));
    g.feedbackWrite(fb, y);
    g.output(y);
    DeviceBank bank(g, 2);
    bank.setParam(k, 0, 0.5f);
    bank.setParam(k, 1, 0.25f);
    bank.commit();
    std::vector<float> in(4 * 2 * 64, 0.f), out(4 * 2 * 64, 0.f);
    in[0] = 1.f;       // voice 0, block 0, sample 0
    in[64 + 3] = 1.f;  // voice 1, block 0, sample 3
    bank.process(in.data(), out.data(), nullptr, 4);
    REQUIRE(out[0] == 1.f && out[2 * 64] == 0.5f && out[4 * 64] == 0.25f && out[6 * 64] == 0.125f);
    REQUIRE(out[64 + 3] == 1.f && out[3 * 64 + 3] == 0.25f && out[5 * 64 + 3] == 0.0625f);
  }

  // ---- Upsampler / Downsampler banks: 2x up then 2x down gives the input back, delayed and low-passed ----
  {
    const int V = 3, T = 24;
    std::vector<float> x((size_t)T * V * 64), up((size_t)2 * T * V * 64), back((size_t)T * V * 64);
    for (int t = 0; t < T; ++t)
      for (int v = 0; v < V; ++v)
        for (int n = 0; n < 64; ++n) x[((size_t)t * V + v) * 64 + n] = std::sin(0.05f * (float)(t * 64 + n) * (float)(v + 1));
    Resampler upper(MLB_RESAMPLE_UP, 1, V), downer(MLB_RESAMPLE_DOWN, 1, V);
    REQUIRE(upper.process(x.data(), up.data(), T) == 2 * T);
    REQUIRE(downer.process(up.data(), back.data(), 2 * T) == T);
    float best = 1e9f;
    for (int d = 0; d < 10; ++d)  // group delay of the two half-band stages: a few samples
    {
      float err = 0.f;
      for (int t = 2; t < T - 1; ++t)
        for (int n = 0; n < 64; ++n)
        {
          const int i = t * 64 + n + d;
          err = std::max(err, std::fabs(back[((size_t)(i / 64) * V + 0) * 64 + i % 64] - x[((size_t)t * V + 0) * 64 + n]));
        }
      best = std::min(best, err);
    }
    REQUIRE(best < 0.05f);
  }

  // ---- Voice bank: one note on at frame 10, velocity 0.8, no glide: gate and pitch step there ----
  {
    const int32_t idx[1] = {1};
    const float glide[1] = {0.f}, drift[1] = {0.f}, bend[1] = {7.f};
    VoiceBank vb(1, 48000.f, idx, glide, drift, bend);
    mlb_voice_events ev[2];
    std::memset(ev, 0, sizeof(ev));
    ev[0].n_events = 1;
    ev[0].time[0] = 10, ev[0].type[0] = MLB_EV_NOTE_ON, ev[0].flags[0] = MLB_EVF_RESET;
    ev[0].value1[0] = 5.f, ev[0].value2[0] = 0.8f;
    std::vector<float> rows((size_t)2 * MLB_VOICE_ROWS * 64);
    vb.process(ev, rows.data(), 2);
    const float* pitch = rows.data();            // row 0
    const float* gate = rows.data() + 64;        // row 1
    const float* voice = rows.data() + 2 * 64;   // row 2
    const float* time1 = rows.data() + (size_t)(MLB_VOICE_ROWS + 7) * 64;  // elapsed time, second vector
    REQUIRE(gate[9] == 0.f && gate[10] == 0.8f && gate[63] == 0.8f);
    REQUIRE(pitch[9] == 0.f && pitch[11] == 5.f);
    REQUIRE(voice[0] == 0.f);  // voiceIndex - 1
    // the note-on's age reset and eventAgeStep = 1 precede the frames written before it (.cpp:152-165)
    REQUIRE(std::fabs(time1[63] - 128.f / 48000.f) < 1e-7f);
  }

  // ---- events -> signals -> voice DSP in one call: sine(kPitch row) * kGate row, two voices, the mix bus
  //      is their sum; voice 1 never gets a note and stays silent ----
  {
    const int32_t idx[2] = {1, 2};
    const float glide[2] = {0.f, 0.f}, drift[2] = {0.f, 0.f}, bend[2] = {0.f, 0.f};
    VoiceBank vb(2, 48000.f, idx, glide, drift, bend);
    Graph g;
    g.output(g.multiply(g.sine(g.input(0)), g.input(1)));
    DeviceBank bank(g, 2);
    bank.commit();
    mlb_voice_events ev[3 * 2];
    std::memset(ev, 0, sizeof(ev));
    ev[0].n_events = 1;
    ev[0].time[0] = 10, ev[0].type[0] = MLB_EV_NOTE_ON, ev[0].flags[0] = MLB_EVF_RESET;
    ev[0].value1[0] = 0.01f, ev[0].value2[0] = 0.5f;
    std::vector<float> out((size_t)3 * 2 * 64), mix((size_t)3 * 64);
    processEvents(vb, bank, ev, out.data(), mix.data(), 3);
    float peak = 0.f;
    bool silent1 = true, mixok = true, pre = true;
    for (int t = 0; t < 3; ++t)
      for (int n = 0; n < 64; ++n)
      {
        const float a = out[((size_t)t * 2 + 0) * 64 + n], b = out[((size_t)t * 2 + 1) * 64 + n];
        if (t == 0 && n < 10 && a != 0.f) pre = false;
        peak = std::max(peak, std::fabs(a));
        silent1 = silent1 && b == 0.f;
        mixok = mixok && mix[(size_t)t * 64 + n] == a + b;
      }
    REQUIRE(pre);                        // gate is 0 before the note on
    REQUIRE(peak > 0.4f && peak < 0.501f);  // a 480 Hz sine at velocity 0.5
    REQUIRE(silent1);
    REQUIRE(mixok);
  }

  std::printf("%s: %d assertions, %d failed, %lld kernels launched\n", g_fail ? "FAILED" : "ALL PASSED", g_checks,
              g_fail, mlb_kernel_launches());
  return g_fail ? 1 : 0;
}
