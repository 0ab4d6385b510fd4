// tests/cpp/test_published.cpp -- CPU-only.  mlb::PublishedSignal (include/mlb200_host.hpp) in lockstep with
// the reference's SignalProcessor::PublishedSignal (source/app/MLSignalProcessor.{h,cpp}, compiled in place)
// on random sequences of writeQuick / writeQuickVert / read / readLatest / peekLatest.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mlb200_host.hpp"
#include "MLSignalProcessor.h"

static int g_fail = 0, g_checks = 0;
#define REQUIRE(c) do { ++g_checks; if (!(c)) { ++g_fail; std::printf("REQUIRE failed: %s line %d\n", #c, __LINE__); } } while (0)
static unsigned rnd(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <size_t CH>
static void run(unsigned seed, int octaves, int maxFrames, int maxVoices)
{
  unsigned s = seed;
  ml::SignalProcessor::PublishedSignal ref(maxFrames, maxVoices, (int)CH, octaves);
  mlb::PublishedSignal mine(maxFrames, maxVoices, (int)CH, octaves);
  float counter = 0.f;
  std::vector<float> a(4096), b(4096);
  for (int step = 0; step < 600; ++step)
  {
    const unsigned op = rnd(s) % 6;
    if (op <= 2)
    {
      ml::DSPVectorArray<CH> v;
      for (size_t j = 0; j < CH; ++j)
        for (int n = 0; n < 64; ++n) v.row(j)[n] = counter++;
      const size_t frames = 1 + rnd(s) % (size_t)std::min(64, maxFrames);
      ref.template writeQuick<CH>(v, frames, 0);
      mine.writeQuick(v.getConstBuffer(), frames, 0);
    }
    else if (op == 3)
    {
      float frame[CH];
      for (size_t j = 0; j < CH; ++j) frame[j] = counter++;
      ref.writeQuickVert(frame, CH, 0);
      mine.writeQuickVert(frame, CH, 0);
    }
    else
    {
      const size_t want = rnd(s) % 40;
      std::fill(a.begin(), a.end(), -1.f);
      std::fill(b.begin(), b.end(), -1.f);
      size_t ra = 0, rb = 0;
      if (op == 4)
        ra = ref.read(a.data(), want), rb = mine.read(b.data(), want);
      else if (rnd(s) & 1)
        ra = ref.readLatest(a.data(), want), rb = mine.readLatest(b.data(), want);
      else
        ref.peekLatest(a.data(), want), mine.peekLatest(b.data(), want);
      REQUIRE(ra == rb);
      REQUIRE(std::memcmp(a.data(), b.data(), a.size() * sizeof(float)) == 0);
    }
    REQUIRE(ref.getReadAvailable() == mine.getReadAvailable());
    REQUIRE(ref.getAvailableFrames() == mine.getAvailableFrames());
  }
}

int main()
{
  for (unsigned seed = 1; seed <= 6; ++seed)
  {
    run<1>(seed, 0, 64, 1);
    run<2>(seed, 2, 32, 4);
    run<3>(seed, 4, 16, 2);
    run<4>(seed, 1, 64, 8);
  }
  std::printf("%s: %d assertions, %d failed (checked against the reference)\n", g_fail ? "FAILED" : "ALL PASSED", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
