// oracle/ref/mle2s.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference EventsToSignals::Voice (source/app/MLEventsToSignals.{h,cpp})
// for a bank of voices.  oracle/Makefile compiles this file together with the reference's own
// MLEventsToSignals.cpp, MLSymbol.cpp and MLText.cpp, in place, into oracle/_ref/libmle2s.so.
// Nothing from the reference is copied into this repository.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load the library.
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "MLEventsToSignals.h"
#include "mlb200.h"

using namespace ml;

namespace
{
struct VoiceBank
{
  std::vector<EventsToSignals::Voice> voices;
  std::vector<EventsToSignals::SmoothedController> pressure;  // controllers[128], one copy per voice
  std::vector<float> pitchBend;
  bool midi{false};
};
}  // namespace

extern "C"
{
struct mle2s_bank
{
  VoiceBank b;
};

// Voice() + reset() + setSampleRate + setPitchGlideInSeconds + setDriftAmount, exactly what
// EventsToSignals' constructor and setters do for each of its voices (.cpp:283-300, 304-318, 858-873)
mle2s_bank* mle2s_bank_create(int V, float sr, const int32_t* voiceIndex, const float* glideSeconds,
                              const float* driftAmount, const float* pitchBend, unsigned flags)
{
  auto* h = new mle2s_bank;
  h->b.voices.resize(V);
  h->b.pressure.resize(V);
  h->b.midi = (flags & MLB_VOICES_MIDI) != 0;
  for (auto& c : h->b.pressure) c.setSampleRate(sr);
  h->b.pitchBend.assign(pitchBend, pitchBend + V);
  for (int v = 0; v < V; ++v)
  {
    auto& vc = h->b.voices[v];
    vc.voiceIndex = voiceIndex[v];
    vc.reset();
    vc.outputs.row(kVoice) = DSPVector((float)voiceIndex[v] - 1);
    vc.setSampleRate(sr);
    vc.setPitchGlideInSeconds(glideSeconds[v]);
    vc.setDriftAmount(driftAmount[v]);
  }
  return h;
}
void mle2s_bank_destroy(mle2s_bank* h) { delete h; }

// T vectors for every voice.  ev [T][V]; out [T][8][V][64].  Returns seconds of wall time.
double mle2s_bank_process(mle2s_bank* h, int T, const mlb_voice_events* ev, float* out, int nthreads)
{
  const int V = (int)h->b.voices.size();
  if (nthreads < 1) nthreads = 1;
  nthreads = std::min(nthreads, std::max(1, V));
  auto worker = [&](int v0, int v1)
  {
    for (int v = v0; v < v1; ++v)
    {
      auto& vc = h->b.voices[v];
      for (int t = 0; t < T; ++t)
      {
        const mlb_voice_events& r = ev[(size_t)t * V + v];
        vc.beginProcess();
        for (int k = 0; k < r.n_events; ++k)
        {
          Event e;
          e.type = r.type[k];
          e.time = r.time[k];
          e.value1 = r.value1[k];
          e.value2 = r.value2[k];
          vc.writeNoteEvent(e, 1, (r.flags[k] & MLB_EVF_GLIDE) != 0, (r.flags[k] & MLB_EVF_RESET) != 0);
        }
        // the instantaneous values event routing writes into the voice (.cpp:640-820)
        if (r.set_mask & MLB_SET_BEND) vc.currentPitchBend = r.bend;
        if (r.set_mask & MLB_SET_MOD) vc.currentMod = r.mod;
        if (r.set_mask & MLB_SET_X) vc.currentX = r.x;
        if (r.set_mask & MLB_SET_Y) vc.currentY = r.y;
        if (r.set_mask & MLB_SET_Z) vc.currentZ = r.z;
        vc.endProcess(h->b.pitchBend[v]);
        if (h->b.midi)
        {
          // processVector's tail in MIDI mode (.cpp:432-447): smooth the channel pressure, add it to z
          auto& pc = h->b.pressure[v];
          if (r.set_mask & MLB_SET_PRESSURE) pc.inputValue = r.pressure;
          pc.process();
          vc.outputs.row(kZ) += pc.output;
        }
        for (int row = 0; row < kNumVoiceOutputRows; ++row)
          store(vc.outputs.constRow(row), out + (((size_t)t * kNumVoiceOutputRows + row) * V + v) * 64);
      }
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1)
    worker(0, V);
  else
  {
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i)
      th.emplace_back(worker, (int)((long long)V * i / nthreads), (int)((long long)V * (i + 1) / nthreads));
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// The complete reference EventsToSignals (allocator included) for ONE instrument, so that tests can
// show what per-voice event records its routing produces for a MIDI phrase.
struct mle2s_full
{
  EventsToSignals e2s;
};
mle2s_full* mle2s_full_create(float sr, int polyphony, float glideSeconds, float driftAmount, int unison, int mpe)
{
  auto* h = new mle2s_full;
  if (mpe) h->e2s.setProtocol(Symbol("MPE"));
  h->e2s.setSampleRate(sr);
  h->e2s.setPolyphony(polyphony);
  h->e2s.setPitchGlideInSeconds(glideSeconds);
  h->e2s.setDriftAmount(driftAmount);
  h->e2s.setUnison(unison != 0);
  return h;
}
void mle2s_full_destroy(mle2s_full* h) { delete h; }
void mle2s_full_add_event(mle2s_full* h, int type, int channel, int sourceIdx, int time, float v1, float v2)
{
  Event e;
  e.type = (uint8_t)type, e.channel = (uint8_t)channel, e.sourceIdx = (uint16_t)sourceIdx;
  e.time = time, e.value1 = v1, e.value2 = v2;
  h->e2s.addEvent(e);
}
// one vector starting at frame `start` of the event buffer's time base; out [polyphony][8][64]
void mle2s_full_process(mle2s_full* h, int start, float* out)
{
  h->e2s.processVector(start);
  const int P = (int)h->e2s.getPolyphony();
  for (int v = 0; v < P; ++v)
    for (int row = 0; row < kNumVoiceOutputRows; ++row)
      store(h->e2s.getVoice(v).outputs.constRow(row), out + ((size_t)v * kNumVoiceOutputRows + row) * 64);
}
void mle2s_full_clear_events(mle2s_full* h) { h->e2s.clearEvents(); }
}  // extern "C"
