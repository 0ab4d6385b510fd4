// voice_kernel.cuh -- K7: EventsToSignals::Voice for V voices.
// Reference: source/app/MLEventsToSignals.cpp:47-263 (E below), MLEventsToSignals.h:99-168,
// glides source/DSP/MLDSPGens.h:433-590 (G).  One lane = one Voice; a launch runs n_blocks vectors:
// beginProcess (drift), the vector's note events (sample-accurate gate / pitch glide / age), the
// instantaneous controller values, endProcess (vector-accurate LinearGlides, bend and drift added to
// pitch).  Input: one 72-byte mlb_voice_events record per voice and vector; output: up to 8 rows.
// Scalar state is SoA [word][V]; the six LinearGlide::mCurrVec rows are only touched while a glide
// is moving (an idle LinearGlide's row equals its target in every lane of the row).
#pragma once
#include "functors.cuh"

namespace mlb
{
enum VoiceGlide { VG_BEND = 0, VG_MOD, VG_X, VG_Y, VG_Z, VG_DRIFT, VG_PRESSURE, VG_COUNT };
enum VoiceState
{
  VS_PG_CURR = 0, VS_PG_STEP, VS_PG_TARGET, VS_PG_REM, VS_PG_PER, VS_PG_DY,  // pitchGlide (G:517-590)
  VS_GL = 6,                       // 6 x {step, target, vectorsRemaining, spare}
  VS_VEL = VS_GL + 4 * VG_COUNT,   // currentVelocity, currentPitch, bend, mod, x, y, z
  VS_PITCH, VS_BEND, VS_MOD, VS_X, VS_Y, VS_Z, VS_PRESSURE,
  VS_AGE, VS_AGE_STEP, VS_SEED, VS_DRIFT_COUNTER, VS_NEXT_DRIFT, VS_CUR_DRIFT,
  VS_COUNT
};
enum VoiceCoef { VC_GLIDE_SAMPLES = 0, VC_DRIFT_AMOUNT, VC_BEND_RANGE, VC_VOICE_ROW, VC_COUNT };

struct VoiceArgs
{
  const mlb_voice_events* ev;  // [T][V]
  float* out;                  // [T][8][V][64]
  uint32_t* state;             // [VS_COUNT][V]
  const float* coef;           // [VC_COUNT][V]
  float* grows;                // [VG_COUNT][V][64]  LinearGlide::mCurrVec
  int V, T;
  unsigned row_mask;
  float sr;                    // Voice::sr (a double in the reference, set from this value)
  float gl_per, gl_dy;         // LinearGlide coefficients of bend/mod/x/y/z (sr * kGlideTimeSeconds)
  float dr_per, dr_dy;         // ... of pitchDriftGlide (sr * kDriftTimeSeconds)
  float pc_per, pc_dy;         // ... of the channel-pressure SmoothedController (int(sr * kControllerGlideTimeSeconds))
  int midi;                    // MLB_VOICES_MIDI: z row += smoothed channel pressure (processVector, E:432-447)
};

struct VoiceRegs
{
  uint32_t pg[4];
  float pg_per_f, pg_dy;
  float vel, pitch;
  uint32_t age, age_step;
  int next_frame;
};

// SampleAccurateLinearGlide::setGlideTimeInSamples, G:527-532
MLB_DEV void voice_set_glide_time(VoiceRegs& r, float t)
{
  int n = cvt_trunc(t);
  if (n < 1) n = 1;
  r.pg_per_f = __int2float_rn(n);
  r.pg_dy = __fdiv_rn(1.0f, __int2float_rn(n));
}
// one output frame, E:134-140: gate, glided pitch, age -> seconds (samplesToSeconds, E:12-18)
// Row tiles in shared memory: [lane][65 floats] per warp -- a lane walks its own row (frames can be
// revisited by a retrigger), then the warp stores the 32 rows with full 128-byte lines.
constexpr int kVoiceTileStride = 65;
constexpr int kVoiceTileFloats = 32 * kVoiceTileStride;

MLB_DEV void voice_frame(VoiceRegs& r, float sr, int t, float gate_v, float* gate, float* pitch, float* tm,
                        bool want_time)
{
  gate[t] = gate_v;
  const float co[2] = {r.pg_per_f, r.pg_dy};
  pitch[t] = sample_glide_tick<true>(r.pitch, r.pg, co);
  r.age += r.age_step;
  if (want_time) tm[t] = __double2float_rn(__ddiv_rn((double)r.age, (double)sr));  // FP64 divide only if the row is wanted
}
// the warp's 32 rows of one output plane: tile[j][0..63] -> plane[(v0 + j)][0..63], two 128-B lines per row
MLB_DEV void voice_store_tile(const float* tile, float* plane, int v0, int V, int lane)
{
  __syncwarp();
#pragma unroll 4
  for (int j = 0; j < 32; ++j)
  {
    if (v0 + j >= V) break;
    float* dst = plane + (size_t)(v0 + j) * MLB_BLOCK;
    __stcs(dst + lane, tile[j * kVoiceTileStride + lane]);
    __stcs(dst + 32 + lane, tile[j * kVoiceTileStride + 32 + lane]);
  }
  __syncwarp();
}

// LinearGlide::operator()(float), G:459-505.  The row mCurrVec lives in delay memory and is touched only
// while the glide moves: an idle glide's row equals DSPVector(target) (landing writes the target into
// every element, construction and setValue leave zeros with target zero).
struct GlidePlan
{
  int mode;  // -1 idle, 0 land on target, 1 start, 2 continue
  float step, target, cv;
};
// st: step, target, vectorsRemaining, (spare)
MLB_DEV GlidePlan glide_plan(uint32_t* st, float f, float per_f, float dy, const float* row)
{
  GlidePlan g;
  g.step = u2f(st[0]), g.target = u2f(st[1]);
  g.cv = 0.f;
  const int rem_prev = (int32_t)st[2];
  int remaining = rem_prev;
  const int per = cvt_trunc(per_f);
  if (f != g.target)
  {
    g.target = f;
    remaining = per;
  }
  if (remaining < 0)
    g.mode = -1;
  else if (remaining == 0)
  {
    g.mode = 0;
    g.step = 0.f;
    remaining--;
  }
  else if (remaining == per)
  {
    g.mode = 1;
    // currentValue = mCurrVec[63]; an idle row is its (previous) target
    g.cv = rem_prev < 0 ? u2f(st[1]) : row[MLB_BLOCK - 1];
    g.step = __fmul_rn(__fsub_rn(g.target, g.cv), dy);
    remaining--;
  }
  else
  {
    g.mode = 2;
    remaining--;
  }
  st[0] = f2u(g.step), st[1] = f2u(g.target), st[2] = (uint32_t)remaining;
  return g;
}
// Run one glide for this lane's vector; emit(q, float4) receives samples 4q..4q+3 when WANT is set.
// The three modes are hoisted out of the sample loop; an idle glide whose row nobody wants costs nothing.
template <bool WANT, class Emit>
MLB_DEV void glide_run(const GlidePlan& g, float* row, bool live, Emit emit)
{
  float4* row4 = reinterpret_cast<float4*>(row);
  if (g.mode <= 0)
  {
    // landing: the row becomes DSPVector(target); it is not stored -- idle rows are never read
    if (WANT)
    {
      const float4 y = make_float4(g.target, g.target, g.target, g.target);
#pragma unroll 4
      for (int q = 0; q < 16; ++q) emit(q, y);
    }
  }
  else if (g.mode == 1)
  {
#pragma unroll 2
    for (int q = 0; q < 16; ++q)
    {
      float4 y;
      y.x = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q), g.step));
      y.y = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q + 1), g.step));
      y.z = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q + 2), g.step));
      y.w = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q + 3), g.step));
      if (live) row4[q] = y;
      if (WANT) emit(q, y);
    }
  }
  else
  {
#pragma unroll 1
    for (int h = 0; h < 2; ++h)
    {
      float4 buf[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) buf[q] = row4[8 * h + q];
#pragma unroll
      for (int q = 0; q < 8; ++q)
      {
        float4 y = buf[q];
        y.x = __fadd_rn(y.x, g.step), y.y = __fadd_rn(y.y, g.step), y.z = __fadd_rn(y.z, g.step), y.w = __fadd_rn(y.w, g.step);
        if (live) row4[8 * h + q] = y;
        if (WANT) emit(8 * h + q, y);
      }
    }
  }
}

__global__ void __launch_bounds__(128, 3) voice_bank_kernel(const VoiceArgs a)
{
  extern __shared__ float voice_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const int v0 = v_raw - lane;
  if (v0 >= a.V) return;                      // whole warp out of range
  const bool live = v_raw < a.V;
  const int v = live ? v_raw : a.V - 1;       // dead lanes shadow the last voice, their stores are masked
  const size_t V = (size_t)a.V;
  const bool want_time = (a.row_mask & 128u) != 0;
  const int n_tiles = want_time ? 3 : 2;  // gate, pitch (+ elapsed time)
  float* const tiles = voice_smem + (size_t)warp * n_tiles * kVoiceTileFloats;
  float* const gate = tiles + lane * kVoiceTileStride;
  float* const pitch = gate + kVoiceTileFloats;
  float* const tm = pitch + kVoiceTileFloats;
  uint32_t st[VS_COUNT];
#pragma unroll
  for (int i = 0; i < VS_COUNT; ++i) st[i] = a.state[(size_t)i * V + v];
  const float glide_samples = a.coef[(size_t)VC_GLIDE_SAMPLES * V + v];  // (float)pitchGlideTimeInSamples
  const float drift_amount = a.coef[(size_t)VC_DRIFT_AMOUNT * V + v];
  const float bend_range = a.coef[(size_t)VC_BEND_RANGE * V + v];
  const float voice_row = a.coef[(size_t)VC_VOICE_ROW * V + v];
  float* grow[VG_COUNT];
#pragma unroll
  for (int i = 0; i < VG_COUNT; ++i) grow[i] = a.grows + ((size_t)i * V + v) * MLB_BLOCK;

  VoiceRegs r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.pg[i] = st[VS_PG_CURR + i];
  r.pg_per_f = u2f(st[VS_PG_PER]), r.pg_dy = u2f(st[VS_PG_DY]);
  r.vel = u2f(st[VS_VEL]), r.pitch = u2f(st[VS_PITCH]);
  r.age = st[VS_AGE], r.age_step = st[VS_AGE_STEP];
  float cur[5] = {u2f(st[VS_BEND]), u2f(st[VS_MOD]), u2f(st[VS_X]), u2f(st[VS_Y]), u2f(st[VS_Z])};
  float cur_pressure = u2f(st[VS_PRESSURE]);
  uint32_t seed = st[VS_SEED];
  int drift_counter = (int)st[VS_DRIFT_COUNTER], next_drift = (int)st[VS_NEXT_DRIFT];
  float cur_drift = u2f(st[VS_CUR_DRIFT]);


  for (int t = 0; t < a.T; ++t)
  {
    // ---- beginProcess, E:90-124 ----
    r.next_frame = 0;
    drift_counter += MLB_BLOCK;
    if (drift_counter >= next_drift)
    {
      const float d = noise_tick(seed);  // RandomScalarSource::getFloat == NoiseGen's LCG step
      const float next_mul = __fadd_rn(1.0f, fabsf(noise_tick(seed)));
      cur_drift = d;
      drift_counter = 0;
      next_drift = __double2int_rz(__dmul_rn(__dmul_rn((double)a.sr, (double)next_mul), 8.0));
    }
    // ---- the vector's note events (E:126-220) and the tail fill of endProcess (E:224-236) ----
    // The reference runs, per event, [pre-actions; frames up to the event; post-actions], then fills the
    // rest of the vector.  Here every lane emits ONE frame per iteration of a warp-uniform loop and
    // steps its own event cursor in between, so lanes with different event times do not serialise.
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(a.ev + ((size_t)t * V + v));
    const uint32_t head = rec[0], times = rec[1], types = rec[2], flags = rec[3];
    const int n_events = min((int)(head & 0xFFu), MLB_VOICE_MAX_EVENTS);
    const unsigned set_mask = (head >> 8) & 0xFFu;
    // Decode the (at most 4) note events once.  ev_d = frame before which the event takes effect
    // (retrigger: its gate-0 frame is ev_d - 1), ev_kind = 0 ignored / 1 on / 2 retrigger / 3 off.
    int ev_d[MLB_VOICE_MAX_EVENTS], ev_kind[MLB_VOICE_MAX_EVENTS];
    bool monotone = true;
    {
      int nf = 0;  // nextFrameToProcess as the reference would see it
#pragma unroll
      for (int k = 0; k < MLB_VOICE_MAX_EVENTS; ++k)
      {
        int dest = (int)((times >> (8 * k)) & 0xFFu);
        dest = dest > MLB_BLOCK ? MLB_BLOCK : dest;
        const int type = (int)((types >> (8 * k)) & 0xFFu);
        int kind = type == MLB_EV_NOTE_ON ? 1 : type == MLB_EV_NOTE_RETRIG ? 2 : type == MLB_EV_NOTE_OFF ? 3 : 0;
        if (k >= n_events) kind = 0;
        if (kind == 2 && dest == 0) dest = 1;
        const int bound = kind == 2 ? dest - 1 : dest;
        if (kind != 0)
        {
          if (bound < nf) monotone = false;  // the reference would step back (e.g. a pedal-release note-off at frame 0)
          nf = dest;
        }
        ev_d[k] = dest, ev_kind[k] = kind;
      }
    }
    if (__all_sync(0xffffffffu, monotone))
    {
      // Fast path: frames 0..63 are each emitted once, in order.  Event k's pre-actions (age reset, glide
      // time) apply as soon as event k-1 is done, its post-actions (new pitch / velocity) before frame ev_d[k].
      int k = 0;
      // the current event's frame and kind are kept in scalars (cur_d = 1000 when there is none), so that the
      // 64-iteration frame loop below never indexes the decoded event arrays
      int cur_d = 1000, cur_kind = 0;
      auto advance = [&]()  // make the next non-ignored event current and apply its pre-actions
      {
        while (k < MLB_VOICE_MAX_EVENTS && ev_kind[k] == 0) ++k;
        cur_d = 1000, cur_kind = 0;
        if (k < MLB_VOICE_MAX_EVENTS)
        {
          cur_d = ev_d[k], cur_kind = ev_kind[k];
          const unsigned fl = (flags >> (8 * k)) & 0xFFu;
          if (cur_kind != 3)
          {
            if (fl & MLB_EVF_RESET) r.age = 0;
            r.age_step = 1;
          }
          if (cur_kind == 1) voice_set_glide_time(r, (fl & MLB_EVF_GLIDE) ? glide_samples : 0.f);
        }
      };
      auto complete = [&]()  // post-actions of the current event, then move on
      {
        if (cur_kind == 3)
          r.vel = 0.f;
        else
          r.pitch = u2f(rec[4 + k]), r.vel = u2f(rec[8 + k]);
        ++k;
        advance();
      };
      advance();
#pragma unroll 1
      for (int f = 0; f < MLB_BLOCK; ++f)
      {
        while (cur_d <= f) complete();
        const bool retrig_frame = cur_kind == 2 && cur_d - 1 == f;
        voice_frame(r, a.sr, f, retrig_frame ? 0.f : r.vel, gate, pitch, tm, want_time);
      }
      while (k < MLB_VOICE_MAX_EVENTS) complete();  // events at frame 64: only their value changes remain
      r.next_frame = MLB_BLOCK;
    }
    else
    {
      int k = 0;
      bool pre_applied = false, retrig_frame_done = false;
      bool active = true;
      while (active)
      {
        int emit_at = -1;
        float emit_gate = 0.f;
        while (k < n_events)
        {
          int dest = (int)((times >> (8 * k)) & 0xFFu);
          dest = dest > MLB_BLOCK ? MLB_BLOCK : dest;
          const int type = (int)((types >> (8 * k)) & 0xFFu);
          const unsigned fl = (flags >> (8 * k)) & 0xFFu;
          const bool is_on = type == MLB_EV_NOTE_ON, is_rt = type == MLB_EV_NOTE_RETRIG, is_off = type == MLB_EV_NOTE_OFF;
          if (!(is_on || is_rt || is_off))
          {
            ++k;  // kNoteSustain & co: writeNoteEvent does nothing
            continue;
          }
          if (!pre_applied)
          {
            if (is_on || is_rt)
            {
              if (fl & MLB_EVF_RESET) r.age = 0;
              r.age_step = 1;
            }
            if (is_on) voice_set_glide_time(r, (fl & MLB_EVF_GLIDE) ? glide_samples : 0.f);
            pre_applied = true;
          }
          if (is_rt && dest == 0) dest = 1;
          const int bound = is_rt ? dest - 1 : dest;  // writeOutputFrames(bound)
          if (r.next_frame < bound)
          {
            emit_at = r.next_frame++;
            emit_gate = r.vel;
            break;
          }
          r.next_frame = bound;  // writeOutputFrames always ends with nextFrameToProcess = endFrame
          if (is_rt && !retrig_frame_done)
          {
            retrig_frame_done = true;  // the retrigger frame: gate 0 at dest - 1 (E:183-189)
            emit_at = dest - 1;
            emit_gate = 0.f;
            break;
          }
          // post-actions
          if (is_off)
            r.vel = 0.f;
          else
          {
            r.pitch = u2f(rec[4 + k]);
            r.vel = u2f(rec[8 + k]);
          }
          if (is_rt) r.next_frame = dest;
          ++k;
          pre_applied = false, retrig_frame_done = false;
        }
        if (emit_at < 0)
        {
          if (r.next_frame < MLB_BLOCK)
          {
            emit_at = r.next_frame++;
            emit_gate = r.vel;
          }
          else
            active = false;
        }
        if (emit_at >= 0) voice_frame(r, a.sr, emit_at, emit_gate, gate, pitch, tm, want_time);
      }
    }
    if (set_mask & MLB_SET_BEND) cur[0] = u2f(rec[12]);
    if (set_mask & MLB_SET_MOD) cur[1] = u2f(rec[13]);
    if (set_mask & MLB_SET_X) cur[2] = u2f(rec[14]);
    if (set_mask & MLB_SET_Y) cur[3] = u2f(rec[15]);
    if (set_mask & MLB_SET_Z) cur[4] = u2f(rec[16]);
    if (set_mask & MLB_SET_PRESSURE) cur_pressure = u2f(rec[17]);
    // ---- endProcess, E:222-262 ----
    if (r.vel == 0.f) cur[4] = 0.f;
    GlidePlan gp[VG_COUNT];
#pragma unroll
    for (int i = 0; i < VG_COUNT; ++i)
    {
      if (i == VG_PRESSURE && !a.midi)
      {
        gp[i].mode = -1, gp[i].step = gp[i].target = gp[i].cv = 0.f;
        continue;
      }
      gp[i] = glide_plan(&st[VS_GL + 4 * i], i == VG_DRIFT ? cur_drift : (i == VG_PRESSURE ? cur_pressure : cur[i]),
                         i == VG_DRIFT ? a.dr_per : (i == VG_PRESSURE ? a.pc_per : a.gl_per),
                         i == VG_DRIFT ? a.dr_dy : (i == VG_PRESSURE ? a.pc_dy : a.gl_dy), grow[i]);
    }
    float* const planes = a.out + (size_t)t * MLB_VOICE_ROWS * V * MLB_BLOCK;
    const size_t plane_floats = V * MLB_BLOCK;
    // pitch += bendGlide * pitchBend * (1/12); pitch += driftSig * driftAmount * kDriftScale  (E:255-261)
    glide_run<true>(gp[VG_BEND], grow[VG_BEND], live, [&](int q, float4 y)
    {
      const float k12 = 1.f / 12;
      float* p = pitch + 4 * q;
      p[0] = __fadd_rn(p[0], __fmul_rn(__fmul_rn(y.x, bend_range), k12));
      p[1] = __fadd_rn(p[1], __fmul_rn(__fmul_rn(y.y, bend_range), k12));
      p[2] = __fadd_rn(p[2], __fmul_rn(__fmul_rn(y.z, bend_range), k12));
      p[3] = __fadd_rn(p[3], __fmul_rn(__fmul_rn(y.w, bend_range), k12));
    });
    glide_run<true>(gp[VG_DRIFT], grow[VG_DRIFT], live, [&](int q, float4 y)
    {
      float* p = pitch + 4 * q;
      p[0] = __fadd_rn(p[0], __fmul_rn(__fmul_rn(y.x, drift_amount), 0.02f));
      p[1] = __fadd_rn(p[1], __fmul_rn(__fmul_rn(y.y, drift_amount), 0.02f));
      p[2] = __fadd_rn(p[2], __fmul_rn(__fmul_rn(y.z, drift_amount), 0.02f));
      p[3] = __fadd_rn(p[3], __fmul_rn(__fmul_rn(y.w, drift_amount), 0.02f));
    });
    if (a.row_mask & 1u) voice_store_tile(tiles + kVoiceTileFloats, planes + 0 * plane_floats, v0, a.V, lane);    // kPitch
    if (a.row_mask & 2u) voice_store_tile(tiles, planes + 1 * plane_floats, v0, a.V, lane);                        // kGate
    if (want_time) voice_store_tile(tiles + 2 * kVoiceTileFloats, planes + 7 * plane_floats, v0, a.V, lane);      // kElapsedTime
    __syncwarp();
    // the glide-only rows reuse the gate tile: kZ(3) kX(4) kY(5) kMod(6) <- z, x, y, mod glides; kVoice(2) constant
    if (a.row_mask & 4u)
    {
#pragma unroll 4
      for (int n = 0; n < MLB_BLOCK; ++n) gate[n] = voice_row;
      voice_store_tile(tiles, planes + 2 * plane_floats, v0, a.V, lane);
    }
    const int glide_of_row[4] = {VG_Z, VG_X, VG_Y, VG_MOD};
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
    {
      const int row = 3 + rr, gi = glide_of_row[rr];
      if ((a.row_mask >> row) & 1u)
      {
        glide_run<true>(gp[gi], grow[gi], live, [&](int q, float4 y)
        { gate[4 * q] = y.x, gate[4 * q + 1] = y.y, gate[4 * q + 2] = y.z, gate[4 * q + 3] = y.w; });
        if (gi == VG_Z && a.midi)  // voices[v].outputs.row(kZ) += controllers[128].output
          glide_run<true>(gp[VG_PRESSURE], grow[VG_PRESSURE], live, [&](int q, float4 y)
          {
            gate[4 * q] = __fadd_rn(gate[4 * q], y.x), gate[4 * q + 1] = __fadd_rn(gate[4 * q + 1], y.y);
            gate[4 * q + 2] = __fadd_rn(gate[4 * q + 2], y.z), gate[4 * q + 3] = __fadd_rn(gate[4 * q + 3], y.w);
          });
        voice_store_tile(tiles, planes + (size_t)row * plane_floats, v0, a.V, lane);
      }
      else
      {
        glide_run<false>(gp[gi], grow[gi], live, [](int, float4) {});  // the glide still advances
        if (gi == VG_Z && a.midi) glide_run<false>(gp[VG_PRESSURE], grow[VG_PRESSURE], live, [](int, float4) {});
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) st[VS_PG_CURR + i] = r.pg[i];
  st[VS_PG_PER] = f2u(r.pg_per_f), st[VS_PG_DY] = f2u(r.pg_dy);
  st[VS_VEL] = f2u(r.vel), st[VS_PITCH] = f2u(r.pitch);
  st[VS_AGE] = r.age, st[VS_AGE_STEP] = r.age_step;
  st[VS_BEND] = f2u(cur[0]), st[VS_MOD] = f2u(cur[1]), st[VS_X] = f2u(cur[2]), st[VS_Y] = f2u(cur[3]), st[VS_Z] = f2u(cur[4]);
  st[VS_PRESSURE] = f2u(cur_pressure);
  st[VS_SEED] = seed, st[VS_DRIFT_COUNTER] = (uint32_t)drift_counter, st[VS_NEXT_DRIFT] = (uint32_t)next_drift;
  st[VS_CUR_DRIFT] = f2u(cur_drift);
  if (live)
  {
#pragma unroll
    for (int i = 0; i < VS_COUNT; ++i) a.state[(size_t)i * V + v] = st[i];
  }
}

// processVector's MPE tail (E:448-460): pitch, x, y, z, mod rows += the rows of the instrument's main voice.
// One thread per float4 of a (block, voice) row; planes [T][8][V][64].
__global__ void __launch_bounds__(256) voice_mpe_add_kernel(float* out, const int32_t* main_voice, int V, int T,
                                                             unsigned row_mask)
{
  const size_t n4 = (size_t)T * V * 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
  {
    const int q = (int)(i & 15u);
    const size_t tv = i >> 4;
    const int v = (int)(tv % (size_t)V);
    const size_t t = tv / (size_t)V;
    const int m = main_voice[v];
    if (m < 0) continue;
    const int rows[5] = {0, 3, 4, 5, 6};  // kPitch, kZ, kX, kY, kMod
#pragma unroll
    for (int k = 0; k < 5; ++k)
    {
      if (!((row_mask >> rows[k]) & 1u)) continue;
      float4* plane = reinterpret_cast<float4*>(out + (t * MLB_VOICE_ROWS + rows[k]) * (size_t)V * MLB_BLOCK);
      float4 a = plane[(size_t)v * 16 + q];
      const float4 b = plane[(size_t)m * 16 + q];
      a.x = __fadd_rn(a.x, b.x), a.y = __fadd_rn(a.y, b.y), a.z = __fadd_rn(a.z, b.z), a.w = __fadd_rn(a.w, b.w);
      plane[(size_t)v * 16 + q] = a;
    }
  }
}

}  // namespace mlb
