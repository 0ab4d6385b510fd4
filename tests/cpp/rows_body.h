// tests/cpp/rows_body.h -- ONE process function, compiled twice from this very file (like kitchen_body.h):
//   * against the reference itself (oracle/ref/mlref.cpp: `using namespace ml;`), and
//   * against the tracing layer (tests/cpp/test_trace.cpp: `using namespace mlb::tr;`).
// DSPVectorArray<ROWS> used as a value, the way a bank of voices is written in the reference: rowwise arithmetic on
// whole arrays, the row operations of MLDSPOps.h:1056-1383 (repeatRows, stretchRows, zeroPadRows, shiftRows, rotateRows,
// shuffleRows, evenRows / oddRows, separateRows, concatRows, addRows, rowIndex, columnIndex, rangeOpen / rangeClosed)
// Bank<T, ROWS> with array arguments, the "1" forms (multiply1, min1, ...), the array lerp, comparisons, select,
// int <-> float conversions and int arithmetic.  One input row (a frequency), two output rows.
#pragma once

struct RowsState
{
  Bank<SineGen, 4> oscs;
  Bank<OnePole, 4> smooth;
  Bank<Lopass, 2> lps;
};

inline void rowsInit(RowsState& s)
{
  for (int j = 0; j < 4; ++j)
  {
    s.oscs[j].clear();
    s.smooth[j].coeffs = OnePole::makeCoeffs(0.01f * (float)(j + 1));
  }
  static volatile float om[2] = {0.07f, 0.19f};
  for (int j = 0; j < 2; ++j) s.lps[j].coeffs = Lopass::makeCoeffs(om[j], 0.8f);
}

template <class ROWS_CTX>
inline void rowsProcess(ROWS_CTX* ctx, void* state)
{
  RowsState* s = static_cast<RowsState*>(state);
  DSPVector f = ctx->inputs[0];
  // four partials: the frequency row repeated, times 1, 2, 3, 4
  DSPVectorArray<4> ratios = rowIndex<4>() + repeatRows<4>(DSPVector(1.f));
  DSPVectorArray<4> freqs = repeatRows<4>(f) * ratios;
  DSPVectorArray<4> tones = s->oscs(freqs);
  // amplitudes 1, 1/2, 1/3, 1/4, with a ramp across the vector on the odd partials
  DSPVectorArray<4> amps = repeatRows<4>(DSPVector(1.f)) / ratios;
  DSPVectorArray<2> sweep = repeatRows<2>(rangeClosed(0.5f, 1.f));
  amps *= shuffleRows(repeatRows<2>(DSPVector(1.f)), sweep);
  DSPVectorArray<4> voiced = s->smooth(tones * amps);
  // split, filter the odd partials as a pair, put everything back and mix
  DSPVectorArray<2> ev = evenRows(voiced), od = s->lps(oddRows(voiced));
  DSPVectorArray<4> back = concatRows(ev, od);
  DSPVectorArray<4> turned = rotateRows(back, 1) - shiftRows(back, -1) * repeatRows<4>(DSPVector(0.25f));
  DSPVectorArray<6> padded = zeroPadRows<6>(turned);
  DSPVectorArray<3> mid = separateRows<1, 4>(padded);
  // comparisons, select, int <-> float conversions and int arithmetic on the way out
  DSPVector mixed = addRows(turned) * 0.25f;
  DSPVectorInt hot = greaterThan(abs(mixed), DSPVector(0.2f));
  DSPVector folded = select(mixed * 0.5f + sign(mixed) * 0.1f, mixed, hot);
  DSPVectorInt steps = addInt32(roundFloatToInt(folded * 8.f), truncateFloatToInt(columnIndex() * 0.125f));
  DSPVector stair = intToFloat(subtractInt32(steps, DSPVectorInt(3))) * 0.001f;
  DSPVector gate = select(DSPVector(1.f), DSPVector(0.5f), lessThanOrEqual(f, DSPVector(0.0025f)));
  folded = folded * gate + select(stair, DSPVector(0.f), notEqual(within(folded, DSPVector(-0.1f), DSPVector(0.1f)), DSPVector(0.f)));
  DSPVectorArray<4> pushed = max1(min1(multiply1(back, gate), DSPVector(0.9f)), DSPVector(-0.9f));
  pushed = lerp(pushed, turned, 0.25f);
  ctx->outputs[0] = folded + addRows(subtract1(pushed, stair)) * 0.01f + rangeOpen(0.f, 0.001f);
  ctx->outputs[1] = addRows(stretchRows<5>(mid)) * 0.2f + columnIndex() * 1e-6f;
}
