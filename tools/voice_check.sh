# one GPU call: voice-bank parity (small + full size + contract E), the K7 timings, optionally an ncu capture
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_voice_bank.py tests/test_gpu_synth.py tests/test_gpu_fullsize.py tests/test_cpp_hostapi.py -m gpu -q -k "voice or synth or events or cpp" 2>&1 | tail -n 8
timeout 300 python tools/bench_configs.py --only voices 2>&1 | tail -n 2 | cut -c1-330
for w in $VOICE_WARPS_SWEEP; do echo "MLB_VOICE_WARPS=$w"; MLB_VOICE_WARPS=$w timeout 300 python tools/bench_configs.py --only voices 2>&1 | tail -n 2 | cut -c1-200; done
if [ -n "$VOICE_NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:voice_bank_kernel -s 1 -c 1 -f -o gpurun_out/prof_voice_r2 \
    python tools/bench_configs.py --only voices --steps 1 > gpurun_out/ncu_voice.log 2>&1
  tail -n 3 gpurun_out/ncu_voice.log
fi
) > gpurun_out/voice_r2.txt 2>&1
cat gpurun_out/voice_r2.txt
