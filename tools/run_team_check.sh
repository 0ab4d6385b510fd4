(python -m pytest tests/test_gpu_chain.py tests/test_gpu_coeff_rows.py tests/test_gpu_fullsize.py tests/test_gpu_robustness.py -x -q 2>&1 | tail -n 3
tail -n 3 gpurun_out/sanitizer_racecheck_r2.log | cut -c1-200
python tools/bench_configs.py --only 2 2>&1 | cut -c1-200
MLB_TEAM_PROF=1 python tools/probe_team.py 2>&1 | grep -E "team prof" | awk "NR%23==1" | head -n 4
for V in 8192 16384 24576 32768; do PROBE_V=$V python tools/probe_team.py 2>/dev/null | head -n 3 | cut -c1-150; done) > gpurun_out/r2j_team.txt 2>&1
