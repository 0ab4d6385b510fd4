(python -m pytest tests/test_gpu_chain.py tests/test_gpu_coeff_rows.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -n 2
for V in 4096 8192 12288 16384 20480; do PROBE_V=$V python tools/probe_team.py 2>&1 | grep -E "case" | head -n 1 | cut -c1-130; done
) > gpurun_out/r2j_team.txt 2>&1
