(python -m pytest tests/test_gpu_chain.py tests/test_gpu_fullsize.py -x -q -k "fdn or config4 or config_4" 2>&1 | tail -n 2
python tools/bench_configs.py --only 4 2>&1 | cut -c1-260
for W in 4 5 6 7 8; do MLB_FDN_WARPS=$W python tools/bench_configs.py --only 4 2>&1 | cut -c60-200; done
) > gpurun_out/r2j_team.txt 2>&1
