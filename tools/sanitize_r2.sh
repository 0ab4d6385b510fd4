#!/bin/bash
# compute-sanitizer on the final build + an ncu summary of the three-warp team kernel (outputs under gpurun_out/)
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_chain.py -x -q -k "config_a_bit_exact or dynamic_work_units or mix_bus or asynchronous" > gpurun_out/sanitizer_racecheck_r2.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_r2.log
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_chain.py tests/test_gpu_coeff_rows.py tests/test_trace.py -x -q -m gpu -k "not full_size and not large_bank" > gpurun_out/sanitizer_memcheck_r2.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_r2.log
ncu --set full --clock-control none --import-source on -k regex:chain_team_kernel -s 2 -c 1 -o gpurun_out/prof_team_r2 python tools/bench_configs.py --only 2 --steps 2 > gpurun_out/ncu_team.log 2>&1
tail -n 4 gpurun_out/sanitizer_racecheck_r2.log gpurun_out/sanitizer_memcheck_r2.log
