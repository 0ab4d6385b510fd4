#!/bin/bash
# One GPU call that regenerates the round-2 evidence under gpurun_out/ (copied into profiles/ afterwards):
#   configs timings at HEAD, the ncu launch list of the bench command, one `ncu --set full` capture of the
#   chain kernel (DRAM traffic per launch), compute-sanitizer memcheck + racecheck on small parity tests.
set -x
python tools/bench_configs.py > gpurun_out/configs_r2.jsonl 2> gpurun_out/configs_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 4 -c 1 -o gpurun_out/prof_chainA_r2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --e2e-steps 1 --no-parity > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_team_kernel -s 2 -c 1 -o gpurun_out/prof_team_r2 \
    python tools/bench_configs.py --only 2 --steps 2 > gpurun_out/ncu_team.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_chain.py tests/test_gpu_coeff_rows.py -x -q -k "not full_size and not large_bank" > gpurun_out/sanitizer_memcheck_r2.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_r2.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_chain.py -x -q -k "config_a_bit_exact or dynamic_work_units or mix_bus" > gpurun_out/sanitizer_racecheck_r2.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_r2.log
tail -3 gpurun_out/sanitizer_memcheck_r2.log gpurun_out/sanitizer_racecheck_r2.log
cut -c1-220 gpurun_out/configs_r2.jsonl
