(python -m pytest tests/test_gpu_chain.py -x -q -k "asynchronous or mix" 2>&1 | tail -n 2
python bench.py --steps 50 --warmup 5 --no-variants --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -n 2 gpurun_out/r2l_bench.err
) > gpurun_out/r2j_team.txt 2>&1
