(python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5
(time python bench.py) > gpurun_out/bench_final_r2.json 2> gpurun_out/bench_final_r2.err; tail -n 4 gpurun_out/bench_final_r2.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_final_r2_reference.json 2>/dev/null
python tools/bench_configs.py > gpurun_out/configs_r2.jsonl 2> gpurun_out/configs_r2.err
) > gpurun_out/final_r2.txt 2>&1
