# One GPU call at the end of a round: the whole -m gpu suite, smoke(), the default bench line and the
# per-config timings, all at HEAD.  Results under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
(python -m pytest tests -m gpu -q 2>&1 | tail -n 6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5
(time python bench.py) > gpurun_out/bench_final_r2b.json 2> gpurun_out/bench_final_r2b.err; tail -n 4 gpurun_out/bench_final_r2b.err
python tools/bench_configs.py > gpurun_out/configs_r2b.jsonl 2> gpurun_out/configs_r2b.err
cut -c1-260 gpurun_out/configs_r2b.jsonl
) > gpurun_out/final_r2b.txt 2>&1
tail -n 40 gpurun_out/final_r2b.txt
