python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2i_multi8.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_multi8.log; tail -n 5 gpurun_out/r2i_multi8.log
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2i_bench$N.json 2> gpurun_out/r2i_bench$N.err; echo "bench $N rc=$?"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench1.json 2> gpurun_out/r2i_bench1.err; echo "bench 1 rc=$?"
