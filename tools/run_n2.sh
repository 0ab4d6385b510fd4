python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2k_multi.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench2.json 2> gpurun_out/r2k_bench2.err; echo "rc=$?" >> gpurun_out/r2k_bench2.err
