for D in 0; do
MLB_BUS_DEBUG=$D python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 50 --warmup 5 --no-parity > gpurun_out/r2k_bench2_d$D.json 2> gpurun_out/r2k_bench2.err
done
