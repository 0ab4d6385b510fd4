# one GPU call: the contract-E tests, the C++ host API test on the device, then the default bench line
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_cpp_hostapi.py tests/test_voice_bank.py -m gpu -x -q 2>&1 | tail -n 15
(time timeout 600 python bench.py) > gpurun_out/bench_synth_r2.json 2> gpurun_out/bench_synth_r2.err; tail -n 4 gpurun_out/bench_synth_r2.err
) > gpurun_out/synth_r2.txt 2>&1
tail -n 30 gpurun_out/synth_r2.txt
