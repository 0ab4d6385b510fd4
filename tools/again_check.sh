# the MLB_AGAIN GPU tests only (what is left of the round's GPU budget)
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_zz_gpu_again.py -m gpu -q -p no:cacheprovider > gpurun_out/again_r2.txt 2>&1
tail -n 15 gpurun_out/again_r2.txt
