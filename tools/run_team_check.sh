(for S in 8 16 24 48 64; do echo "stages $S"; MLB_STAGES=$S python tools/bench_configs.py --only 5,6 2>&1 | cut -c1-175; done) > gpurun_out/r2j_team.txt 2>&1
