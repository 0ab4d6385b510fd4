(python -m pytest tests/test_trace.py -x -q -m gpu 2>&1 | tail -n 4) > gpurun_out/r2j_team.txt 2>&1
