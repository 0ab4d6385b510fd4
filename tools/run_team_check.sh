(python -m pytest tests/test_gpu_chain.py tests/test_gpu_fullsize.py -x -q -k "fdn or config4 or config_4" 2>&1 | tail -n 2
python tools/bench_configs.py --only 4 2>&1 | cut -c1-260
MLB_TEAM_PROF=0 python - <<'PY'
import sys; sys.path.insert(0,'.')
from madronalib_b200 import api
import ctypes
PY
) > gpurun_out/r2j_team.txt 2>&1
