cd $GRAFT_REPO_ROOT
timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, helpers
from chemtensor_b200 import workloads
lib = helpers.load("cuda")
for name, model, L, params, sector, D, _ in bench.SWEEP_CASES:
    print(name, json.dumps(bench.sweep_seconds(lib, model, L, params, sector, D)), flush=True)
PY
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu_r1ad.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_r1ad.log
