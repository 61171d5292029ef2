cd $GRAFT_REPO_ROOT
CTB_SVD_BLOCK_ROWS=8 timeout 25 python - <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, helpers
from chemtensor_b200 import workloads
lib = helpers.load("cuda")
name, model, L, params, sector, D, _ = bench.SWEEP_CASES[1]
r = bench.sweep_seconds(lib, model, L, params, sector, D)
print("rows8", name, r["s_per_sweep"], r["energies"], r["phases_s"]["svd_split"], r["phases_s"]["lanczos_incl_plans"], flush=True)
PY
