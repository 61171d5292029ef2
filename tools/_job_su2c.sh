#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_zz_su2.py -m gpu -x -q > gpurun_out/pytest_su2_r2s.log 2>&1; tail -2 gpurun_out/pytest_su2_r2s.log
CTB_BENCH_SWEEP_ONLY=su2 timeout 215 python bench.py > gpurun_out/bench_su2_r2s.json 2> gpurun_out/bench_su2_r2s.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_su2_r2s.json").read().strip().splitlines()[-1])
    print("value", d.get("value"), "e2e", d.get("e2e", {}).get("value"))
    for rec in d.get("sweep", []):
        b = rec.get("b200") or {}
        print(rec["config"], b.get("sweep_s"), b.get("energies"), rec.get("max_energy_diff_vs_reference"), b.get("max_multiplets"), (b.get("stats") or {}))
except Exception as exc:
    print("bench line not parsed:", exc)
PY
tail -3 gpurun_out/bench_su2_r2s.err | cut -c1-300
