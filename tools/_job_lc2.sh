#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 30 python -m pytest tests/test_zz_su2.py -m gpu -x -q 2>&1 | tail -1
timeout 20 python - 2>&1 <<'PY' | tail -2
import sys
sys.path.insert(0, "tests")
import helpers, test_zz_blocklc as t
eng = helpers.load("cuda")
t.test_contiguous_blocks_all_lengths_and_parities(eng, False); t.test_contiguous_blocks_all_lengths_and_parities(eng, True)
t.test_permutations_and_stacking_are_exact(eng); t.test_in_place_scaling(eng); print("lc direct ok")
PY
timeout 25 python tools/su2_run.py cuda 200 2048 --sweeps 2 --lanczos 10 --degen 8 --out gpurun_out/su2_oneupload_r2v.json > gpurun_out/su2_oneupload_r2v.log 2>&1; grep -o '"sweep_s": [^]]*]' gpurun_out/su2_oneupload_r2v.log; grep -o '"energies": [^]]*]' gpurun_out/su2_oneupload_r2v.log; grep -o '"local_solve_s": [0-9.]*' gpurun_out/su2_oneupload_r2v.log
