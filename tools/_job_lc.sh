#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 90 python - > gpurun_out/lc_direct_r2t.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "tests")
import helpers, test_zz_blocklc as t
eng = helpers.load("cuda")
t.test_contiguous_blocks_all_lengths_and_parities(eng, False); print("contiguous f64 ok")
t.test_contiguous_blocks_all_lengths_and_parities(eng, True); print("contiguous c128 ok")
t.test_permutations_and_stacking_are_exact(eng); print("permutations ok")
t.test_in_place_scaling(eng); print("in-place ok")
PY
cat gpurun_out/lc_direct_r2t.log | tail -6
