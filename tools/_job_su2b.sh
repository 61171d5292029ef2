#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 230 python tools/su2_run.py cuda 200 8192 --sweeps 4 --lanczos 10 --degen 8 --out gpurun_out/su2_L200_D8192_r2r.json > gpurun_out/su2_L200_D8192_r2r.log 2>&1; cut -c1-900 gpurun_out/su2_L200_D8192_r2r.log | tail -12
timeout 150 python -m pytest tests/test_factorizations.py tests/test_sweep_parity.py tests/test_zz_su2.py -m gpu -x -q > gpurun_out/pytest_svd_su2_r2r.log 2>&1; tail -3 gpurun_out/pytest_svd_su2_r2r.log
