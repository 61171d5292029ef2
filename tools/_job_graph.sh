#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_zz_su2.py -m gpu -x -q > gpurun_out/pytest_su2_graph_r2u.log 2>&1; tail -2 gpurun_out/pytest_su2_graph_r2u.log
timeout 25 python tools/su2_run.py cuda 200 2048 --sweeps 2 --lanczos 10 --degen 8 --out gpurun_out/su2_graph_on_r2u.json > gpurun_out/su2_graph_on_r2u.log 2>&1; grep -o '"sweep_s": [^]]*]' gpurun_out/su2_graph_on_r2u.log; grep -o '"energies": [^]]*]' gpurun_out/su2_graph_on_r2u.log; grep -o '"local_solve_s": [0-9.]*' gpurun_out/su2_graph_on_r2u.log
CTB_SU2_NO_GRAPH=1 timeout 25 python tools/su2_run.py cuda 200 2048 --sweeps 2 --lanczos 10 --degen 8 --out gpurun_out/su2_graph_off_r2u.json > gpurun_out/su2_graph_off_r2u.log 2>&1; grep -o '"sweep_s": [^]]*]' gpurun_out/su2_graph_off_r2u.log; grep -o '"energies": [^]]*]' gpurun_out/su2_graph_off_r2u.log; grep -o '"local_solve_s": [0-9.]*' gpurun_out/su2_graph_off_r2u.log
