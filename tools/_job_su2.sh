#!/bin/bash
# GPU job of the SU(2) half of the round: parity tests, lc-kernel roofline, config-5 sweeps, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_zz_su2.py -m gpu -x -q > gpurun_out/pytest_su2_r2q.log 2>&1; tail -2 gpurun_out/pytest_su2_r2q.log
timeout 60 python tools/su2_lc_bench.py > gpurun_out/su2_lc_bench_r2q.jsonl 2>&1; cat gpurun_out/su2_lc_bench_r2q.jsonl | cut -c1-220
timeout 240 python tools/su2_run.py cuda 200 8192 --sweeps 5 --lanczos 10 --degen 8 --out gpurun_out/su2_L200_D8192_r2q.json > gpurun_out/su2_L200_D8192_r2q.log 2>&1; grep "sweep" gpurun_out/su2_L200_D8192_r2q.log | cut -c1-330; tail -c 400 gpurun_out/su2_L200_D8192_r2q.log
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/su2_launches_r2q.csv python tools/su2_run.py cuda 32 512 --sweeps 2 --lanczos 10 --degen 8 > gpurun_out/su2_ncu_r2q.log 2>&1; tail -2 gpurun_out/su2_ncu_r2q.log | cut -c1-200; wc -l gpurun_out/su2_launches_r2q.csv
