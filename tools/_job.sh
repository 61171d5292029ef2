cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest_gpu_r1p.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_r1p.log
timeout 300 python tools/trace_e2e.py fh_L64_D4096 2>&1 | tail -n 6
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['per_step_ms'], d['e2e'])"
