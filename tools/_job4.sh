cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_configs.py tests/test_sweep_parity.py tests/test_factorizations.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/pytest_cfg_r2.log 2>&1; grep -v OpenBLAS gpurun_out/pytest_cfg_r2.log | tail -n 12
(time timeout 1500 python bench.py) > gpurun_out/bench_default_r2a.json 2> gpurun_out/bench_default_r2a.err; tail -c 7000 gpurun_out/bench_default_r2a.json; grep -v OpenBLAS gpurun_out/bench_default_r2a.err | tail -n 5
