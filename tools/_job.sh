cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_r1f.log 2>&1; tail -n 5 gpurun_out/pytest_gpu_r1f.log
(timeout 300 python tools/gemm_sweep.py quick > gpurun_out/gemm_sweep_r1f.jsonl 2> gpurun_out/gemm_sweep_r1f.err); cat gpurun_out/gemm_sweep_r1f.jsonl; tail -n 3 gpurun_out/gemm_sweep_r1f.err
(timeout 900 python bench.py --steps 10 --warmup 3 --sweep > gpurun_out/bench_conv_r1f.log 2>&1); tail -n 2 gpurun_out/bench_conv_r1f.log | cut -c1-4000
