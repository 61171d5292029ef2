cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest_gpu_r1x.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_r1x.log
(timeout 600 python bench.py --sweep > gpurun_out/bench_default_r1x.json 2> gpurun_out/bench_default_r1x.err); tail -c 600 gpurun_out/bench_default_r1x.json
(timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1x.json 2> /dev/null); tail -c 300 gpurun_out/bench_ref_r1x.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1x.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_r1x.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"grouped_gemm_kernel|mix_kernel" --launch-skip 9 -c 3 -f -o gpurun_out/heff_full_r1x python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r1x.log 2>&1
ls -la gpurun_out/*.ncu-rep
