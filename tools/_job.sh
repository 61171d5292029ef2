# Template of a GPU job (run with: gpurun --timeout 1500 -- 'bash tools/_job.sh'): the full GPU suite, the default bench line, the
# reference arm, the ncu launch list and one full capture of the three Heff kernels.  Outputs land in gpurun_out/.
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
(timeout 600 python bench.py --sweep > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err); tail -c 600 gpurun_out/bench_default.json
(timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> /dev/null); tail -c 300 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"grouped_gemm_kernel|mix_kernel" --launch-skip 9 -c 3 -f -o gpurun_out/heff_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
# candidates measured at the end of round 1: CTB_SVD_BLOCK_ROWS=8 python -m pytest tests -m gpu -q ; then bench.py --sweep
