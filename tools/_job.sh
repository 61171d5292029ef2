cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_r1e.log 2>&1; tail -n 5 gpurun_out/pytest_gpu_r1e.log
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conv_r1e.log 2>&1); tail -n 2 gpurun_out/bench_conv_r1e.log | cut -c1-2600
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_conv_2gpu_r1e.log 2>&1); tail -n 3 gpurun_out/bench_conv_2gpu_r1e.log | cut -c1-2600
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_r1e.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grouped_gemm|mix_kernel" -s 9 -c 3 -o gpurun_out/prof_heff_r1e python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu2_r1e.log 2>&1)
ls -la gpurun_out/
