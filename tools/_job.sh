cd $GRAFT_REPO_ROOT
nvidia-smi -L
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_r1d.log 2>&1; tail -n 5 gpurun_out/pytest_gpu_r1d.log
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conv_r1d.log 2>&1); tail -n 2 gpurun_out/bench_conv_r1d.log | cut -c1-2500
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_conv_2gpu_r1d.log 2>&1); tail -n 3 gpurun_out/bench_conv_2gpu_r1d.log | cut -c1-2500
(timeout 300 python tools/trace_e2e.py fh_L64_D4096 > gpurun_out/trace_e2e_r1d.log 2>&1); tail -n 6 gpurun_out/trace_e2e_r1d.log
