cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
(timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q -s 2>&1 | tail -15) > gpurun_out/pytest_gpu2_r2a.log 2>&1; tail -n 15 gpurun_out/pytest_gpu2_r2a.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-sweep) > gpurun_out/bench_2gpu_r2a.json 2> gpurun_out/bench_2gpu_r2a.err; tail -c 2500 gpurun_out/bench_2gpu_r2a.json; tail -n 5 gpurun_out/bench_2gpu_r2a.err
(CTB_NO_MULTICAST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-sweep --no-cpu-baseline) > gpurun_out/bench_2gpu_uc_r2a.json 2> gpurun_out/bench_2gpu_uc_r2a.err; tail -c 1200 gpurun_out/bench_2gpu_uc_r2a.json
