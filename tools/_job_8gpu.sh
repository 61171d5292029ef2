cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-sweep) > gpurun_out/bench_8gpu_r2a.json 2> gpurun_out/bench_8gpu_r2a.err; tail -c 2600 gpurun_out/bench_8gpu_r2a.json; grep -v "OpenBLAS\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_8gpu_r2a.err | tail -n 5
(CTB_NO_MULTICAST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-sweep --no-cpu-baseline --no-e2e) > gpurun_out/bench_8gpu_uc_r2a.json 2> gpurun_out/bench_8gpu_uc_r2a.err; tail -c 1500 gpurun_out/bench_8gpu_uc_r2a.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-sweep --no-cpu-baseline) > gpurun_out/bench_4gpu_r2a.json 2> gpurun_out/bench_4gpu_r2a.err; tail -c 1800 gpurun_out/bench_4gpu_r2a.json
