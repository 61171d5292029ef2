cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_r1j.log 2>&1; tail -n 5 gpurun_out/pytest_gpu_r1j.log
for c in 0 6; do
  echo "== class $c"; CTB_GEMM_CLASS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['per_step_ms'], d['roofline']['per_step_tflops'])"
done
for n in 2; do
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_conv_${n}gpu_r1j.log 2>&1); tail -n 3 gpurun_out/bench_conv_${n}gpu_r1j.log | grep '"metric"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['per_step_ms'], d['roofline']['exchange_ms'])"
done
tail -n 5 gpurun_out/bench_conv_2gpu_r1j.log | cut -c1-400
