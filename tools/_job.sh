cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu2_r1u.log 2>&1; tail -n 3 gpurun_out/pytest_gpu2_r1u.log
run() { # name, nproc, env...
  name=$1; n=$2; shift 2
  (env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_${name}.log 2>&1)
  grep '"metric"' gpurun_out/bench_${name}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],4), [round(x,4) for x in d['roofline']['per_step_ms']], round(d['roofline']['exchange_ms'],4))" || tail -n 5 gpurun_out/bench_${name}.log
}
run r1u_2gpu_fused 2 CTB_EXCHANGE=fused
run r1u_2gpu_fused_ncclbar 2 CTB_EXCHANGE=fused CTB_NCCL_BARRIER=1
