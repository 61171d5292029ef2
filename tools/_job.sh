cd $GRAFT_REPO_ROOT
for c in 7 0; do
  echo "== class $c"; CTB_GEMM_CLASS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['per_step_ms'], d['roofline']['per_step_tflops'])"
done
(timeout 300 python -m pytest tests/test_tensor_ops.py -m gpu -x -q 2>&1 | tail -2)
