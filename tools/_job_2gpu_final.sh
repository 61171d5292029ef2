cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | grep -v OpenBLAS | tail -6) > gpurun_out/pytest_gpu2_final.log 2>&1; tail -n 4 gpurun_out/pytest_gpu2_final.log
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/bench_2gpu_final.json 2> gpurun_out/bench_2gpu_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'parity', d.get('parity_checked'), d.get('parity_rel_err_vs_reference'))
for s in d.get('sweep', []):
    b=s['b200']; print(s['config'], b['s_per_sweep'], b['energies'], {k:(round(v,2) if isinstance(v,float) else v) for k,v in b['phases_s'].items()})
PY
