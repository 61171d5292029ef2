cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(CTB_BENCH_VERBOSE=1 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/bench_2gpu_r2d.json 2> gpurun_out/bench_2gpu_r2d.err
grep -v "OpenBLAS\|OMP_NUM\|\*\*\*\*\|^$" gpurun_out/bench_2gpu_r2d.err | head -n 30
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu_r2d.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'parity', d.get('parity_checked'), d.get('parity_rel_err_vs_reference'))
for s in d.get('sweep', []):
    b=s['b200']; print(s['config'], b['s_per_sweep'], b['energies'], {k:(round(v,2) if isinstance(v,float) else v) for k,v in b['phases_s'].items()})
PY
