cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | grep -v OpenBLAS | tail -8) > gpurun_out/pytest_gpu_r2b.log 2>&1; tail -n 8 gpurun_out/pytest_gpu_r2b.log
(time timeout 1500 python bench.py) > gpurun_out/bench_default_r2b.json 2> gpurun_out/bench_default_r2b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_r2b.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for s in d['sweep']:
    b=s.get('b200') or s.get('b200_merged_pair_tensor')
    print(s['config'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in b.items() if k not in ('phases_s','bond_dims','mpo_bond_dims')})
    if b.get('phases_s'): print('   ', {k:(round(v,3) if isinstance(v,float) else v) for k,v in b['phases_s'].items()})
PY
grep real gpurun_out/bench_default_r2b.err
