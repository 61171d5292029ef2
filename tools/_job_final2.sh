cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1500 python bench.py) > gpurun_out/bench_default_final2.json 2> gpurun_out/bench_default_final2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_final2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'launches', d['gpu_launches'], d['clocks'])
for s in d['sweep']:
    for key in ('b200', 'b200_merged_pair_tensor', 'b200_pair_form'):
        b = s.get(key)
        if b: print(s['config'], key, {k:(round(v,3) if isinstance(v,float) else v) for k,v in b.items() if k not in ('phases_s','bond_dims','mpo_bond_dims')}, (b.get('phases_s') or {}).get('per_sweep_s'))
PY
grep real gpurun_out/bench_default_final2.err
