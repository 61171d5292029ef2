# final validation, one GPU: full GPU suite, default bench line, reference arm, ncu launch list of the bench command, full capture of the Heff kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v OpenBLAS | tail -6) > gpurun_out/pytest_gpu_final.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_final.log
(time timeout 1500 python bench.py) > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'launches', d['gpu_launches'], d['clocks'])
for s in d['sweep']:
    for key in ('b200', 'b200_merged_pair_tensor', 'b200_pair_form'):
        b = s.get(key)
        if b: print(s['config'], key, {k:(round(v,3) if isinstance(v,float) else v) for k,v in b.items() if k not in ('phases_s','bond_dims','mpo_bond_dims')})
PY
grep real gpurun_out/bench_default_final.err
(timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.json 2> /dev/null); tail -c 400 gpurun_out/bench_ref_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep > gpurun_out/ncu_launches_final.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"grouped_gemm_kernel|mix_kernel" --launch-skip 9 -c 3 -f -o gpurun_out/heff_full_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep > gpurun_out/ncu_full_final.log 2>&1; tail -2 gpurun_out/ncu_full_final.log | cut -c1-200
