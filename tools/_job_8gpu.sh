cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(CTB_BENCH_VERBOSE=1 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/bench_8gpu_r2b.json 2> gpurun_out/bench_8gpu_r2b.err
grep -v "OpenBLAS\|OMP_NUM\|\*\*\*\*\|^$" gpurun_out/bench_8gpu_r2b.err | grep "rank 0\|Error\|error" | head -n 12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_8gpu_r2b.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_step_ms'], 'e2e', d['e2e']['ms_per_step'], 'parity', d.get('parity_checked'), d.get('parity_rel_err_vs_reference'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for s in d.get('sweep', []):
    b=s['b200']; print(s['config'], b['s_per_sweep'], b['energies'], {k:(round(v,2) if isinstance(v,float) else v) for k,v in b['phases_s'].items()})
PY
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-sweep) > gpurun_out/bench_4gpu_r2b.json 2> gpurun_out/bench_4gpu_r2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_4gpu_r2b.json').read().strip().splitlines()[-1])
print('N=4 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
PY
