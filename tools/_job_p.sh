cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time CTB_TRACE_PLAN=1 CTB_BENCH_SWEEP_ONLY=mol_n24 timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/bench_n24only.json 2> gpurun_out/bench_n24only.err); python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n24only.json').read().strip().splitlines()[-1])
for s in d['sweep']: print(s['config'], s['b200']['s_per_sweep'], s['b200']['phases_s'])
PY
grep -A3 "contraction plans" gpurun_out/bench_n24only.err; grep real -A2 gpurun_out/bench_n24only.err
(time CTB_BENCH_SWEEP_ONLY=fh_L64 timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/bench_fh64only.json 2> gpurun_out/bench_fh64only.err); python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fh64only.json').read().strip().splitlines()[-1])
for s in d['sweep']: print(s['config'], s['b200'])
PY
tail -3 gpurun_out/bench_fh64only.err
