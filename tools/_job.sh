cd $GRAFT_REPO_ROOT
(timeout 500 python bench.py --sweep --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_sweep_r1ac.json 2> gpurun_out/bench_sweep_r1ac.err); python -c "
import json
d=json.loads(open('gpurun_out/bench_sweep_r1ac.json').read().strip().splitlines()[-1])
print(json.dumps(d['sweep'], indent=1))
"
tail -n 3 gpurun_out/bench_sweep_r1ac.err
