cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mode in hash dense; do
  if [ $mode = dense ]; then export CTB_GRID_DENSE_MAX=100000000000; else unset CTB_GRID_DENSE_MAX; fi
  (CTB_TRACE_PLAN=1 timeout 400 python tools/sweep_run.py fermi_hubbard 64 4096 1 10 > gpurun_out/fh64_ab_$mode.json 2> gpurun_out/fh64_ab_$mode.err)
  python - <<PY
import json
d=json.loads(open('gpurun_out/fh64_ab_$mode.json').read().strip().splitlines()[-1])
print('$mode', d['s_per_sweep'], d['energies'], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['phases_s'].items() if k!='per_sweep_s'})
PY
  grep -A1 "contraction plans" gpurun_out/fh64_ab_$mode.err | cut -c1-250
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
