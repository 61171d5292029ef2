# job G: plan-building profile at scale after the packed-matrix signature fix
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(CTB_TRACE_PLAN=1 timeout 300 python tools/molecular_run.py 16 512 1 > gpurun_out/mol16_r2k.json 2> gpurun_out/mol16_r2k.err); python -c "
import json; d=json.loads(open('gpurun_out/mol16_r2k.json').read().strip().splitlines()[-1]); print(d['s_per_sweep'], d['energies'], d['phases_s'])"; grep -A3 "contraction plans" gpurun_out/mol16_r2k.err
(CTB_TRACE_PLAN=1 timeout 300 python tools/sweep_run.py fermi_hubbard 32 1024 2 10 > gpurun_out/fh32_r2k.json 2> gpurun_out/fh32_r2k.err); cut -c1-900 gpurun_out/fh32_r2k.json; grep -A3 "contraction plans" gpurun_out/fh32_r2k.err
(CTB_TRACE_PLAN=1 timeout 500 python tools/molecular_run.py 24 1024 1 > gpurun_out/mol24_r2k.json 2> gpurun_out/mol24_r2k.err); python -c "
import json; d=json.loads(open('gpurun_out/mol24_r2k.json').read().strip().splitlines()[-1]); print(d['s_per_sweep'], d['energies'], d['phases_s'])"; grep -A3 "contraction plans" gpurun_out/mol24_r2k.err
