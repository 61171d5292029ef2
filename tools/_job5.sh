cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(CTB_TRACE_PLAN=1 timeout 900 python tools/sweep_run.py fermi_hubbard 64 4096 1) > gpurun_out/sweep_fh64_r2d.json 2> gpurun_out/sweep_fh64_r2d.err; tail -c 600 gpurun_out/sweep_fh64_r2d.json; grep "contraction plans" gpurun_out/sweep_fh64_r2d.err
