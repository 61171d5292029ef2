cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu_r2a.log 2>&1; tail -n 8 gpurun_out/pytest_gpu_r2a.log
(timeout 300 python tools/sweep_run.py fermi_hubbard 16 256) > gpurun_out/sweep_fh16_r2a.json 2> gpurun_out/sweep_fh16_r2a.err; tail -c 900 gpurun_out/sweep_fh16_r2a.json
(timeout 600 python tools/sweep_run.py fermi_hubbard 32 1024) > gpurun_out/sweep_fh32_r2a.json 2> gpurun_out/sweep_fh32_r2a.err; tail -c 900 gpurun_out/sweep_fh32_r2a.json
(timeout 900 python tools/sweep_run.py fermi_hubbard 64 4096 1) > gpurun_out/sweep_fh64_r2a.json 2> gpurun_out/sweep_fh64_r2a.err; tail -c 900 gpurun_out/sweep_fh64_r2a.json; tail -n 3 gpurun_out/sweep_fh64_r2a.err
