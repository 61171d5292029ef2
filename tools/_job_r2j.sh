# job F: re-blocking kernel with four runs in flight; complex128 molecular sweeps at 16 and 24 orbitals
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "not dmrg and not sweep and not fullsize and not config" 2>&1 | grep -v OpenBLAS | tail -5) > gpurun_out/pytest_gpu_r2j.log 2>&1; tail -n 2 gpurun_out/pytest_gpu_r2j.log
(timeout 300 python tools/level1_bench.py > gpurun_out/level1_r2j.jsonl 2> gpurun_out/level1_r2j.err); grep remap gpurun_out/level1_r2j.jsonl | cut -c1-330
(time timeout 300 python tools/molecular_run.py 16 512 1 > gpurun_out/mol16_r2j.json 2> gpurun_out/mol16_r2j.err); tail -c 1500 gpurun_out/mol16_r2j.json; tail -n 6 gpurun_out/mol16_r2j.err
(time timeout 500 python tools/molecular_run.py 24 1024 1 > gpurun_out/mol24_r2j.json 2> gpurun_out/mol24_r2j.err); tail -c 1500 gpurun_out/mol24_r2j.json; tail -n 6 gpurun_out/mol24_r2j.err
nvidia-smi --query-gpu=memory.used --format=csv
