cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu2_r1s.log 2>&1; tail -n 3 gpurun_out/pytest_gpu2_r1s.log
