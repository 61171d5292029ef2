cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_fullsize.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_fullsize_r1y.log 2>&1; tail -n 6 gpurun_out/pytest_fullsize_r1y.log
