cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_pair_form.py tests/test_dmrg.py tests/test_callers.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_pair_r1aa.log 2>&1; tail -n 6 gpurun_out/pytest_pair_r1aa.log
