cd $GRAFT_REPO_ROOT
(timeout 200 python -m pytest tests/test_factorizations.py tests/test_golden_engine.py tests/test_dmrg.py -m gpu -x -q 2>&1 | tail -4)
