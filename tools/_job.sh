cd $GRAFT_REPO_ROOT
(timeout 100 python -m pytest tests/test_zz_svd_workspace_gpu.py -m gpu -x -q 2>&1 | tail -12)
