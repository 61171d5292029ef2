cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SMALL="560,540;500,520;480,470;300,310;290,280;200,210;120,130"
(CTB_SVD_EIG_TIMING=1 CTB_SVD_PROFILE=1 timeout 600 python tools/svd_check.py real "$SMALL" 14) > gpurun_out/svd_eig_timing.log 2>&1; grep -v cycle gpurun_out/svd_eig_timing.log | tail -n 8
