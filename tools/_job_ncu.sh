cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SMALL="560,540;500,520;480,470;300,310;290,280;200,210;120,130"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bj_eig_kernel" --launch-skip 20 -c 1 -f -o gpurun_out/bj_eig_small python tools/svd_check.py real "$SMALL" 14 > gpurun_out/ncu_eig.log 2>&1; tail -n 2 gpurun_out/ncu_eig.log
