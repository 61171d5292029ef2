cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BIG="2168,2166;1468,1460;1444,1450;1418,1396;1390,1378;988,1016;982,932;928,936;938,926;404,390;388,394"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bj_eig_kernel" --launch-skip 70 -c 1 -f -o gpurun_out/bj_eig_full python tools/svd_check.py real "$BIG" 14 > gpurun_out/ncu_eig.log 2>&1; tail -n 2 gpurun_out/ncu_eig.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bj_update_kernel|bj_gram_kernel" --launch-skip 140 -c 2 -f -o gpurun_out/bj_gemm_full python tools/svd_check.py real "$BIG" 14 > gpurun_out/ncu_gemm.log 2>&1; tail -n 2 gpurun_out/ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bj_qr_panel_kernel|bj_qr_apply_kernel" --launch-skip 4 -c 2 -f -o gpurun_out/bj_qr_full python tools/svd_check.py real "$BIG" 14 > gpurun_out/ncu_qr.log 2>&1; tail -n 2 gpurun_out/ncu_qr.log
ls -la gpurun_out/*.ncu-rep
