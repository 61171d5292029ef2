cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BIG="2168,2166;1468,1460;1444,1450;1418,1396;1390,1378;988,1016;982,932;928,936;938,926;404,390;388,394"
SMALL="560,540;500,520;480,470;300,310;290,280;200,210;120,130"
(timeout 300 python tools/svd_check.py both "200,200;300,260;150,400;97,33" 12) > gpurun_out/svd_check_a.log 2>&1; tail -n 4 gpurun_out/svd_check_a.log
(CTB_SVD_PROFILE=1 timeout 600 python tools/svd_check.py real "$BIG" 14) > gpurun_out/svd_prof_d.log 2>&1; grep -v cycle gpurun_out/svd_prof_d.log | tail -n 4
(CTB_SVD_PROFILE=1 timeout 600 python tools/svd_check.py both "1024,1024" 14) > gpurun_out/svd_prof_1k.log 2>&1; grep -v cycle gpurun_out/svd_prof_1k.log | tail
(CTB_SVD_PROFILE=1 timeout 600 python tools/svd_check.py real "$SMALL" 14) > gpurun_out/svd_prof_s.log 2>&1; grep -v cycle gpurun_out/svd_prof_s.log | tail -n 4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/svd_launches_small.csv python tools/svd_check.py real "$SMALL" 14 > gpurun_out/svd_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/svd_launches.csv python tools/svd_check.py real "$BIG" 14 > gpurun_out/svd_ncu.log 2>&1
(timeout 600 python tools/sweep_run.py fermi_hubbard 32 1024) > gpurun_out/sweep_fh32_r2c.json 2> gpurun_out/sweep_fh32_r2c.err; tail -c 700 gpurun_out/sweep_fh32_r2c.json
(timeout 600 python tools/sweep_run.py xxz 100 1024 1) > gpurun_out/sweep_xxz_r2c.json 2> gpurun_out/sweep_xxz_r2c.err; tail -c 700 gpurun_out/sweep_xxz_r2c.json
