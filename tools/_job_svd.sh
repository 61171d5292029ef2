# GPU job: accuracy + timing of the block-Jacobi SVD, then the factorization / golden / DMRG GPU tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python tools/svd_check.py real "200,200;300,260;150,400;97,33" 12 --old) > gpurun_out/svd_check_a.log 2>&1; tail -n 6 gpurun_out/svd_check_a.log
(timeout 300 python tools/svd_check.py complex "200,200;300,260;150,400;97,33" 12 --old) > gpurun_out/svd_check_b.log 2>&1; tail -n 6 gpurun_out/svd_check_b.log
(timeout 600 python tools/svd_check.py real "1024,1024;700,900;512,512;512,512" 14 --old) > gpurun_out/svd_check_c.log 2>&1; tail -n 6 gpurun_out/svd_check_c.log
(timeout 600 python tools/svd_check.py real "2168,2166;1468,1460;1444,1450;1418,1396;1390,1378;988,1016;982,932;928,936;938,926;404,390;388,394" 14) > gpurun_out/svd_check_d.log 2>&1; tail -n 6 gpurun_out/svd_check_d.log
(timeout 900 python -m pytest tests/test_factorizations.py tests/test_golden_engine.py tests/test_dmrg.py tests/test_zz_golden_more.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_svd.log 2>&1; tail -n 15 gpurun_out/pytest_svd.log
