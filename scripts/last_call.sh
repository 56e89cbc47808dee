#!/bin/bash
set -u
O=gpurun_out/last; mkdir -p $O
timeout 110 python -m pytest tests -m gpu -q > $O/tests_default.log 2>&1; echo "tests_default rc=$?" | tee -a $O/summary.log
tail -n 4 $O/tests_default.log
timeout 40 python scripts/ww_sweep.py quick > $O/sweep_default.log 2>&1; cat $O/sweep_default.log
MADRL_B200_LIB=madrl_b200/variants/libmadrl_b200_deferred.so timeout 40 python scripts/ww_sweep.py quick > $O/sweep_deferred.log 2>&1; cat $O/sweep_deferred.log
MADRL_B200_LIB=madrl_b200/variants/libmadrl_b200_deferred.so timeout 60 python -m pytest tests/test_api_gpu.py tests/test_waterworld_gpu.py tests/test_hostage_gpu.py -m gpu -q > $O/tests_deferred.log 2>&1; echo "tests_deferred rc=$?" | tee -a $O/summary.log
tail -n 4 $O/tests_deferred.log
