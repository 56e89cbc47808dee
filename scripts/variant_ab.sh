#!/bin/bash
# A/B of experiment variants on the GPU box (run under gpurun).  Everything is bounded by `timeout`;
# output under gpurun_out/ab/.  Order = priority (the call may be cut short by the GPU budget).
set -u
O=gpurun_out/ab; mkdir -p $O
V=madrl_b200/variants
timeout 420 python -m pytest tests -m gpu -x -q > $O/tests_default.log 2>&1; echo "tests_default rc=$?" | tee -a $O/summary.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.log
timeout 240 python bench.py > $O/bench_waterworld.json 2> $O/bench_waterworld.err; echo "bench rc=$?" | tee -a $O/summary.log
for v in "$@"; do
  MADRL_B200_LIB=$V/libmadrl_b200_$v.so timeout 300 python -m pytest tests/test_waterworld_gpu.py tests/test_edge_sizes_gpu.py tests/test_api_gpu.py -m gpu -x -q > $O/tests_$v.log 2>&1
  echo "tests_$v rc=$?" | tee -a $O/summary.log
done
timeout 100 python scripts/ww_sweep.py quick > $O/sweep_default.log 2>&1
for v in "$@"; do
  MADRL_B200_LIB=$V/libmadrl_b200_$v.so timeout 100 python scripts/ww_sweep.py quick > $O/sweep_$v.log 2>&1
done
timeout 100 python scripts/hw_sweep.py > $O/sweep_hostage.log 2>&1
timeout 200 python bench.py --workload hostage --no-cpu > $O/bench_hostage.json 2> $O/bench_hostage.err
timeout 200 python bench.py --workload waterworld_c4 --no-cpu > $O/bench_c4.json 2> $O/bench_c4.err
tail -n 3 $O/tests_*.log; cat $O/sweep_*.log; cat $O/bench_waterworld.json
