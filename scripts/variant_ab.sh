#!/bin/bash
# A/B of experiment builds on one B200 (run under gpurun, after scripts/build_variants.sh here):
#   gpurun --timeout 900 -- 'bash scripts/variant_ab.sh w1 all'
# For the default library and every named variant (madrl_b200/variants/libmadrl_b200_<name>.so):
# the GPU parity suite, then the three throughput sweeps.  Everything is bounded by `timeout`;
# output under gpurun_out/ab/ (summary.log first).
set -u
O=gpurun_out/ab; mkdir -p $O; : > $O/summary.log
V=madrl_b200/variants
run() {  # name, lib-or-empty
  local n=$1 lib=$2
  ( [ -n "$lib" ] && export MADRL_B200_LIB=$lib
    timeout 300 python -m pytest tests -m gpu -q -x > $O/tests_$n.log 2>&1; echo "tests_$n rc=$? $(tail -n 1 $O/tests_$n.log)" >> $O/summary.log
    timeout 90 python scripts/ww_sweep.py quick > $O/sweep_ww_$n.log 2>&1
    timeout 90 python scripts/pe_sweep.py quick > $O/sweep_pe_$n.log 2>&1
    timeout 90 python scripts/hw_sweep.py quick > $O/sweep_hw_$n.log 2>&1
    for f in ww pe hw; do sed "s/^/$n: /" $O/sweep_${f}_$n.log >> $O/summary.log; done )
}
run default ""
for v in "$@"; do run $v $V/libmadrl_b200_$v.so; done
cat $O/summary.log
