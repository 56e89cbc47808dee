#!/bin/bash
# A/B of experiment builds on one B200 (run under gpurun, after scripts/build_variants.sh here):
#   gpurun --timeout 900 -- 'bash scripts/variant_ab.sh w1:ww,pe,hw split:hw'
# For the default library and every named variant (madrl_b200/variants/libmadrl_b200_<name>.so):
# the GPU parity tests of the listed kernel families, then their throughput sweeps.  Everything is
# bounded by `timeout`; output under gpurun_out/ab/ (summary.log first).
set -u
O=gpurun_out/ab; mkdir -p $O; : > $O/summary.log
V=madrl_b200/variants
declare -A TESTS=( [ww]=tests/test_waterworld_gpu.py [pe]=tests/test_pursuit_gpu.py [hw]=tests/test_hostage_gpu.py )
run() {  # name, lib-or-empty, families
  local n=$1 lib=$2 fams=$3
  ( [ -n "$lib" ] && export MADRL_B200_LIB=$lib
    local files="tests/test_edge_sizes_gpu.py"
    for f in ${fams//,/ }; do files="$files ${TESTS[$f]}"; done
    timeout 300 python -m pytest $files -m gpu -q -x > $O/tests_$n.log 2>&1; echo "tests_$n rc=$? $(tail -n 1 $O/tests_$n.log)" >> $O/summary.log
    for f in ${fams//,/ }; do
      timeout 90 python scripts/${f}_sweep.py quick > $O/sweep_${f}_$n.log 2>&1
      sed "s/^/$n: /" $O/sweep_${f}_$n.log >> $O/summary.log
    done )
}
run default "" ${DEFAULT_FAMS:-ww,pe,hw}
for spec in "$@"; do
  v=${spec%%:*}; fams=${spec#*:}; [ "$fams" = "$spec" ] && fams=ww,pe,hw
  run $v $V/libmadrl_b200_$v.so $fams
done
cat $O/summary.log
