#!/bin/bash
# ncu evidence for one library build (run under gpurun, ONE GPU):
#   gpurun --timeout 900 -- 'bash scripts/ncu_capture.sh r2 [madrl_b200/variants/libmadrl_b200_x.so]'
# Writes gpurun_out/<tag>_{ww_c2,ww_c4,pe,hw}.ncu-rep (--set full, one launch of each env kernel at its
# bench shape) and gpurun_out/<tag>_launches.csv (launch list of a short bench.py run).  Summaries for
# profiles/ are made here afterwards with scripts/ncu_summary.py.
set -u
TAG=${1:-r2}; LIB=${2:-}
[ -n "$LIB" ] && export MADRL_B200_LIB=$LIB
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --kernel-name-base demangled -c 1"
timeout 250 ncu $COMMON -s 3 -k 'regex:ww_kernel<float, .int.1,' -o gpurun_out/${TAG}_ww_c2 -f python scripts/ww_sweep.py c2one > gpurun_out/${TAG}_ncu_ww_c2.log 2>&1
timeout 250 ncu $COMMON -s 3 -k 'regex:ww_kernel<float, .int.4,' -o gpurun_out/${TAG}_ww_c4 -f python scripts/ww_sweep.py c4one > gpurun_out/${TAG}_ncu_ww_c4.log 2>&1
timeout 250 ncu $COMMON -s 4 -k regex:pe_kernel -o gpurun_out/${TAG}_pe -f python scripts/pe_sweep.py quick > gpurun_out/${TAG}_ncu_pe.log 2>&1
timeout 250 ncu $COMMON -s 4 -k regex:hw_kernel -o gpurun_out/${TAG}_hw -f python scripts/hw_sweep.py quick > gpurun_out/${TAG}_ncu_hw.log 2>&1
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ls -la gpurun_out/${TAG}_*
