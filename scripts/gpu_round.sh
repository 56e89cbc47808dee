#!/bin/bash
# One full evidence pass on a 1-GPU box: GPU parity suite, smoke, throughput sweeps, ncu captures.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_round.sh r2'
set -u
TAG=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -x > $O/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -n 1 $O/${TAG}_tests.log)"
timeout 120 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$? $(tail -n 1 $O/${TAG}_smoke.log)"
for f in ww pe hw; do timeout 120 python scripts/${f}_sweep.py quick 2>&1 | tee $O/${TAG}_sweep_$f.log; done
bash scripts/ncu_capture.sh $TAG > $O/${TAG}_ncu_capture.log 2>&1
tail -n 8 $O/${TAG}_ncu_capture.log
