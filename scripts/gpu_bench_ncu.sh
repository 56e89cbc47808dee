#!/bin/bash
# bench.py (default run) + the Waterworld ncu captures on a 1-GPU box.
set -u
TAG=${1:-r2b}; O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 $O/${TAG}_bench.err
COMMON="--set full --clock-control none --import-source on --kernel-name-base demangled -c 1"
timeout 250 ncu $COMMON -s 3 -k 'regex:ww_kernel<float, .int.1,' -o $O/${TAG}_ww_c2 -f python scripts/ww_sweep.py c2one > $O/${TAG}_ncu_ww_c2.log 2>&1
timeout 250 ncu $COMMON -s 3 -k 'regex:ww_kernel<float, .int.4,' -o $O/${TAG}_ww_c4 -f python scripts/ww_sweep.py c4one > $O/${TAG}_ncu_ww_c4.log 2>&1
ls -la $O/${TAG}_*
python -c "
import json;d=json.loads(open('$O/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']); print(d['clocks']); print(d.get('e2e')); print(d.get('cpu_baseline'))
for k,v in d.get('workloads',{}).items(): print(k, v['value'], v['roofline']['frac'], v.get('e2e',{}).get('value'))
"
