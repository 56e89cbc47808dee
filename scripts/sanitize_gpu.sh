#!/bin/bash
# compute-sanitizer passes over small invocations of the three env kernels (run under gpurun):
#   gpurun --timeout 600 -- 'bash scripts/sanitize_gpu.sh'
# memcheck: out-of-bounds / misaligned global and shared accesses; racecheck: shared-memory hazards
# (Pursuit's cell words, the staged candidates of MADRL_WW_SMEM_MIN_OPL); synccheck: divergent
# barriers.  Output under gpurun_out/sanitize/.  The CPU-side counterparts are tests/emu/sanitize.py
# and the emulator's hazard checker.
set -u
O=gpurun_out/sanitize; mkdir -p $O
cat > /tmp/san_drive.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from madrl_b200 import BatchedMAWaterWorld, BatchedPursuitEvade, BatchedHostageWorld
torch.manual_seed(0)
ww = BatchedMAWaterWorld(64, n_pursuers=5, n_evaders=5, seed=1); ww.reset()
ww.rollout(torch.randn(6, 64, 5, 2, device='cuda') * 0.5)
c4 = BatchedMAWaterWorld(16, n_pursuers=20, n_evaders=50, n_poison=50, seed=1); c4.reset()
c4.rollout(torch.randn(3, 16, 20, 2, device='cuda') * 0.5)
maps = np.load(os.path.join("maps", "map_pool16.npy"))
pe = BatchedPursuitEvade(64, maps, n_evaders=30, n_pursuers=8, obs_range=7, sample_maps=True, reward_mech='local',
                         catchr=0.1, seed=1, max_path_length=5)
pe.reset(); pe.rollout(torch.randint(0, 5, (12, 64, 8), dtype=torch.int32, device='cuda'))
hw = BatchedHostageWorld(64, 10, 16, 16, 4, 2, seed=1, max_path_length=4); hw.reset()
hw.rollout(torch.randn(9, 64, 10, 2, device='cuda'))
# closed-loop (POLICY) instantiations, the stand-alone generators, path packing
o = ww.reset(); ww.rollout_heuristic(6, o)
o = c4.reset(); c4.rollout_heuristic(3, o)
o = pe.reset(); pe.rollout_heuristic(12, o); pe.rollout_heuristic(5, o, py2_division=False)
from madrl_b200.heuristics import waterworld_heuristic, pursuit_heuristic
waterworld_heuristic(ww.reset(), 30); pursuit_heuristic(pe.reset(), obs_range=7)
from madrl_b200.postproc import pack_paths
a = torch.randn(6, 64, 5, 2, device='cuda') * 0.5
ob, rw, dn, inf = ww.rollout(a)
pack_paths(ob, a, rw, dn)
torch.cuda.synchronize(); print("driver done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 170 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_drive.py > $O/$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/$tool.log | tail -n 1)"
done
