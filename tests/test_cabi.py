"""The C-ABI library loads on a CPU-only machine and exports every symbol include/*.h declares
(no compute calls: those need a GPU)."""
import ctypes as C
import glob
import os
import re

from conftest import ROOT


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(madrl_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from madrl_b200 import _lib
    lib = _lib.lib()
    names = declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert lib.madrl_version() >= 100


def test_waterworld_layout_is_host_only_and_sane():
    from madrl_b200 import _lib
    lib = _lib.lib()
    cfg = _lib.WWConfig(n_envs=4096, n_pursuers=5, n_evaders=5, n_poison=10, n_sensors=30, n_coop=2,
                        addid=1, speed_features=1, timestep_limit=1000)
    lay = _lib.WWLayout()
    assert lib.madrl_ww_state_layout(C.byref(cfg), C.byref(lay)) == 0
    assert lay.obs_dim == 213 and lay.n_obj == 20 and lay.real_bytes == 4
    offs = [lay.rng_counter, lay.objs, lay.obst, lay.timestep, lay.path_len, lay.sensors]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and lay.total_bytes > offs[-1]
    bad = _lib.WWConfig(n_envs=4, n_pursuers=40, n_evaders=5, n_poison=10, n_sensors=30, n_coop=2,
                        timestep_limit=1000)
    assert lib.madrl_ww_state_layout(C.byref(bad), C.byref(lay)) == -1
    assert b"n_pursuers" in lib.madrl_last_error()
