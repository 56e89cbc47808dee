"""profiles/r2_traffic.json from the `ncu --set full` captures of scripts/ncu_capture.sh (read here, on
the CPU box):  python scripts/ncu_traffic.py r2a
For each kernel: dram__bytes_read.sum + dram__bytes_write.sum of the captured launch divided by the
env-steps that launch computed, plus the hash of the kernel's source files at the time this script ran
(run it right after the capture, on the tree that was captured) -- bench.py reports `roofline.traffic` /
`frac_dram` only while the kernel is still built from those files.
usage: python scripts/ncu_traffic.py TAG [TAG...]   (later tags win)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from madrl_b200.build import kernel_hash  # noqa: E402

# capture name -> env-steps of the captured launch (scripts/{ww,pe,hw}_sweep.py shapes)
SHAPES = {"ww_c2": 4096 * 256, "ww_c4": 4096 * 16, "pe": 65536 * 32, "hw": 8192 * 32}
KEYS = {"ww_c2": "ww_c2", "ww_c4": "ww_c4", "pe": "pe_c3", "hw": "hw_c5"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    tags = sys.argv[1:] or ["r2"]
    out = {"kernels": {}}
    for cap, steps in SHAPES.items():
        reps = [os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (t, cap)) for t in tags]
        reps = [r for r in reps if os.path.exists(r)]
        if not reps:
            continue
        rep = reps[-1]
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[-1])}
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        t_ns = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        t_ns *= {"nsecond": 1, "usecond": 1e3, "msecond": 1e6, "ns": 1, "us": 1e3, "ms": 1e6}.get(d["gpu__time_duration.sum"][1], 1)
        out["kernels"][KEYS[cap]] = {
            "kernel": d.get("Kernel Name", ("?",))[0], "env_steps_per_launch": steps,
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_env_step": (rd + wr) / steps,
            "duration_under_ncu_ms": t_ns / 1e6, "capture": os.path.relpath(rep, ROOT),
            "kernel_hash": kernel_hash(cap.split("_")[0])}
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
