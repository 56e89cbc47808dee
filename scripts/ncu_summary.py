"""Summarise an .ncu-rep capture (read here, on the CPU box): key metrics + hottest SASS lines.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [units_per_launch] [top_n]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'not_selected', 'math_pipe_throttle', 'lg_throttle',
          'mio_throttle', 'dispatch_stall', 'branch_resolving', 'no_instruction', 'barrier', 'selected', 'drain',
          'imc_miss', 'tex_throttle', 'sleeping', 'membar', 'misc']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[-1]


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hdr, unit, val = raw(rep)
    d = {h: (v, u) for h, v, u in zip(hdr, val, unit)}
    print("kernel:", d.get('Kernel Name', ('?',))[0])
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
    for s in STALLS:
        k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
        if k in d:
            print("| stall %s (warps per issue) | %s | |" % (s, d[k][0]))
    if units and 'smsp__inst_executed.sum' in d:
        print("| warp-instructions per unit | %.1f | (units per launch = %g) |" % (float(d['smsp__inst_executed.sum'][0].replace(',', '')) / units, units))
    if top:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[1]
        iS, iEx, iSa = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
        data = rows[2:]
        tot_s = sum(int(r[iSa]) for r in data)
        print("\nhottest instructions by stall samples (of %d):" % tot_s)
        for r in sorted(data, key=lambda r: -int(r[iSa]))[:top]:
            print("%6d %5.2f%%  x%-8.2f %s" % (int(r[iSa]), 100.0 * int(r[iSa]) / max(1, tot_s),
                                              (int(r[iEx]) / units) if units else int(r[iEx]), r[iS].strip()[:90]))


if __name__ == "__main__":
    main()
