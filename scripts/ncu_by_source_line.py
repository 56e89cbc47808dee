"""Per-SOURCE-LINE dynamic warp-instruction counts and stall shares of one kernel (read here, on the CPU box):
`nvdisasm -g` of the cubin (line info per SASS instruction, needs -lineinfo) joined, instruction by instruction, with
the per-SASS "Instructions Executed" / stall samples of an `ncu --set full` capture of the same build.

    cuobjdump -xelf all madrl_b200/build/pursuit.o && nvdisasm -g pursuit.sm_100a.cubin > pe_all.txt
    ncu -i gpurun_out/r2g_pe.ncu-rep --page source --csv > pe_src.csv
    python scripts/ncu_by_source_line.py pe_all.txt <mangled kernel name> pe_src.csv <units per launch>

This is what found Pursuit's 85-instruction capture test, the per-launch grid rebuild and the layout branch that kept
the window loop from unrolling (C3 0.57 -> 0.70 of the roofline)."""
import csv, re, sys
dis, sec, srccsv, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lines=open(dis).read().split("\n")
start=[i for i,l in enumerate(lines) if l.startswith(".text."+sec+":")][0]
cur=None; seq=[]
for l in lines[start+1:]:
    if l.startswith(".text.") or l.startswith("\t.section"): 
        if seq: break
    m=re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m=re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m: seq.append((int(m.group(1),16), cur, m.group(2)))
rows=list(csv.reader(open(srccsv))); hdr=rows[1]
ie=hdr.index("Instructions Executed"); ia=hdr.index("Warp Stall Sampling (All Samples)")
ex=[(int(r[ie]), int(r[ia]), r[1].strip()) for r in rows[2:] if len(r)>ie]
assert len(ex)==len(seq), (len(ex), len(seq))
agg={}
for (addr, loc, txt), (n, st, t2) in zip(seq, ex):
    a=agg.setdefault(loc, [0,0]); a[0]+=n; a[1]+=st
tot=sum(v[0] for v in agg.values()); tots=sum(v[1] for v in agg.values())
print("total warp instr / unit: %.1f" % (tot/units))
src={}
for k,v in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1]) if kv[0] else ("",0)):
    if v[0]/units >= 3:
        f,ln=k if k else ("?",0)
        if f not in src:
            try: src[f]=open("/root/repo/madrl_b200/csrc/"+f).read().split("\n")
            except Exception: src[f]=[]
        text=src[f][ln-1].strip()[:90] if src[f] and ln-1 < len(src[f]) else ""
        print("%-14s %4d  %7.1f instr  %5.1f%% stalls  %s" % (f, ln, v[0]/units, 100*v[1]/tots, text))
