#!/bin/bash
# Development aid: source-level instruction counts of one kernel (ncu --import-source, --page source), top lines only.
cd "$(dirname "$0")/.."
K=${KERNEL:-k_partition_coarse}; LG=${LG:-24}
SIZES=$LG NOPINT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$K -s 1 -c 1 -f -o gpurun_out/pc_src python tools/quick_bench.py > /dev/null 2>&1
ncu -i gpurun_out/pc_src.ncu-rep --page source --csv > gpurun_out/pc_source_full.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/pc_src.ncu-rep > gpurun_out/pc_summary.csv
rm -f gpurun_out/pc_src.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/pc_source_full.csv")))
hdr=None
for i,r in enumerate(rows):
    if "Source" in r and any("Instructions Executed" in c for c in r): hdr=i; break
if hdr is None:
    print("no source table; header candidates:", [r[:6] for r in rows[:5]]); raise SystemExit
h=rows[hdr]; si=h.index("Source"); ii=[k for k,c in enumerate(h) if c=="Instructions Executed"][0]
li=h.index("#") if "#" in h else 0
out=[]
for r in rows[hdr+1:]:
    try: out.append((float(r[ii].replace(",","")), r[li], r[si].strip()[:110]))
    except Exception: pass
tot=sum(o[0] for o in out) or 1
out.sort(reverse=True)
with open("gpurun_out/pc_source_top.txt","w") as f:
    for v,l,s in out[:45]:
        f.write("%5.1f%%  line %s  %s\n"%(100*v/tot,l,s))
print(open("gpurun_out/pc_source_top.txt").read())
PY
rm -f gpurun_out/pc_source_full.csv
