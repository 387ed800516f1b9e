#!/usr/bin/env python3
"""usage: profiles/ncu_lines.py report.ncu-rep [top]
Per-source-line warp-stall samples of an `ncu --set full --import-source on`
report (kernels built with -lineinfo), read on the CPU box."""
import csv, subprocess, sys, io, collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None
hdr = None
agg = collections.Counter()
inst = collections.Counter()
src = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        key = (cur_file, int(r[0]))
        try:
            agg[key] += int(r[i_samp])
            inst[key] += int(r[i_inst])
        except ValueError:
            pass
        src[key] = r[1].strip()
total = sum(agg.values()) or 1
tinst = sum(inst.values()) or 1
print(f"total samples {total}, warp instructions {tinst}")
for key, v in agg.most_common(top):
    print(f"{100*v/total:5.1f}% samp {100*inst[key]/tinst:5.1f}% inst  {key[0]}:{key[1]:<5d} {src[key][:110]}")
