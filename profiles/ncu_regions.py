#!/usr/bin/env python3
"""usage: profiles/ncu_regions.py report.ncu-rep source.cu
Warp-stall samples and warp instructions per region of the source file; a
region starts at each line holding a '// @region name' marker."""
import csv, subprocess, sys, io, collections, re
rep, src = sys.argv[1], sys.argv[2]
marks = []
for n, line in enumerate(open(src), 1):
    m = re.search(r"@region\s+(\S+)", line)
    if m:
        marks.append((n, m.group(1)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
samp = collections.Counter(); inst = collections.Counter()
base = src.split("/")[-1]
def region(line):
    name = "(before)"
    for n, nm in marks:
        if line >= n: name = nm
        else: break
    return name
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        key = region(int(r[0])) if cur == base else cur
        try: samp[key] += int(r[i_s]); inst[key] += int(r[i_i])
        except ValueError: pass
ts = sum(samp.values()) or 1; ti = sum(inst.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for k, v in sorted(samp.items(), key=lambda kv: -kv[1]):
    print(f"{k:28s} samples {100*v/ts:5.1f}%   instructions {100*inst[k]/ti:5.1f}%")
