#!/usr/bin/env python
"""usage: tools/make_traffic.py gpurun_out/rXX_solve.ncu-rep [workload lanes limit]
Writes profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of
the captured k_solve / k_optimize launch) -- what bench.py reports as roofline.traffic."""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "spp"
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
limit = int(sys.argv[4]) if len(sys.argv) > 4 else 100
mode = sys.argv[5] if len(sys.argv) > 5 else "solve"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, row = rows[0], rows[1], rows[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def val(name):
    i = hdr.index(name)
    return float(row[i].replace(",", "")) * scale[units[i]]


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
i = hdr.index("gpu__time_duration.sum")
d = {"workload": workload, "lanes": lanes, "limit": limit, "mode": mode, "kernel": row[hdr.index("Kernel Name")],
     "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
     "gpu_time": row[i] + " " + units[i], "source": os.path.basename(rep)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "profiles", "traffic.json"), "w") as f:
    json.dump(d, f, indent=1)
print(json.dumps(d))
