# The other BASELINE configurations through bench.py (device-resident number
# only; they are parity-test cases, not the bench line).
mkdir -p gpurun_out
rm -f gpurun_out/r1l_other_workloads.jsonl
timeout 300 python bench.py --workload scp --lanes 1024 --limit 100 --steps 3 --no-cpu-baseline --no-e2e >> gpurun_out/r1l_other_workloads.jsonl 2>> gpurun_out/other.err
timeout 300 python bench.py --workload csplib022 --lanes 296 --limit 20 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/r1l_other_workloads.jsonl 2>> gpurun_out/other.err
timeout 300 python bench.py --workload telebus --lanes 1024 --limit 50 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/r1l_other_workloads.jsonl 2>> gpurun_out/other.err
timeout 400 python bench.py --workload rail-small --lanes 148 --limit 5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/r1l_other_workloads.jsonl 2>> gpurun_out/other.err
python - <<'PY'
import json
for l in open('gpurun_out/r1l_other_workloads.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:40], d['config']['lanes_per_gpu'], round(d['value']/1e6,3),'Mupd/s', round(d['ms_per_step'],1),'ms', 'frac',round(d['roofline']['frac'],3), 'sweeps/s', round(d['sweeps_per_s']))
PY
tail -3 gpurun_out/other.err
