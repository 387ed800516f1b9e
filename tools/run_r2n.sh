# Round-2 measurement on one B200: the whole GPU suite, both bench arms, the ncu launch list of
# the bench command, one full capture of the dominant kernel, the other BASELINE workloads.
R=${1:-r2n}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1
grep -v "^$" gpurun_out/${R}_pytest_gpu.log | tail -n 5
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
cut -c1-700 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
cut -c1-300 gpurun_out/${R}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_bench.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_optimize -s 1 -c 1 -o gpurun_out/${R}_optimize -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_full.log 2>&1
tail -n 2 gpurun_out/${R}_ncu_full.log | cut -c1-200
rm -f gpurun_out/${R}_other_workloads.jsonl
timeout 300 python bench.py --workload telebus --mode solve --limit 50 --steps 3 --cpu-lanes 256 --no-e2e >> gpurun_out/${R}_other_workloads.jsonl 2>> gpurun_out/${R}_other.err
timeout 400 python bench.py --workload telebus --rounds 1 --steps 2 --warmup 1 --cpu-seconds 10 --no-e2e >> gpurun_out/${R}_other_workloads.jsonl 2>> gpurun_out/${R}_other.err
timeout 300 python bench.py --workload scp --mode solve --limit 100 --steps 3 --no-cpu-baseline --no-e2e >> gpurun_out/${R}_other_workloads.jsonl 2>> gpurun_out/${R}_other.err
timeout 400 python bench.py --workload csplib022 --mode solve --lanes 296 --limit 20 --steps 2 --warmup 3 --cpu-lanes 32 --no-e2e >> gpurun_out/${R}_other_workloads.jsonl 2>> gpurun_out/${R}_other.err
python - <<PY
import json
for l in open('gpurun_out/${R}_other_workloads.jsonl'):
    d=json.loads(l); c=d.get('cpu_baseline') or {}
    print(d['config']['workload'][:34], d['config']['mode'][:8], round(d['value']/1e6,3),'Mupd/s', round(d['ms_per_step'],1),'ms', 'frac',round(d['roofline']['frac'],4), 'cpu', c.get('value'), c.get('cores'))
PY
tail -n 3 gpurun_out/${R}_other.err
