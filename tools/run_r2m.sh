R=r2m
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_integration.py tests/test_gpu_optimizer.py tests/test_gpu_edges.py -x -q ) > gpurun_out/${R}_pytest.log 2>&1
grep -v "^$" gpurun_out/${R}_pytest.log | tail -n 8
timeout 600 python bench.py --workload telebus --mode solve --limit 50 --steps 3 --cpu-lanes 256 > gpurun_out/${R}_telebus_solve.json 2> gpurun_out/${R}_telebus.err
cut -c1-400 gpurun_out/${R}_telebus_solve.json; tail -n 3 gpurun_out/${R}_telebus.err
timeout 600 python bench.py --workload telebus --steps 3 --rounds 2 --cpu-seconds 10 > gpurun_out/${R}_telebus_opt.json 2>> gpurun_out/${R}_telebus.err
python - <<PY
import json
for f in ("gpurun_out/${R}_telebus_solve.json","gpurun_out/${R}_telebus_opt.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "frac", d["roofline"]["frac"], "ms", d["ms_per_step"], "cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"), "e2e", d.get("e2e",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
