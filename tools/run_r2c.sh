R=${1:-r2c}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_optimizer.py -x -q ) > gpurun_out/${R}_pytest.log 2>&1
grep -v "^$" gpurun_out/${R}_pytest.log | tail -6
timeout 600 python bench.py --mode solve --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/${R}_bench_solve.json 2> gpurun_out/${R}_bench.err
cut -c1-300 gpurun_out/${R}_bench_solve.json; tail -3 gpurun_out/${R}_bench.err
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/${R}_bench_opt.json 2>> gpurun_out/${R}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${R}_bench_solve.json","gpurun_out/${R}_bench_opt.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["roofline"]["frac"], d["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
