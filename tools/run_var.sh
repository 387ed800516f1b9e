mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; python -c "
import json; d=json.load(open('gpurun_out/bench_check.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_check2.json 2>> gpurun_out/bench_check.err; python -c "
import json; d=json.load(open('gpurun_out/bench_check2.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
