mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 ) > gpurun_out/r2s_pytest_gpu.log 2>&1
timeout 300 python tools/exp_perf.py --workload csplib022 --lanes 296 --limit 20 --reps 2 > gpurun_out/r2s_csplib.log 2>&1
timeout 100 python tools/exp_perf.py --lanes 1024 --reps 2 > gpurun_out/r2s_c2.log 2>&1
timeout 200 python bench.py --workload csplib022 --mode solve --lanes 296 --limit 20 --steps 2 --warmup 3 --no-e2e > gpurun_out/r2s_bench_csplib.json 2> gpurun_out/r2s_bench_csplib.err
cut -c1-300 gpurun_out/r2s_pytest_gpu.log gpurun_out/r2s_csplib.log gpurun_out/r2s_c2.log
