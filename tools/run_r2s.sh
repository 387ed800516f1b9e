timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py tests/test_gpu_optimizer.py -x -q 2>&1 | tail -n 3
timeout 300 python tools/exp_perf.py --workload csplib022 --lanes 296 --limit 20 --reps 2 | cut -c1-200
timeout 100 python tools/exp_perf.py --lanes 1024 --reps 2 | cut -c1-120
timeout 300 python bench.py --workload rail --lanes 148 --rounds 5 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline | cut -c1-200
