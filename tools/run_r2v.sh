mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 12 ) > gpurun_out/r2v_pytest_gpu.log 2>&1
timeout 300 python tools/exp_perf.py --workload csplib022 --lanes 296 --limit 20 --reps 2 > gpurun_out/r2v_csplib.log 2>&1
timeout 100 python tools/exp_perf.py --lanes 1024 --reps 2 > gpurun_out/r2v_c2.log 2>&1
cut -c1-400 gpurun_out/r2v_pytest_gpu.log gpurun_out/r2v_csplib.log gpurun_out/r2v_c2.log
