mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_gpu_edges.py tests/test_gpu_fullsize.py "tests/test_gpu_parity.py::test_sweeps_long_rows_with_ties" -x -q 2>&1 | tail -n 12 ) > gpurun_out/r2u_pytest.log 2>&1
timeout 200 python tools/exp_perf.py --workload csplib022 --lanes 148 --limit 20 --reps 2 > gpurun_out/r2u_csplib_148.log 2>&1
timeout 200 python tools/exp_perf.py --workload csplib022 --lanes 296 --limit 20 --reps 1 > gpurun_out/r2u_csplib_296.log 2>&1
timeout 100 python tools/exp_perf.py --lanes 1024 --reps 2 > gpurun_out/r2u_c2.log 2>&1
cut -c1-400 gpurun_out/r2u_pytest.log gpurun_out/r2u_csplib_148.log gpurun_out/r2u_csplib_296.log gpurun_out/r2u_c2.log
