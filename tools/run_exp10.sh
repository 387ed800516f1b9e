mkdir -p gpurun_out
timeout 600 python tools/_exp/rail_check.py 1.0 > gpurun_out/rail_new.log 2>&1
tail -3 gpurun_out/rail_new.log | cut -c1-400
BXG_THREADS=256 timeout 600 python tools/_exp/rail_check.py 0.25 > gpurun_out/rail_new_q.log 2>&1
tail -3 gpurun_out/rail_new_q.log | cut -c1-400
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/exp_perf.py --lanes 1024 > gpurun_out/exp10.log 2>&1
cat gpurun_out/exp10.log | cut -c1-400
