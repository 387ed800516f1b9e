mkdir -p gpurun_out
rm -f gpurun_out/exp4.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( BXG_GATHER=entry timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/pytest_gpu_entry.log 2>&1
tail -2 gpurun_out/pytest_gpu_entry.log
for cfg in "" "BXG_GATHER=entry" "BXG_THREADS=96" "BXG_THREADS=256"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp4.log 2>&1
done
cat gpurun_out/exp4.log
