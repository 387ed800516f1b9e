mkdir -p gpurun_out
rm -f gpurun_out/exp8.log
for cfg in "" "BXG_GATHER=entry" "BXG_THREADS=96" "BXG_THREADS=256"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp8.log 2>&1
done
cat gpurun_out/exp8.log
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
