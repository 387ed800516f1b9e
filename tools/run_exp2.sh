mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/exp2.log
for cfg in "" "BXG_THREADS=128" "BXG_THREADS=192" "BXG_THREADS=256"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp2.log 2>&1
done
cat gpurun_out/exp2.log
