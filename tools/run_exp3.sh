mkdir -p gpurun_out
rm -f gpurun_out/exp3.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spp or scp or eq01 or golden or smoke" > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
for cfg in "" "BXG_THREADS=96" "BXG_THREADS=256"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp3.log 2>&1
done
cat gpurun_out/exp3.log
