mkdir -p gpurun_out
rm -f gpurun_out/exp12.log
for cfg in "" "BXG_LIB=tools/_exp/libbx_base.so" "" "BXG_LIB=tools/_exp/libbx_base.so"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp12.log 2>&1
done
cut -c1-330 gpurun_out/exp12.log
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_population.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
