mkdir -p gpurun_out
rm -f gpurun_out/exp5.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for cfg in "" "BXG_LIB=tools/_exp/libbx_u8.so" "BXG_THREADS=96" "BXG_THREADS=96 BXG_LIB=tools/_exp/libbx_u8.so"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp5.log 2>&1
done
BXG_LIB=tools/_exp/libbx_u8.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_u8.log 2>&1
tail -2 gpurun_out/pytest_gpu_u8.log
cat gpurun_out/exp5.log
