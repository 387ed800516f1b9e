mkdir -p gpurun_out
rm -f gpurun_out/exp6.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for cfg in "" "BXG_THREADS=96"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp6.log 2>&1
done
cat gpurun_out/exp6.log
timeout 600 python bench.py > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err
cat gpurun_out/r1i_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1i_bench_reference.json 2>> gpurun_out/r1i_bench.err
cat gpurun_out/r1i_bench_reference.json
tail -5 gpurun_out/r1i_bench.err
