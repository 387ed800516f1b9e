mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/exp1.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
for cfg in "" "BXG_MINB6=1" "BXG_THREADS=96" "BXG_THREADS=64" "BXG_THREADS=256" "BXG_THREADS=192"; do
  env $cfg timeout 300 python tools/exp_perf.py --lanes 1024 >> gpurun_out/exp1.log 2>&1
done
BXG_MINB6=1 timeout 300 python tools/exp_perf.py --lanes 888 >> gpurun_out/exp1.log 2>&1
timeout 300 python tools/exp_perf.py --lanes 2048 >> gpurun_out/exp1.log 2>&1
BXG_THREADS=96 timeout 300 python tools/exp_perf.py --lanes 2048 >> gpurun_out/exp1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve -c 1 -o gpurun_out/r1h_solve -f python tools/exp_perf.py --lanes 1024 --reps 1 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/exp1.log
