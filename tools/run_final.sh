# Round-end measurement on one B200: GPU tests, both bench arms, the ncu launch
# list of the bench command and one full capture of the dominant kernel.
R=${1:-r1j}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
cat gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
cat gpurun_out/${R}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 1 -c 1 -o gpurun_out/${R}_solve -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_full.log 2>&1
tail -3 gpurun_out/${R}_ncu_full.log
python tools/exp_perf.py --lanes 1024 > gpurun_out/${R}_phases.json 2>&1
cat gpurun_out/${R}_phases.json
