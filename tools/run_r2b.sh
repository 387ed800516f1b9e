R=r2b
mkdir -p gpurun_out
lscpu | head -20 > gpurun_out/${R}_lscpu.txt
( time timeout 600 python -m pytest tests/test_gpu_optimizer.py tests/test_gpu_parity.py -x -q ) > gpurun_out/${R}_pytest_opt.log 2>&1
tail -4 gpurun_out/${R}_pytest_opt.log
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
cut -c1-1500 gpurun_out/${R}_bench.json; tail -3 gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
cut -c1-400 gpurun_out/${R}_bench_reference.json
timeout 600 python bench.py --mode solve --no-cpu-baseline --steps 3 > gpurun_out/${R}_bench_solve.json 2>> gpurun_out/${R}_bench.err
cut -c1-300 gpurun_out/${R}_bench_solve.json
( time timeout 900 python bench.py --workload rail --lanes 148 --rounds 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e ) > gpurun_out/${R}_bench_rail.json 2> gpurun_out/${R}_bench_rail.err
cut -c1-1500 gpurun_out/${R}_bench_rail.json; tail -5 gpurun_out/${R}_bench_rail.err
