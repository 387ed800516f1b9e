mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1l_pytest_gpu.log 2>&1
tail -3 gpurun_out/r1l_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1l_smoke.log 2>&1; tail -2 gpurun_out/r1l_smoke.log
timeout 900 python bench.py > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; cut -c1-300 gpurun_out/r1l_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1l_bench_reference.json 2>> gpurun_out/r1l_bench.err; cut -c1-200 gpurun_out/r1l_bench_reference.json
