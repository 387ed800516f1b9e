# Final check of round 2 on one B200: smoke(), the whole GPU suite, the bench
# line of the committed build, the csplib022-shaped line.
mkdir -p gpurun_out
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 5 ) > gpurun_out/r2w_smoke.log 2>&1
( time timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 12 ) > gpurun_out/r2w_pytest_gpu.log 2>&1
timeout 200 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
timeout 150 python bench.py --workload csplib022 --mode solve --lanes 296 --limit 20 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_csplib022.json 2>> gpurun_out/r2w_bench.err
cut -c1-700 gpurun_out/r2w_smoke.log gpurun_out/r2w_pytest_gpu.log gpurun_out/r2w_bench.json gpurun_out/r2w_csplib022.json
tail -n 5 gpurun_out/r2w_bench.err
