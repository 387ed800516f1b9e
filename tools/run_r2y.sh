# Launch list of the bench command on the final build, then compute-sanitizer
# (memcheck, racecheck) over smoke(): sweeps, a whole solve and optimizer
# restarts of a small instance.
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2y_ncu_bench.log 2>&1
grep -c k_ gpurun_out/r2y_launches.csv
( time timeout 120 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 25 ) > gpurun_out/r2y_memcheck_smoke.log 2>&1
( time timeout 120 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 25 ) > gpurun_out/r2y_racecheck_smoke.log 2>&1
tail -n 8 gpurun_out/r2y_memcheck_smoke.log gpurun_out/r2y_racecheck_smoke.log
