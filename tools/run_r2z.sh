# racecheck with the hazards grouped by source line: smoke(), then the
# long-row rank-select parity test (cta_rank_select, chunked gather).
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool racecheck --print-limit 400 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_race_smoke.full 2>&1
grep -o "in [a-z_]*\.cuh*:[0-9]*" gpurun_out/r2z_race_smoke.full | sort | uniq -c | sort -rn > gpurun_out/r2z_racecheck_smoke_lines.txt
grep -E "RACECHECK SUMMARY|smoke ok" gpurun_out/r2z_race_smoke.full >> gpurun_out/r2z_racecheck_smoke_lines.txt
timeout 130 compute-sanitizer --tool racecheck --print-limit 400 python -m pytest tests/test_gpu_edges.py -x -q -k "rank_select and None" > gpurun_out/r2z_race_rank.full 2>&1
grep -o "in [a-z_]*\.cuh*:[0-9]*" gpurun_out/r2z_race_rank.full | sort | uniq -c | sort -rn > gpurun_out/r2z_racecheck_rank_select_lines.txt
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2z_race_rank.full >> gpurun_out/r2z_racecheck_rank_select_lines.txt
rm -f gpurun_out/r2z_race_smoke.full.gz; gzip -f gpurun_out/r2z_race_smoke.full gpurun_out/r2z_race_rank.full
cat gpurun_out/r2z_racecheck_smoke_lines.txt gpurun_out/r2z_racecheck_rank_select_lines.txt
