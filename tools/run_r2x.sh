# The Z-family and long-row shapes as full bench lines (e2e + the reference's
# CPU arm on the same box), for DESIGN.md section 4's table.
mkdir -p gpurun_out
rm -f gpurun_out/r2x_other_workloads.jsonl
timeout 170 python bench.py --workload telebus --mode solve --lanes 1024 --limit 50 --steps 3 --warmup 3 >> gpurun_out/r2x_other_workloads.jsonl 2>> gpurun_out/r2x.err
timeout 200 python bench.py --workload csplib022 --mode solve --lanes 296 --limit 20 --steps 2 --warmup 3 >> gpurun_out/r2x_other_workloads.jsonl 2>> gpurun_out/r2x.err
cut -c1-300 gpurun_out/r2x_other_workloads.jsonl
tail -n 5 gpurun_out/r2x.err
