# rail4284-shaped (C5), 1 GPU: engine arm (optimizer, 148 workers) with e2e, then the reference's optimizer on the host cores
R=r2k
mkdir -p gpurun_out
( time timeout 900 python bench.py --workload rail --lanes 148 --rounds 5 --steps 3 --warmup 1 --no-cpu-baseline ) > gpurun_out/${R}_rail_n1.json 2> gpurun_out/${R}_rail_n1.err
cut -c1-1800 gpurun_out/${R}_rail_n1.json; tail -n 4 gpurun_out/${R}_rail_n1.err
( time timeout 900 python bench.py --impl reference --workload rail --lanes 148 --cpu-seconds 60 --cpu-builds o3 --steps 1 --warmup 0 ) > gpurun_out/${R}_rail_reference.json 2> gpurun_out/${R}_rail_reference.err
cut -c1-1500 gpurun_out/${R}_rail_reference.json; tail -n 4 gpurun_out/${R}_rail_reference.err
