# rail4284-shaped (C5) on 8 GPUs: exchange check, the optimizer with the incumbent exchange in the loop,
# then the reference's optimizer on all host cores of the same box
R=r2l
mkdir -p gpurun_out
nproc > gpurun_out/${R}_nproc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/check_exchange.py > gpurun_out/${R}_exchange.log 2>&1; tail -n 3 gpurun_out/${R}_exchange.log
( time timeout 900 $TR bench.py --gpus 8 --workload rail --lanes 148 --rounds 5 --steps 5 --warmup 1 --no-e2e ) > gpurun_out/${R}_rail_n8.json 2> gpurun_out/${R}_rail_n8.err
cut -c1-1800 gpurun_out/${R}_rail_n8.json; tail -n 4 gpurun_out/${R}_rail_n8.err
( time timeout 600 python bench.py --impl reference --gpus 8 --workload rail --lanes 148 --cpu-seconds 60 --cpu-builds o3 --steps 1 --warmup 0 ) > gpurun_out/${R}_rail_reference.json 2> gpurun_out/${R}_rail_reference.err
cut -c1-1500 gpurun_out/${R}_rail_reference.json; tail -n 4 gpurun_out/${R}_rail_reference.err
