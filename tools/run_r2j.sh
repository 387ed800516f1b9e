# 2 GPUs: exchange check, then the default bench (optimizer on C2, exchange in the loop) at N=2
R=r2j
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_exchange.py > gpurun_out/${R}_exchange.log 2>&1; tail -n 8 gpurun_out/${R}_exchange.log
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err
cut -c1-1300 gpurun_out/${R}_bench_n2.json; tail -n 3 gpurun_out/${R}_bench_n2.err
