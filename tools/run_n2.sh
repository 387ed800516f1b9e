mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1j_bench_n2.json 2> gpurun_out/r1j_bench_n2.err
cat gpurun_out/r1j_bench_n2.json | cut -c1-900
tail -5 gpurun_out/r1j_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r1j_bench_reference_n2.json 2>> gpurun_out/r1j_bench_n2.err
cat gpurun_out/r1j_bench_reference_n2.json | cut -c1-400
