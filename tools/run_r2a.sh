# r2a: new optimizer tests first (fast feedback), then the whole GPU suite and a quick bench.
R=r2a
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_optimizer.py -x -q ) > gpurun_out/${R}_pytest_opt.log 2>&1
tail -25 gpurun_out/${R}_pytest_opt.log
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_optimizer.py ) > gpurun_out/${R}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
cat gpurun_out/${R}_bench.json | cut -c1-600
