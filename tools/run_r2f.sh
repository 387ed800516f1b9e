R=${1:-r2f}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/${R}_pytest.log 2>&1
grep -v "^$" gpurun_out/${R}_pytest.log | tail -n 8
bash tools/run_exp.sh ${R} product pipe_t
