mkdir -p gpurun_out
cd tools/_exp/old && timeout 600 python ../rail_check.py 1.0 > ../../../gpurun_out/rail_old.log 2>&1; cd ../../..
tail -4 gpurun_out/rail_old.log
timeout 600 python tools/_exp/rail_check.py 1.0 > gpurun_out/rail_new.log 2>&1
tail -4 gpurun_out/rail_new.log
timeout 600 python tools/_exp/rail_check.py 0.25 > gpurun_out/rail_new_q.log 2>&1
tail -4 gpurun_out/rail_new_q.log
