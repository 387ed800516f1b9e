mkdir -p gpurun_out
export BXG_LIB=$PWD/baryonyx_b200/libbaryonyx_b200_timers.so
timeout 200 python tools/exp_perf.py --workload csplib022 --lanes 148 --limit 20 --reps 1 > gpurun_out/r2t_csplib_timers_148.log 2>&1
timeout 200 python tools/exp_perf.py --workload csplib022 --lanes 296 --limit 20 --reps 1 > gpurun_out/r2t_csplib_timers_296.log 2>&1
cut -c1-600 gpurun_out/r2t_csplib_timers_148.log gpurun_out/r2t_csplib_timers_296.log
