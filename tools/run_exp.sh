# usage: bash tools/run_exp.sh TAG lib1 lib2 ...   (names under baryonyx_b200/exp/, "product" = the built library)
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = product ]; then unset BXG_LIB; else export BXG_LIB=$PWD/baryonyx_b200/exp/lib_$v.so; fi
  echo "== $v" | tee -a gpurun_out/${TAG}_exp.jsonl
  timeout 300 python tools/exp_perf.py --lanes 1024 2>&1 | tail -n 1 | tee -a gpurun_out/${TAG}_exp.jsonl
done
