#!/bin/bash
# ncu captures for the hot kernels (run under gpurun, one GPU).  Usage: profiles/run_ncu.sh <tag>
# 1) launch list with device time per launch (cold-cache, serialised: compare SHARES, not absolutes)
# 2) one --set full capture of each hot kernel, source-correlated (-lineinfo)
TAG=${1:-r01}
WL=${2:-splash256}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
for K in ${3:-k_p2g_tile k_g2p_brick k_assemble k_p2g_finalize k_build_index}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${K}_${TAG} $CMD > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out
