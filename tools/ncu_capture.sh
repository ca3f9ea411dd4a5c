#!/bin/bash
# One gpurun call: the launch list of a bench step and `ncu --set full` captures of the hot kernels (round 3).
# Usage (on the GPU box, from the repo root):  bash tools/ncu_capture.sh <tag>
# Every command runs plain first (exit code checked) and only then under ncu (B200_PROFILING.md).
set -u
TAG=${1:-r3}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --configs none"
timeout 200 $B > $OUT/${TAG}_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 200 $B > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_popc_kernel -s 3 -c 1 -o $OUT/${TAG}_k1 $B > $OUT/${TAG}_ncu_k1.log 2>&1
run_cfg() {   # name, kernel regex, bench_configs args...
  local name=$1 re=$2; shift 2
  local C="python tools/bench_configs.py $* --steps 1"
  timeout 200 $C > $OUT/${TAG}_plain_${name}.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -o $OUT/${TAG}_${name} $C > $OUT/${TAG}_ncu_${name}.log 2>&1
}
run_cfg k2_fp4 count_gemm_kernel c4 --rows 400000
run_cfg k2_i8 count_gemm_kernel c4 --rows 400000 --engine gemm
run_cfg k3b_fet pvalue_kernel c2b --rows 1000000
run_cfg k3b_chi pvalue_kernel c2 --rows 1000000
run_cfg k4_single coop_sort_kernel c2 --rows 1000000
run_cfg k4_multi coop_sort_kernel c3 --rows 500000
run_cfg k1_snp count_popc_ring_kernel c3 --rows 500000
ls -la $OUT | grep ${TAG}_ | head -40
