#!/bin/bash
# Round-2 ncu evidence (run inside ONE gpurun call on one GPU; reports land in gpurun_out/):
#   1. launch list of the bench command (device time per launch)
#   2. DRAM traffic of the DP kernel in the bench (one launch)
#   3. ncu --set full of the throughput kernel and of the latency kernel on small cases
set -x
cd "$(dirname "$0")/.."
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/r02_ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
T="python tools/prof_case.py 1024 0 1"
$T > gpurun_out/r02_ncu_plain_traffic.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fpop_dp_kernel -c 1 --csv --log-file gpurun_out/r02_dp_traffic.csv $T > gpurun_out/r02_ncu_traffic.log 2>&1
C="python tools/prof_case.py 480 4000 1"
$C > gpurun_out/r02_ncu_plain_thr.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fpop_dp_kernel -c 1 -f -o gpurun_out/r02_thr $C > gpurun_out/r02_ncu_thr.log 2>&1
export PSD_LATENCY_MODE=1
L="python tools/prof_case.py 29 20000 1"
$L > gpurun_out/r02_ncu_plain_lat.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fpop_dp_lat_kernel -c 1 -f -o gpurun_out/r02_lat $L > gpurun_out/r02_ncu_lat.log 2>&1
ls -la gpurun_out/ | tail -12
