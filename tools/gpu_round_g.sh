#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
T=${1:-r2p}
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:decode_hq_kernel' -c 1 -f -o gpurun_out/${T}_ncu_hq python bench.py --workload decode --format 0x83 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_hq.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:decode_normal_kernel' -c 1 -f -o gpurun_out/${T}_ncu_normal python bench.py --workload decode --format 0x81 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_normal.log
ls -la gpurun_out/${T}_ncu_*
