#!/usr/bin/env bash
# ncu --set full of the six decoder kernels in their final form (one launch each)
set -u
mkdir -p gpurun_out
T=${1:-r2w}
cap() { # tag format regex
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -c 1 -f -o gpurun_out/${T}_ncu_$1 python bench.py --workload decode --format $2 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_$1.log
}
cap normal 0x81 decode_normal_kernel
cap express 0x82 'decode_capsule_kernel<.int.0'
cap hq 0x83 decode_hq_kernel
cap ultra 0x84 'decode_capsule_kernel<.int.1'
cap dense 0x85 decode_dense_kernel
cap ultradense 0x86 'decode_capsule_kernel<.int.2'
ls -la gpurun_out/${T}_ncu_*.ncu-rep
