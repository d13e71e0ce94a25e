#!/usr/bin/env bash
# compute-sanitizer over a subset of the GPU tests: memcheck (out-of-bounds / misaligned) and racecheck (shared-memory
# hazards) on the shared-memory scan kernels, the PointCloud2 chain, the decoders and the framing kernel
set -u
mkdir -p gpurun_out
T=${1:-r2san}
SUB='shared_final or mode_a_duplicate or mixed_batch or extreme_values or odd_stride'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan_parity.py -q -x -k "$SUB" > gpurun_out/${T}_memcheck_scan.txt 2>&1; echo "memcheck scan rc=$?"; tail -4 gpurun_out/${T}_memcheck_scan.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decode_formats.py tests/test_gpu_framing.py tests/test_gpu_cloud.py -q -x > gpurun_out/${T}_memcheck_dec.txt 2>&1; echo "memcheck decode/framing/cloud rc=$?"; tail -4 gpurun_out/${T}_memcheck_dec.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_scan_parity.py -q -x -k "shared_final or mode_a_duplicate or mixed_batch" > gpurun_out/${T}_racecheck_scan.txt 2>&1; echo "racecheck scan rc=$?"; grep -c "Race reported\|hazard" gpurun_out/${T}_racecheck_scan.txt; tail -6 gpurun_out/${T}_racecheck_scan.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_decode_formats.py tests/test_gpu_cloud.py -q -x -k "batched_ragged or smoothing_chain or chain_strides or ties_and_dense or ragged_and_empty" > gpurun_out/${T}_racecheck_dec.txt 2>&1; echo "racecheck decode/cloud rc=$?"; grep -c "Race reported\|hazard" gpurun_out/${T}_racecheck_dec.txt; tail -6 gpurun_out/${T}_racecheck_dec.txt
