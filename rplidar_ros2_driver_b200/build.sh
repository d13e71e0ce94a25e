#!/usr/bin/env bash
# Builds librplidar_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="${RPL_OUT:-$HERE/librplidar_b200.so}"
SRCS=("$HERE"/csrc/rpl_capi.cu "$HERE"/csrc/scan_fast.cu "$HERE"/csrc/scan_tma.cu "$HERE"/csrc/scan_general.cu \
      "$HERE"/csrc/synth.cu "$HERE"/csrc/cloud.cu "$HERE"/csrc/decode.cu "$HERE"/csrc/assemble.cu "$HERE"/csrc/decode_formats.cu "$HERE"/csrc/timestamps.cu "$HERE"/csrc/cdr.cu)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false \
       -Xcompiler -fPIC,-O2,-ffp-contract=off,-Wall,-Wno-address-of-packed-member -shared -cudart shared)
if [[ "${RPL_PTXAS_V:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
# extra -D tuning switches for experiments, e.g. RPL_DEFS="-DRPL_TMA_CH=1024 -DRPL_TMA_STAGES=4"
if [[ -n "${RPL_DEFS:-}" ]]; then FLAGS+=(${RPL_DEFS}); fi
"$NVCC" "${FLAGS[@]}" -o "$OUT" "${SRCS[@]}"
echo "built $OUT"
