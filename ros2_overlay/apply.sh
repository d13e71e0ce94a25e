#!/usr/bin/env bash
# Applies the B200 overlay to a checkout of frozenreboot/rplidar_ros2_driver.
#   ros2_overlay/apply.sh <path-to-rplidar_ros2_driver checkout> [<output dir>]
# Without an output dir the checkout is patched in place; with one, the four files the overlay touches are
# copied there (src/, include/) and patched -- what the tests and ros2_overlay/CMakeLists.txt do.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$1"; OUT="${2:-}"
FILES=(src/rplidar_node.cpp src/lidar_driver_wrapper.cpp include/rplidar_node.hpp include/lidar_driver_wrapper.hpp)
if [[ -n "$OUT" ]]; then
  mkdir -p "$OUT/src" "$OUT/include"
  for f in "${FILES[@]}"; do cp "$SRC/$f" "$OUT/$f"; done
  TARGET="$OUT"
else
  TARGET="$SRC"
fi
for p in "$HERE"/patches/*.patch; do patch -d "$TARGET" -p1 --forward --no-backup-if-mismatch < "$p"; done
echo "overlay applied to $TARGET"
