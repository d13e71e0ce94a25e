#!/bin/bash
# DRAM bytes + duration of the dominant kernel for every experimental build in tools/libs
for lib in tools/libs/*.so ""; do
  export RPL_B200_LIB=$([ -n "$lib" ] && echo $PWD/$lib)
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:scan_tma -s 3 -c 1 --csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra 2>/dev/null | grep -E "scan_tma" | awk -F, '{print $(NF-2), $(NF-1), $NF}' | tr "\n" " "; echo " <- ${lib:-default}"
done
