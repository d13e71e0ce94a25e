#!/usr/bin/env bash
# Mode A iteration: all GPU tests, the 3200-node lines (Mode B line carries Mode A and the ascended buffer under extra),
# ncu of the Mode A kernel
set -u
mkdir -p gpurun_out
T=${1:-r2t}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 600 python bench.py --nodes 3200 --scans 40960 --steps 50 --no-cpu --no-cloud --no-e2e > gpurun_out/${T}_scan3200.json 2> gpurun_out/${T}_scan3200.err; tail -c 300 gpurun_out/${T}_scan3200.err
timeout 600 python bench.py --nodes 3200 --scans 40960 --mode a --steps 50 --no-cpu --no-cloud --no-e2e --no-extra > gpurun_out/${T}_scan3200_a.json 2> gpurun_out/${T}_scan3200_a.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:scan_small_kernel<.int.1' -c 1 -f -o gpurun_out/${T}_ncu_scan3200_a python bench.py --nodes 3200 --scans 40960 --mode a --steps 1 --no-cpu --no-cloud --no-e2e --no-extra > /dev/null 2> gpurun_out/${T}_ncu_scan3200_a.log
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), {k:(v if not isinstance(v,dict) else round(v['mpoints_s'])) for k,v in d.get('extra',{}).items() if k in ('mode_a_mpoints_s','with_ascended_nodes_out')})
    except Exception as e:
        print(f,'ERR',e)
PY
