#!/usr/bin/env bash
# PointCloud2 chain: tests, bench lines, ncu
set -u
mkdir -p gpurun_out
T=${1:-r2s}
timeout 900 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_cdr.py -q 2>&1 | tail -4
timeout 600 python bench.py --workload cloud --steps 20 --no-cpu > gpurun_out/${T}_cloud.json 2> gpurun_out/${T}_cloud.err; tail -c 300 gpurun_out/${T}_cloud.err
timeout 600 python bench.py --workload cloud --sor 8 --steps 20 --no-cpu > gpurun_out/${T}_cloud_sor.json 2> gpurun_out/${T}_cloud_sor.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_small -c 1 -f -o gpurun_out/${T}_ncu_cloud python bench.py --workload cloud --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_cloud.log
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'compute_ms', d.get('extra',{}).get('compute_ms'))
    except Exception as e:
        print(f,'ERR',e)
PY
