#!/usr/bin/env bash
# one GPU visit: parity tests, smoke, bench lines, ncu captures of the new kernels -> gpurun_out/
set -u
mkdir -p gpurun_out
T=${1:-r2a}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.txt
cat gpurun_out/${T}_pytest.txt | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --workload cloud --steps 20 > gpurun_out/${T}_cloud.json 2> gpurun_out/${T}_cloud.err; tail -c 600 gpurun_out/${T}_cloud.err
timeout 600 python bench.py --workload cloud --sor 8 --steps 20 --no-cpu > gpurun_out/${T}_cloud_sor.json 2> gpurun_out/${T}_cloud_sor.err
timeout 600 python bench.py --nodes 3200 --scans 40960 --steps 50 --no-cpu --no-cloud --no-e2e > gpurun_out/${T}_scan3200.json 2> gpurun_out/${T}_scan3200.err; tail -c 400 gpurun_out/${T}_scan3200.err
timeout 900 python bench.py --steps 50 > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err; tail -c 400 gpurun_out/${T}_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_small -c 3 -f -o gpurun_out/${T}_ncu_cloud python bench.py --workload cloud --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_cloud.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_small -c 2 -f -o gpurun_out/${T}_ncu_scan3200 python bench.py --nodes 3200 --scans 40960 --steps 1 --no-cpu --no-cloud --no-e2e --no-extra > /dev/null 2> gpurun_out/${T}_ncu_scan3200.log
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'))
        x=d.get('extra',{})
        for k in ('cloud','with_ascended_nodes_out','mode_a_mpoints_s','compute_ms','by_exchange'):
            if k in x: print('   ',k, json.dumps(x[k])[:700])
    except Exception as e:
        print(f,'ERR',e)
PY
