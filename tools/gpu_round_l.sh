#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
T=${1:-r2ab}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --nodes 3200 --scans 40960 --steps 50 --no-cpu --no-cloud --no-e2e > gpurun_out/${T}_scan3200.json 2> gpurun_out/${T}_scan3200.err; tail -c 300 gpurun_out/${T}_scan3200.err
timeout 600 python bench.py --workload chain --steps 50 > gpurun_out/${T}_chain.json 2> gpurun_out/${T}_chain.err; tail -c 300 gpurun_out/${T}_chain.err
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e', (d.get('e2e') or {}).get('value'), {k:(round(v,4) if isinstance(v,float) else (round(v['mpoints_s']) if isinstance(v,dict) and 'mpoints_s' in v else v)) for k,v in d.get('extra',{}).items() if k in ('ms_decode','ms_assemble','ms_scan','mode_a_mpoints_s','with_ascended_nodes_out')})
    except Exception as e:
        print(f,'ERR',e)
PY
