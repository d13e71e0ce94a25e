#!/usr/bin/env bash
# all GPU tests + the six decoder lines + the chain line
set -u
mkdir -p gpurun_out
T=${1:-r2x}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
for f in 0x81 0x82 0x83 0x84 0x85 0x86; do
  timeout 300 python bench.py --workload decode --format $f --steps 30 > gpurun_out/${T}_dec_${f}.json 2> gpurun_out/${T}_dec_${f}.err; tail -c 200 gpurun_out/${T}_dec_${f}.err
done
timeout 600 python bench.py --workload chain --steps 50 > gpurun_out/${T}_chain.json 2> gpurun_out/${T}_chain.err
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', round(d.get('roofline',{}).get('frac',0),4), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY
