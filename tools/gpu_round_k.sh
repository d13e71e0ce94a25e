#!/usr/bin/env bash
# the capsule -> LaserScan chain: per-launch durations (ncu launch list) next to the bench's own stage times
set -u
mkdir -p gpurun_out
T=${1:-r2z}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_chain_launches.csv python bench.py --workload chain --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_chain_under_ncu.log 2>&1
timeout 300 python bench.py --workload chain --steps 50 --no-cpu > gpurun_out/${T}_chain.json 2> gpurun_out/${T}_chain.err
T=$T python - <<'PY'
import csv,os,json
T=os.environ.get('T','r2z')
rows=list(csv.reader(open(f'gpurun_out/{T}_chain_launches.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r;continue
    if hdr and len(r)==len(hdr):
        print(r[hdr.index('ID')], r[hdr.index('Kernel Name')][:70], r[hdr.index('Grid Size')], r[hdr.index('Metric Value')])
d=json.loads(open(f'gpurun_out/{T}_chain.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['extra'])
PY
