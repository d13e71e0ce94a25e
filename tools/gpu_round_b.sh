#!/usr/bin/env bash
# GPU visit: parity tests, smoke, bench lines (cloud, cloud+SOR, 3200-node scans, chain, default), ncu of the cloud kernel
set -u
mkdir -p gpurun_out
T=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest.txt
tail -8 gpurun_out/${T}_pytest.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --workload cloud --steps 20 --no-cpu > gpurun_out/${T}_cloud.json 2> gpurun_out/${T}_cloud.err; tail -c 600 gpurun_out/${T}_cloud.err
timeout 600 python bench.py --workload cloud --sor 8 --steps 20 --no-cpu > gpurun_out/${T}_cloud_sor.json 2> gpurun_out/${T}_cloud_sor.err
timeout 600 python bench.py --nodes 3200 --scans 40960 --steps 50 --no-cpu --no-cloud --no-e2e > gpurun_out/${T}_scan3200.json 2> gpurun_out/${T}_scan3200.err; tail -c 400 gpurun_out/${T}_scan3200.err
timeout 600 python bench.py --workload chain --steps 50 > gpurun_out/${T}_chain.json 2> gpurun_out/${T}_chain.err; tail -c 400 gpurun_out/${T}_chain.err
timeout 600 python bench.py --workload chain --chain-copy --steps 50 --no-e2e > gpurun_out/${T}_chain_copy.json 2> gpurun_out/${T}_chain_copy.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_small -c 1 -f -o gpurun_out/${T}_ncu_cloud python bench.py --workload cloud --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_cloud.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_small -c 1 -f -o gpurun_out/${T}_ncu_cloud_sor python bench.py --workload cloud --sor 8 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_cloud_sor.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_capsule -c 1 -f -o gpurun_out/${T}_ncu_dec84 python bench.py --workload decode --format 0x84 --steps 1 > /dev/null 2> gpurun_out/${T}_ncu_dec84.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_capsule -c 1 -f -o gpurun_out/${T}_ncu_dec86 python bench.py --workload decode --format 0x86 --steps 1 > /dev/null 2> gpurun_out/${T}_ncu_dec86.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tma -c 1 -f -o gpurun_out/${T}_ncu_modea python bench.py --mode a --steps 1 --no-cpu --no-cloud --no-e2e --no-extra > /dev/null 2> gpurun_out/${T}_ncu_modea.log
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'))
        x=d.get('extra',{})
        for k in ('with_ascended_nodes_out','mode_a_mpoints_s','compute_ms','ms_decode','ms_assemble','ms_scan'):
            if k in x: print('   ',k, json.dumps(x[k])[:300])
    except Exception as e:
        print(f,'ERR',e)
PY
