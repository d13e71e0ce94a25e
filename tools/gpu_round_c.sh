#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
T=${1:-r2final2}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_pytest.txt
tail -12 gpurun_out/${T}_pytest.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --workload cloud --steps 20 --no-cpu > gpurun_out/${T}_cloud.json 2> gpurun_out/${T}_cloud.err; tail -c 600 gpurun_out/${T}_cloud.err
timeout 600 python bench.py --nodes 3200 --scans 40960 --steps 50 --no-cpu --no-cloud --no-e2e > gpurun_out/${T}_scan3200.json 2> gpurun_out/${T}_scan3200.err; tail -c 400 gpurun_out/${T}_scan3200.err
timeout 600 python bench.py --nodes 3200 --scans 40960 --mode a --steps 50 --no-cpu --no-cloud --no-e2e --no-extra > gpurun_out/${T}_scan3200_a.json 2> gpurun_out/${T}_scan3200_a.err
timeout 900 python bench.py --steps 100 > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err; tail -c 400 gpurun_out/${T}_default.err
timeout 600 python bench.py --workload cloud --sor 8 --steps 20 --no-cpu > gpurun_out/${T}_cloud_sor.json 2> gpurun_out/${T}_cloud_sor.err
timeout 600 python bench.py --workload chain --steps 50 > gpurun_out/${T}_chain.json 2> gpurun_out/${T}_chain.err; tail -c 300 gpurun_out/${T}_chain.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:scan_small_kernel<.int.1' -c 1 -f -o gpurun_out/${T}_ncu_scan3200_a python bench.py --nodes 3200 --scans 40960 --steps 1 --no-cpu --no-cloud --no-e2e > /dev/null 2> gpurun_out/${T}_ncu_scan3200_a.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:scan_small_kernel<.int.0, .bool.1' -c 1 -f -o gpurun_out/${T}_ncu_scan3200_emit python bench.py --nodes 3200 --scans 40960 --steps 1 --no-cpu --no-cloud --no-e2e > /dev/null 2> gpurun_out/${T}_ncu_scan3200_emit.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_launches_bench.log 2>&1
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'))
        x=d.get('extra',{})
        for k in ('with_ascended_nodes_out','mode_a_mpoints_s','mode_b_mpoints_s','compute_ms','ms_decode','ms_assemble','ms_scan'):
            if k in x: print('   ',k, json.dumps(x[k])[:300])
    except Exception as e:
        print(f,'ERR',e)
PY
