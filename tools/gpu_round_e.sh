#!/usr/bin/env bash
# quick iteration: all GPU tests, the 3200-node LaserScan line, ultra / ultra-dense decoder lines + ncu
set -u
mkdir -p gpurun_out
T=${1:-r2m}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.txt
tail -4 gpurun_out/${T}_pytest.txt
timeout 600 python bench.py --nodes 3200 --scans 40960 --steps 50 --no-cpu --no-cloud --no-e2e > gpurun_out/${T}_scan3200.json 2> gpurun_out/${T}_scan3200.err; tail -c 400 gpurun_out/${T}_scan3200.err
for f in 0x82 0x85; do
  timeout 300 python bench.py --workload decode --format $f --steps 30 --no-cpu > gpurun_out/${T}_dec_${f}.json 2> gpurun_out/${T}_dec_${f}.err; tail -c 300 gpurun_out/${T}_dec_${f}.err
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:decode_capsule_kernel<.int.1' -c 1 -f -o gpurun_out/${T}_ncu_dec84 python bench.py --workload decode --format 0x84 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_dec84.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:decode_capsule_kernel<.int.2' -c 1 -f -o gpurun_out/${T}_ncu_dec86 python bench.py --workload decode --format 0x86 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_dec86.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:scan_small_kernel<.int.0, .bool.1' -c 1 -f -o gpurun_out/${T}_ncu_scan3200_emit python bench.py --nodes 3200 --scans 40960 --steps 1 --no-cpu --no-cloud --no-e2e > /dev/null 2> gpurun_out/${T}_ncu_scan3200_emit.log
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('input_buffer'))
        x=d.get('extra',{})
        for k in ('with_ascended_nodes_out','mode_a_mpoints_s'):
            if k in x: print('   ',k, json.dumps(x[k])[:300])
    except Exception as e:
        print(f,'ERR',e)
PY
