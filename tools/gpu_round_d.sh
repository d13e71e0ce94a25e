#!/usr/bin/env bash
# decoders: parity tests, bench lines for every answer type, ncu of ultra / ultra-dense
set -u
mkdir -p gpurun_out
T=${1:-r2h}
timeout 900 python -m pytest tests -m gpu -q -k "decode or capsule or framing or chain or assemble or wire" 2>&1 | tail -15 > gpurun_out/${T}_pytest_dec.txt
tail -5 gpurun_out/${T}_pytest_dec.txt
for f in 0x81 0x82 0x83 0x84 0x85 0x86; do
  timeout 300 python bench.py --workload decode --format $f --steps 30 > gpurun_out/${T}_dec_${f}.json 2> gpurun_out/${T}_dec_${f}.err; tail -c 300 gpurun_out/${T}_dec_${f}.err
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:decode_capsule_kernel<.int.1' -c 1 -f -o gpurun_out/${T}_ncu_dec84 python bench.py --workload decode --format 0x84 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_dec84.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:decode_capsule_kernel<.int.2' -c 1 -f -o gpurun_out/${T}_ncu_dec86 python bench.py --workload decode --format 0x86 --steps 1 --no-cpu > /dev/null 2> gpurun_out/${T}_ncu_dec86.log
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_dec_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY
