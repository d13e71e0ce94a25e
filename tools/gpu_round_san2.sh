#!/usr/bin/env bash
# racecheck on the shared-memory kernels without bulk-TMA staging + decoders (tools/racecheck_small.py), then the
# regular GPU tests (dense decoder change) and the chain / dense decoder lines
set -u
mkdir -p gpurun_out
T=${1:-r2san2}
timeout 300 python tools/racecheck_small.py > gpurun_out/${T}_driver_plain.txt 2>&1; tail -2 gpurun_out/${T}_driver_plain.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python tools/racecheck_small.py > gpurun_out/${T}_racecheck.txt 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/${T}_racecheck.txt; grep "Race reported" gpurun_out/${T}_racecheck.txt | sort | uniq -c | head; tail -4 gpurun_out/${T}_racecheck.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python bench.py --workload decode --format 0x85 --steps 30 --no-cpu > gpurun_out/${T}_dec_0x85.json 2> gpurun_out/${T}_dec_0x85.err
timeout 600 python bench.py --workload chain --steps 50 --no-cpu > gpurun_out/${T}_chain.json 2> gpurun_out/${T}_chain.err
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), {k:round(v,4) for k,v in d.get('extra',{}).items() if k.startswith('ms_')})
    except Exception as e:
        print(f,'ERR',e)
PY
