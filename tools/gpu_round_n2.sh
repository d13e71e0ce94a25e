#!/usr/bin/env bash
# two GPUs: the multi-GPU parity check (push kernel + C++ exchange vs torch all-gather), the default bench line at
# N=2 (carries extra.cloud with the three exchanges), the cloud workload on its own line, the reference arm at N=2
set -u
mkdir -p gpurun_out
T=${1:-r2n2final}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi_push.py -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.txt; tail -5 gpurun_out/${T}_pytest.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29521 bench.py --gpus 2 --steps 50 > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err; tail -c 1500 gpurun_out/${T}_default.err
timeout 600 $RUN --master-port 29522 bench.py --gpus 2 --workload cloud --steps 30 > gpurun_out/${T}_cloud.json 2> gpurun_out/${T}_cloud.err; tail -c 800 gpurun_out/${T}_cloud.err
timeout 600 $RUN --master-port 29523 bench.py --gpus 2 --workload cloud --sor 8 --steps 20 > gpurun_out/${T}_cloud_sor.json 2> gpurun_out/${T}_cloud_sor.err
NCCL_DEBUG=INFO timeout 600 $RUN --master-port 29524 bench.py --gpus 2 --workload cloud --steps 3 > /dev/null 2> gpurun_out/${T}_nccl_info.log; grep -iE "NVLS|via P2P|Channel 00|NCCL version|Connected all" gpurun_out/${T}_nccl_info.log | sort | uniq -c | head -12
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'))
        x=d.get('extra',{})
        c=x.get('cloud', x if 'by_exchange' in x else None)
        if c:
            if 'error' in c: print('   CLOUD ERROR', c['error'], c.get('trace','')[-800:])
            else:
                print('   cloud impl',c['impl'],'ms',c['ms_per_step'],'compute',c['compute_ms'],'identical',c['exchanges_bit_identical'])
                for k,v in c['by_exchange'].items(): print('     ',k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items()})
    except Exception as e:
        print(f,'ERR',e)
PY
