#!/usr/bin/env bash
# eight-GPU box, the end-to-end (host buffer) leg only: where does the N>=4 ceiling come from?
#   four ranks on ONE socket's GPUs (0-3) vs four ranks spread over both sockets (0,1,4,5), two ranks same / other
#   socket, and the write-combined input buffer
set -u
mkdir -p gpurun_out
T=${1:-r2n8b}
B="bench.py --steps 10 --no-cloud --no-extra --no-cpu"
run() { # tag devices nproc extra
  CUDA_VISIBLE_DEVICES=$2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) $B --gpus $3 $4 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  tail -c 300 gpurun_out/${T}_$1.err
}
run n4_one_socket 0,1,2,3 4 ""
run n4_two_sockets 0,1,4,5 4 ""
run n4_one_socket_wc 0,1,2,3 4 "--e2e-wc"
run n8_wc 0,1,2,3,4,5,6,7 8 "--e2e-wc"
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/'+os.environ['T']+'_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, 'n',d['n_gpus'],'e2e',round(e['value']),'ms',round(e['ms_per_step'],2),'per-gpu GB/s each way',round(e['h2d_bytes_per_step']/e['ms_per_step']/1e6,1), e.get('input_buffer'), 'numa', d['extra'].get('numa',{}).get('numa_node'))
    except Exception as ex:
        print(f,'ERR',ex)
PY
