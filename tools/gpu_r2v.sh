#!/bin/bash
# 8-GPU: C5 over (transport, pipelined, post CTAs) + the default bench line (what the driver's scaling run records).
G=8
O=gpurun_out
mkdir -p $O
run() {  # tag, extra args
  tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --workload c5 --no-cpu --no-compare --steps 3 "$@" > $O/b_c5_8gpu_$tag.json 2> $O/b_c5_8gpu_$tag.err
  python - $O/b_c5_8gpu_$tag.json <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'fps %.2f'%d['value'], 'pipelined', d['config'].get('pipelined'), d['config']['transport'], [(k['kernel'][:12], round(k['ms_per_launch'],2)) for k in d['roofline']['kernels']], 'exch ms %.2f'%d['roofline']['nvlink']['exchange_ms_per_frame'], 'nvlink frac %.2f'%d['roofline']['nvlink']['frac'], d['config']['slab_vs_single_gpu_check']['ok'])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
run peer_pipe2
run peer_nopipe --no-pipeline
run peer_pipe4 --post-ctas 4
run peer_pipe1 --post-ctas 1
run a2a_pipe --transport alltoall
run a2a_nopipe --transport alltoall --no-pipeline
S=$(date +%s); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G > $O/b_default_8gpu_r2v.json 2> $O/b_default_8gpu_r2v.err; echo "default 8gpu exit $? wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/b_default_8gpu_r2v.json") if l.startswith('{')][-1])
print('C2', d['value'], {k:(round(v['value'],1), v['scaling']) for k,v in d.get('configs',{}).items()})
PY
