#!/bin/bash
O=gpurun_out
mkdir -p $O
for v in "2" "1" "4" "0"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload c5 --no-cpu --no-compare --steps 3 --post-ctas $v > $O/b_c5_2gpu_pc$v.json 2> $O/b_c5_2gpu_pc$v.err
  echo "c5 2gpu post-ctas $v exit $?"
  python - $O/b_c5_2gpu_pc$v.json <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'fps %.2f'%d['value'], 'pipelined', d['config'].get('pipelined'), [(k['kernel'][:12], round(k['ms_per_launch'],2)) for k in d['roofline']['kernels']], d['config']['slab_vs_single_gpu_check']['ok'])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
