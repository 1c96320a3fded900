#!/bin/bash
# multi-GPU: slab tests + C5 bench over (transport, pipelined) at the box's GPU count (pass it as $1)
G=${1:-2}
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q -k "nvlink" > $O/pytest_r2t_$G.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2t_$G.log )
tail -4 $O/pytest_r2t_$G.log
for tr in peer alltoall; do for v in "" "--no-pipeline"; do
  tag=${tr}_pipe; [ -n "$v" ] && tag=${tr}_nopipe
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --workload c5 --transport $tr --no-cpu --no-compare --steps 3 $v > $O/b_c5_${G}gpu_$tag.json 2> $O/b_c5_${G}gpu_$tag.err
  python - $O/b_c5_${G}gpu_$tag.json <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'fps %.2f'%d['value'], 'pipelined', d['config'].get('pipelined'), d['config']['transport'], [(k['kernel'][:12], round(k['ms_per_launch'],2)) for k in d['roofline']['kernels']], 'exch ms %.2f'%d['roofline']['nvlink']['exchange_ms_per_frame'], d['config']['slab_vs_single_gpu_check']['ok'])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done; done
