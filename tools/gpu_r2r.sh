#!/bin/bash
# 2-GPU: multi-rank slab tests (pipelined peer transport), C5 bench pipelined vs not.
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_slab.py -m gpu -x -q -k "nvlink or one_gpu" > $O/pytest_r2r.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2r.log )
tail -8 $O/pytest_r2r.log
for v in "" "--no-pipeline"; do
  tag=pipe; [ -n "$v" ] && tag=nopipe
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload c5 --no-cpu --no-compare --steps 3 $v > $O/b_c5_2gpu_$tag.json 2> $O/b_c5_2gpu_$tag.err
  echo "c5 2gpu $tag exit $?"; tail -2 $O/b_c5_2gpu_$tag.err
  python - $O/b_c5_2gpu_$tag.json <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'fps %.2f'%d['value'], 'pipelined', d['config'].get('pipelined'), [(k['kernel'][:12], round(k['ms_per_launch'],2)) for k in d['roofline']['kernels']], 'nvlink frac %.3f'%d['roofline']['nvlink']['frac'], d['config']['slab_vs_single_gpu_check'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
