#!/bin/bash
G=${1:-2}
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_big.py tests/test_gpu_slab.py -m gpu -x -q > $O/pytest_r2aa.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2aa.log )
tail -4 $O/pytest_r2aa.log
for cfg in "peer -1" "peer 0"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --workload c5 --transport $1 --post-ctas $2 --no-cpu --no-compare --steps 3 > $O/b_c5_${G}gpu_aa_$1_$2.json 2> $O/b_c5_${G}gpu_aa_$1_$2.err
  python - $O/b_c5_${G}gpu_aa_$1_$2.json <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'fps %.2f'%d['value'], 'pipelined', d['config'].get('pipelined'), d['config']['transport'], [(k['kernel'][:12], round(k['ms_per_launch'],2)) for k in d['roofline']['kernels']], 'nvlink frac %.2f'%d['roofline']['nvlink']['frac'], d['config']['slab_vs_single_gpu_check']['ok'])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
timeout 600 python bench.py --workload c5 --no-cpu --no-compare --steps 2 > $O/b_c5_1gpu_aa.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/b_c5_1gpu_aa.json') if l.startswith('{')][-1]); print('c5 1gpu', round(d['value'],2), [(k['kernel'][:12], round(k['ms_per_launch'],2)) for k in d['roofline']['kernels']])"
