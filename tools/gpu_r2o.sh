#!/bin/bash
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_big.py tests/test_gpu_slab.py -m gpu -x -q > $O/pytest_r2o.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2o.log )
tail -15 $O/pytest_r2o.log
for m in 0 1 3 7; do
  timeout 600 python bench.py --workload c5 --line-clusters $m --no-cpu --no-compare --steps 2 --warmup 1 > $O/b_c5_lc$m.json 2> $O/b_c5_lc$m.err
  python - $O/b_c5_lc$m.json <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'fps %.2f'%d['value'], 'clusters', d['config'].get('line_clusters'), [(k['kernel'][:12], round(k['ms_per_launch'],2)) for k in d['roofline']['kernels']], 'frame frac %.3f'%d['roofline']['frame']['frac'])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
