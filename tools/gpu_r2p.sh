#!/bin/bash
# 2-GPU validation of the default bench under torchrun (what the driver's scaling run does) + reference arm + C4 shard probe.
O=gpurun_out
mkdir -p $O
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > $O/b_default_2gpu_r2p.json 2> $O/b_default_2gpu_r2p.err
echo "default 2gpu exit $? wall $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 > $O/b_ref_2gpu_r2p.json 2> $O/b_ref_2gpu_r2p.err
echo "ref 2gpu exit $? wall $(( $(date +%s) - S )) s"
tail -3 $O/b_default_2gpu_r2p.err
python - <<'PY'
import json
for f in ("gpurun_out/b_default_2gpu_r2p.json","gpurun_out/b_ref_2gpu_r2p.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d.get('impl'), d['value'], d['n_gpus'], d.get('cpu_baseline',{}).get('cores'), {k:(v['value'], v['n_gpus'], v['scaling']) for k,v in d.get('configs',{}).items()})
    except Exception as e: print(f,'ERR',e)
PY
timeout 300 python bench.py --workload c4 --c4-shard-of 8 --no-cpu --no-compare --steps 9 > $O/b_c4s8_graph_r2p.json 2>/dev/null
timeout 300 python bench.py --workload c4 --c4-shard-of 8 --no-cpu --no-compare --steps 9 --no-graph > $O/b_c4s8_nograph_r2p.json 2>/dev/null
timeout 300 python bench.py --workload c4 --no-cpu --no-compare --steps 5 > $O/b_c4_r2p.json 2>/dev/null
python - <<'PY'
import json
for f in ("b_c4s8_graph_r2p","b_c4s8_nograph_r2p","b_c4_r2p"):
    d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith('{')][-1]); print(f, round(d['value']), d['ms_per_step'], d['config'].get('submission'))
PY
