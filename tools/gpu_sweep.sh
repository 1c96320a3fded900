#!/bin/bash
# bench.py over (workload, streams, group) combinations; prints fps per combination.  usage: bash tools/gpu_sweep.sh "c2 c3" "1 2 3" "0 8"
O=gpurun_out; mkdir -p $O
for w in $1; do for s in $2; do for g in $3; do
  timeout 300 python bench.py --workload $w --no-cpu --streams $s --group $g --steps 3 > $O/sw.json 2> $O/sw.err || tail -3 $O/sw.err
  python - <<PY
import json
try:
    d=json.load(open("$O/sw.json")); print("$w streams=$s group=$g  fps %.0f  ms/step %.3f  frame_frac %.3f seq %s"%(d["value"],d["ms_per_step"],d["roofline"]["frame"]["frac"],d["config"]["single_slot_sequential_fps"]))
except Exception as e: print("$w $s $g ERR",e)
PY
done; done; done
