#!/bin/bash
# Round-2 call C: does re-using a small set of intermediate buffers (few slots per call) keep the intermediate L2-resident?
O=gpurun_out
mkdir -p $O
for cfg in "1 1" "2 2" "3 3" "4 3" "8 3" "32 3"; do
  set -- $cfg
  for ck in 1 3; do
    timeout 300 python bench.py --workload c3 --slots $1 --streams $2 --col-kernel $ck --fused 0 --no-cpu --no-compare --steps 3 > $O/b_c3_sl$1_ck$ck.json 2> $O/b_c3_sl$1_ck$ck.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/b_c3_sl$1_ck$ck.json") if l.startswith("{")][-1])
    print("c3 slots=$1 streams=$2 col=$ck : %.0f fps"%d["value"], [(k["kernel"][3:6], round(k["ms_per_launch"]*1e3,1)) for k in d["roofline"]["kernels"]])
except Exception as e: print("c3 slots=$1 col=$ck ERR", e)
PY
  done
done
for cfg in "8 1 8" "16 2 8" "24 3 8" "48 3 16" "128 3 0"; do
  set -- $cfg
  timeout 300 python bench.py --workload c2 --slots $1 --streams $2 --group $3 --no-cpu --no-compare --steps 3 > $O/b_c2_sl$1.json 2> $O/b_c2_sl$1.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/b_c2_sl$1.json") if l.startswith("{")][-1])
    print("c2 slots=$1 streams=$2 group=$3 : %.0f fps"%d["value"], [(k["kernel"][3:6], round(k["ms_per_launch"]*1e3,1)) for k in d["roofline"]["kernels"]])
except Exception as e: print("c2 slots=$1 ERR", e)
PY
done
