#!/bin/bash
# A/B: balanced persistent grids (default lib) vs grids = min(items, resident) (variant lib "nb"), alternating, same box.
O=gpurun_out
L=$PWD/fft-ocean-waves_b200/lib
one() {  # label, lib, args...
  lbl=$1; lib=$2; shift; shift
  OCEANWAVES_LIB=$lib timeout 300 python bench.py --no-cpu --no-compare "$@" 2>/dev/null > /tmp/ab.json
  python - "$lbl" "$*" <<'PY'
import json,sys
d=json.loads([l for l in open('/tmp/ab.json') if l.startswith('{')][-1]); print(sys.argv[1], sys.argv[2], round(d['value']), round(d['ms_per_step'],4))
PY
}
for i in 1 2 3; do
  one balanced $L/liboceanwaves.so --workload c4 --c4-shard-of 8 --steps 15
  one plain    $L/liboceanwaves_nb.so --workload c4 --c4-shard-of 8 --steps 15
done
one balanced $L/liboceanwaves.so --workload c4 --steps 5
one plain    $L/liboceanwaves_nb.so --workload c4 --steps 5
one balanced $L/liboceanwaves.so --workload c3 --steps 5
one plain    $L/liboceanwaves_nb.so --workload c3 --steps 5
