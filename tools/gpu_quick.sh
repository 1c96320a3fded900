#!/bin/bash
# Quick GPU check: parity tests (optional) + the three bench workloads without the CPU leg.
# usage: bash tools/gpu_quick.sh [tag] [notest]
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
if [ "$2" != "notest" ]; then
  ( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log ); tail -4 $O/pytest_$TAG.log
fi
for w in c2 c3 c4; do
  timeout 600 python bench.py --workload $w --no-cpu > $O/b_$w.json 2> $O/b_$w.err; echo "bench $w exit $?"; tail -2 $O/b_$w.err
done
python tools/summ.py c2 c3 c4
