#!/bin/bash
# One gpurun call: GPU parity tests, the three bench workloads, ncu launch list + full captures.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_$TAG.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log )
tail -3 $O/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
for w in c2 c3 c4 c5; do
  timeout 900 python bench.py --workload $w > $O/b_$w.json 2> $O/b_$w.err; echo "bench $w exit $?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/b_ref.json 2> $O/b_ref.err
python tools/summ.py c2 c3 c4 c5
# ncu: launch list (shares) for the default workload and C3, then full captures of each kernel at C3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c2_$TAG.csv \
    python bench.py --workload c2 --profile --steps 1 --warmup 1 > $O/ncu_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_$TAG.csv \
    python bench.py --workload c3 --profile --steps 1 --warmup 1 > $O/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ow_ -s 30 -c 6 -f -o $O/prof_c3_$TAG \
    python bench.py --workload c3 --profile --steps 1 --warmup 1 > $O/ncu_c3_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ow_ -s 30 -c 6 -f -o $O/prof_c2_$TAG \
    python bench.py --workload c2 --profile --steps 1 --warmup 1 > $O/ncu_c2_full.log 2>&1
ls -la $O
