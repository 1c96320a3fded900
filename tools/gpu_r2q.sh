#!/bin/bash
# Round-2 profile call: bench lines of every workload, ncu launch lists, ncu --set full captures (C3 and C2), column-kernel timeline.
O=gpurun_out
TAG=r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_$TAG.txt 2>&1
S=$(date +%s); timeout 900 python bench.py > $O/b_default_$TAG.json 2> $O/b_default_$TAG.err; echo "bench default exit $? wall $(( $(date +%s) - S )) s"
for w in c3 c4 c5; do
  timeout 900 python bench.py --workload $w --no-cpu > $O/b_${w}_$TAG.json 2> $O/b_${w}_$TAG.err; echo "bench $w exit $?"
done
timeout 600 python bench.py --impl reference > $O/b_ref_$TAG.json 2> $O/b_ref_$TAG.err
python tools/summ.py default_$TAG c3_$TAG c4_$TAG c5_$TAG 2>&1 | tee $O/summ_$TAG.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c2_$TAG.csv \
    python bench.py --workload c2 --profile --steps 1 --warmup 1 > $O/ncu_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_$TAG.csv \
    python bench.py --workload c3 --profile --steps 1 --warmup 1 > $O/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ow_ -s 30 -c 6 -f -o $O/prof_c3_$TAG \
    python bench.py --workload c3 --profile --steps 1 --warmup 1 > $O/ncu_c3_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ow_ -s 30 -c 6 -f -o $O/prof_c2_$TAG \
    python bench.py --workload c2 --profile --steps 1 --warmup 1 > $O/ncu_c2_full.log 2>&1
( cd tools/tune && ./trace 2048 c 8 && ./trace 1024 c 8 && ./trace 512 c 32 ) > $O/col_cta_timeline_$TAG.txt 2>&1
ls -la $O/*$TAG* | tail -20
