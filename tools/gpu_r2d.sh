#!/bin/bash
# Round-2 call D: where does the fused column kernel lose its time? (ncu --set full with source), plus stream-count probes.
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ow_col2 -s 6 -c 2 -f -o $O/prof_c2_col2fused_r2d \
    python bench.py --workload c2 --col-kernel 2 --fused 1 --profile --steps 1 --warmup 1 > $O/ncu_r2d_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ow_col2 -s 6 -c 2 -f -o $O/prof_c2_col3_r2d \
    python bench.py --workload c2 --col-kernel 3 --fused 0 --profile --steps 1 --warmup 1 > $O/ncu_r2d_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ow_row_bulk -s 3 -c 2 -f -o $O/prof_c2_rowbulk_r2d \
    python bench.py --workload c2 --row-kernel 3 --profile --steps 1 --warmup 1 > $O/ncu_r2d_c.log 2>&1
for s in 3 4; do
  timeout 300 python bench.py --workload c3 --streams $s --no-cpu --no-compare --steps 3 > $O/b_c3_st$s.json 2> $O/b_c3_st$s.err
  python -c "
import json
d=json.loads([l for l in open('$O/b_c3_st$s.json') if l.startswith('{')][-1]); print('c3 streams=$s %.0f fps'%d['value'])"
done
for g in "0 0" "1 8" "2 4" ; do
  set -- $g
  timeout 300 python bench.py --workload c4 --c4-shard-of 8 --streams ${1/0/3} --group $2 --no-cpu --no-compare --steps 5 > $O/b_c4s8_$1_$2.json 2> $O/b_c4s8_$1_$2.err
  python -c "
import json
d=json.loads([l for l in open('$O/b_c4s8_$1_$2.json') if l.startswith('{')][-1]); print('c4 shard-of-8 streams=$1 group=$2 %.0f fps (8 cascades)'%d['value'], d['config']['launch_groups_per_step'])"
done
ls -la $O/*.ncu-rep | tail -4
