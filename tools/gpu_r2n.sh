#!/bin/bash
O=gpurun_out
mkdir -p $O
for n in 256 512 1024; do
  timeout 120 python tools/single_frame_probe.py $n 2>&1 | tee -a $O/single_frame_r2n.txt
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_single_$n.csv python tools/single_frame_probe.py $n --plain > /dev/null 2>&1
  python - $O/launches_single_$n.csv <<'PY' | tee -a $O/single_frame_r2n.txt
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
for r in rows[-6:]: print("   ", r[4].split('<')[0].replace('void ow::',''), r[8], r[7], r[14], r[13])
PY
done
timeout 600 python tools/sweep_modes.py c4 --c4-shard-of 8 --rows 0 --cols 0:0 --streams 3,4 --groups 0 --reps 9 --prio 2>&1 | tee $O/sweep_prio_r2n.txt
timeout 600 python tools/sweep_modes.py c2 c3 c4 --rows 0 --cols 0:0 --streams 3 --groups 0 --reps 5 --prio 2>&1 | tee -a $O/sweep_prio_r2n.txt
