#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python tools/sweep_modes.py c4 --c4-shard-of 8 --rows 0 --cols 0:0 --streams 3,4 --groups 0 --reps 9 --graph 2>&1 | tee $O/sweep_r2w.txt
timeout 600 python tools/sweep_modes.py c4 --c4-shard-of 8 --rows 0 --cols 0:0 --streams 3,4 --groups 0 --reps 9 2>&1 | tee -a $O/sweep_r2w.txt
timeout 600 python tools/sweep_modes.py c4 --rows 0 --cols 0:0 --streams 3,4 --groups 0 --reps 5 --graph 2>&1 | tee -a $O/sweep_r2w.txt
timeout 600 python tools/sweep_modes.py c2 c3 --rows 0 --cols 0:0 --streams 3,4 --groups 0 --reps 5 2>&1 | tee -a $O/sweep_r2w.txt
