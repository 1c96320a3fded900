#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python tools/sweep_modes.py c3 --rows 0 --cols 4:0 --streams 3 --groups 1,3,4 --l2 0,1 --reps 5 2>&1 | tee $O/sweep_r2l.txt
timeout 600 python tools/sweep_modes.py c2 --rows 0 --cols 1:0 --streams 3 --groups 0 --l2 0,1 --reps 5 2>&1 | tee -a $O/sweep_r2l.txt
timeout 600 python tools/sweep_modes.py c4 --rows 0 --cols 1:0,4:0 --streams 3 --groups 0 --l2 0 --reps 5 2>&1 | tee -a $O/sweep_r2l.txt
