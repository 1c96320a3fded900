#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python tools/sweep_modes.py c3 c2 c4 --rows 0 --cols 1:0,4:0 --streams 3 --groups 0 --reps 5 2>&1 | tee $O/sweep_r2k.txt
timeout 600 python tools/sweep_modes.py c3 --rows 0 --cols 4:0 --streams 3,4 --groups 2,3,4,8 --reps 5 2>&1 | tee -a $O/sweep_r2k.txt
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_r2k.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2k.log )
tail -3 $O/pytest_r2k.log
