#!/bin/bash
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_r2m.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2m.log )
tail -3 $O/pytest_r2m.log
timeout 600 python tools/sweep_modes.py c3 c2 c4 --rows 0 --cols 0:0 --streams 3 --groups 0 --reps 5 2>&1 | tee $O/sweep_r2m.txt
timeout 600 python tools/sweep_modes.py c4 --c4-shard-of 8 --rows 0 --cols 0:0,4:0 --streams 3,4 --groups 0 --reps 9 2>&1 | tee -a $O/sweep_r2m.txt
