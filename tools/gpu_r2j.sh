#!/bin/bash
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "variants" > $O/pytest_r2j.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2j.log )
tail -3 $O/pytest_r2j.log
timeout 600 python tools/sweep_modes.py c3 c2 c4 --rows 0 --cols 1:0,4:0 --streams 3 --groups 0 --reps 5 2>&1 | tee $O/sweep_r2j.txt
timeout 600 python tools/sweep_modes.py c3 --rows 0 --cols 4:0 --streams 1,2,4 --groups 0,1,2 --caps 0:0 --reps 5 2>&1 | tee -a $O/sweep_r2j.txt
