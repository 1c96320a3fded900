#!/bin/bash
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "variants or fused" > $O/pytest_r2e.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2e.log )
tail -5 $O/pytest_r2e.log
timeout 600 python tools/sweep_modes.py c2 --rows 1,2 --cols 1:0,2:1,3:1 > $O/sweep_c2_r2e.txt 2>&1; cat $O/sweep_c2_r2e.txt
timeout 600 python tools/sweep_modes.py c3 --rows 1 --cols 1:0,2:0,2:1,3:1 > $O/sweep_c3_r2e.txt 2>&1; cat $O/sweep_c3_r2e.txt
timeout 600 python tools/sweep_modes.py c4 --rows 2 --cols 1:0,2:1,3:1 > $O/sweep_c4_r2e.txt 2>&1; cat $O/sweep_c4_r2e.txt
OCEANWAVES_LIB=$PWD/fft-ocean-waves_b200/lib/liboceanwaves_c2mb2.so timeout 600 python tools/sweep_modes.py c4 --rows 2 --cols 2:1,3:1 > $O/sweep_c4_c2mb2_r2e.txt 2>&1; cat $O/sweep_c4_c2mb2_r2e.txt
OCEANWAVES_LIB=$PWD/fft-ocean-waves_b200/lib/liboceanwaves_c2s4.so timeout 600 python tools/sweep_modes.py c2 --rows 1 --cols 2:0,2:1,3:1 > $O/sweep_c2_c2s4_r2e.txt 2>&1; cat $O/sweep_c2_c2s4_r2e.txt
