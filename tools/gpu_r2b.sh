#!/bin/bash
# Round-2 call B: correctness of the new kernels (bulk-async row kernel, TMA-staged / fused column kernel), then mode sweeps.
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "variants or fused or graph" > $O/pytest_r2b.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2b.log )
tail -15 $O/pytest_r2b.log
timeout 600 python tools/sweep_modes.py c2 --check > $O/sweep_c2_r2b.txt 2>&1; tail -20 $O/sweep_c2_r2b.txt
timeout 600 python tools/sweep_modes.py c3 --check --streams 1,3 > $O/sweep_c3_r2b.txt 2>&1; tail -34 $O/sweep_c3_r2b.txt
timeout 600 python tools/sweep_modes.py c4 --check > $O/sweep_c4_r2b.txt 2>&1; tail -20 $O/sweep_c4_r2b.txt
OCEANWAVES_LIB=$PWD/fft-ocean-waves_b200/lib/liboceanwaves_c2mb2.so timeout 600 python tools/sweep_modes.py c4 --rows 2,3 --cols 2:0,2:1,3:0,3:1 > $O/sweep_c4_c2mb2_r2b.txt 2>&1; tail -10 $O/sweep_c4_c2mb2_r2b.txt
