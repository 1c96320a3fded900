#!/bin/bash
O=gpurun_out
mkdir -p $O
L=$PWD/fft-ocean-waves_b200/lib
( timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "variants or fused" > $O/pytest_r2f.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2f.log )
tail -3 $O/pytest_r2f.log
echo "== default lib"; timeout 300 python tools/sweep_modes.py c2 c3 --rows 1,2 --cols 1:0 2>&1 | tee $O/sweep_r2f_default.txt
for v in v1 v2 v3 v4 v5; do
  echo "== $v"; OCEANWAVES_LIB=$L/liboceanwaves_$v.so timeout 300 python tools/sweep_modes.py c2 --rows 1,2 --cols 1:0 2>&1 | tee $O/sweep_r2f_$v.txt
done
