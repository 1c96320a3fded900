#!/bin/bash
# Round-2 call A: new GPU tests, default bench line (with configs.*), stream-count and L2 experiments at C3.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_r2a.txt 2>&1
python - > $O/l2props_r2a.txt 2>&1 <<'PY'
import ctypes, torch
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")
for name, a in (("L2CacheSize", 38), ("MaxPersistingL2CacheSize", 108), ("MaxAccessPolicyWindowSize", 109), ("MultiProcessorCount", 16), ("MaxSharedMemoryPerBlockOptin", 97)):
    v = ctypes.c_int(0); rc = rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0); print(name, rc, v.value)
PY
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_r2a.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2a.log )
tail -5 $O/pytest_r2a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2a.log 2>&1; tail -2 $O/smoke_r2a.log
( time timeout 900 python bench.py ) > $O/b_default_r2a.json 2> $O/b_default_r2a.err; echo "bench default exit $?"; tail -4 $O/b_default_r2a.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/b_ref_r2a.json 2> $O/b_ref_r2a.err
for s in 1 2 3; do
  timeout 300 python bench.py --workload c3 --streams $s --no-cpu --no-compare > $O/b_c3_s$s.json 2> $O/b_c3_s$s.err; echo "c3 streams $s exit $?"
done
timeout 300 python bench.py --workload c3 --streams 1 --discard --no-cpu --no-compare > $O/b_c3_s1_discard.json 2> $O/b_c3_s1_discard.err
timeout 300 python bench.py --workload c2 --streams 1 --no-cpu --no-compare > $O/b_c2_s1.json 2> $O/b_c2_s1.err
timeout 300 python bench.py --workload c2 --streams 1 --group 8 --no-cpu --no-compare > $O/b_c2_s1_g8.json 2> $O/b_c2_s1_g8.err
timeout 300 python bench.py --workload c2 --streams 3 --group 8 --no-cpu --no-compare > $O/b_c2_s3_g8.json 2> $O/b_c2_s3_g8.err
# DRAM bytes per kernel with the L2 state carried from kernel to kernel (no flush between ncu's launches; one pass, no replay)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read.sum --cache-control none --clock-control none -s 30 -c 60 --csv --log-file $O/l2carry_c3_s1_r2a.csv \
    python bench.py --workload c3 --streams 1 --profile --steps 1 --warmup 1 > $O/ncu_l2carry.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -s 30 -c 60 --csv --log-file $O/l2carry_c2_g8_r2a.csv \
    python bench.py --workload c2 --streams 1 --group 8 --profile --steps 1 --warmup 1 > $O/ncu_l2carry_c2.log 2>&1
python tools/summ.py default_r2a c3_s1 c3_s2 c3_s3 c3_s1_discard c2_s1 c2_s1_g8 c2_s3_g8
ls -la $O | tail -30
