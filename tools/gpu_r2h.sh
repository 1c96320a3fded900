#!/bin/bash
# Round-2 call H: do persistent row/column kernels with CAPPED grids (room left on every SM for the other kernels of frames in flight)
# overlap better than kernels that each fill the machine? + C4 shard-of-8 launch-shape probes.
O=gpurun_out
mkdir -p $O
timeout 900 python tools/sweep_modes.py c3 --rows 1 --cols 1:0 --streams 3,4 --groups 0,1 > $O/sweep_c3_r2h_base.txt 2>&1; cat $O/sweep_c3_r2h_base.txt
timeout 900 python tools/sweep_modes.py c3 --rows 2,3 --cols 2:0,3:0 --streams 3,4 --groups 0,1 --caps 0:0,1:1,2:1,3:1 > $O/sweep_c3_r2h.txt 2>&1; cat $O/sweep_c3_r2h.txt
timeout 600 python tools/sweep_modes.py c3 --rows 2,3 --cols 1:0 --streams 3,4 --groups 0,1 --caps 1:0,2:0 > $O/sweep_c3_r2h_b.txt 2>&1; cat $O/sweep_c3_r2h_b.txt
timeout 600 python tools/sweep_modes.py c4 --c4-shard-of 8 --rows 2 --cols 1:0 --streams 3,4 --groups 0,1,2,4 --reps 7 > $O/sweep_c4s8_r2h.txt 2>&1; cat $O/sweep_c4s8_r2h.txt
timeout 600 python tools/sweep_modes.py c4 --c4-shard-of 8 --rows 2 --cols 1:0 --streams 3,4 --groups 0,1,2,4 --reps 7 --graph > $O/sweep_c4s8_graph_r2h.txt 2>&1; cat $O/sweep_c4s8_graph_r2h.txt
