#!/bin/bash
# ncu --set full of one bytetrack_step_kernel launch (C2 workload, 296 streams x 20 frames = the bench's occupancy) ->
# summary + per-source-line breakdown in gpurun_out/$1 (text only; the .ncu-rep stays on the box)
O=gpurun_out; T=/tmp/ncu_reps; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:bytetrack_step -s 4 -c 1 -o $T/bt -f python bench.py --no-e2e --no-cpu --steps 2 --warmup 8 --streams 296 --frames 20 > $O/ncu_bt.log 2>&1
{ python tools/ncu_summary.py $T/bt.ncu-rep; echo "# per-source-line / per-phase breakdown (tools/ncu_lines.py) of the same capture"; python tools/ncu_lines.py $T/bt.ncu-rep motcpp_b200/libmotb200.so bytetrack_step_kernel 60; } > $O/$1 2>&1
