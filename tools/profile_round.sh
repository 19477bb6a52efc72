#!/bin/bash
# Round evidence in one GPU call: bench lines, microbenchmarks, the bench's ncu launch list and ncu --set full
# summaries (text only: the .ncu-rep files stay on the box) -> gpurun_out/
set -x
O=gpurun_out
python bench.py > $O/r1_bench_line.json 2> $O/bench.err
python bench.py --impl reference > $O/r1_bench_reference_arm.json 2>> $O/bench.err
python tools/microbench.py > $O/r1_microbench.jsonl 2> $O/mb.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/b.log 2>&1
T=/tmp/ncu_reps; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:bytetrack_step -s 4 -c 1 -o $T/bt -f python bench.py --no-e2e --no-cpu --steps 2 --warmup 8 --streams 148 --frames 20 > $O/ncu_bt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ocsort_step -s 2 -c 1 -o $T/oc -f python tools/microbench.py --only engine_ocsort --quick > $O/ncu_oc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:botsort_step -s 1 -c 1 -o $T/bot -f python tools/microbench.py --only engine_botsort --quick > $O/ncu_bot.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:strongsort_step -s 30 -c 1 -o $T/ss -f python tools/microbench.py --only engine_strongsort --quick > $O/ncu_ss.log 2>&1
ncu --set full --clock-control none -k regex:"cosine|ocm_cost|kf_|iou_cost|lap_dense|gate_cost|iou_tlwh|nn_" -c 60 -o $T/micro -f python tools/microbench.py --quick > $O/ncu_micro.log 2>&1
for k in bt:bytetrack_step_kernel oc:ocsort_step_kernel bot:botsort_step_kernel ss:strongsort_step_kernel; do
  f=${k%%:*}; n=${k##*:}
  { python tools/ncu_summary.py $T/$f.ncu-rep; echo "# per-source-line / per-phase breakdown (tools/ncu_lines.py) of the same capture"; python tools/ncu_lines.py $T/$f.ncu-rep motcpp_b200/libmotb200.so $n 40; } > $O/r1_${n}_ncu_full.txt 2>&1
done
python tools/ncu_summary.py $T/micro.ncu-rep --last > $O/r1_microkernels_ncu_full.txt 2>&1
ls -la $O
