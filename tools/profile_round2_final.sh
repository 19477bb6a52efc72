#!/bin/bash
# Round-2 evidence refresh (one gpurun call): bench lines (N = 1, reference arm, BASELINE configs[4] in both readings),
# ncu launch list of the bench command, ncu --set full of bytetrack_step_kernel and boosttrack_step_kernel, the event-timed
# microbenchmarks, smoke().  Text / JSON only -> gpurun_out/
O=gpurun_out; T=/tmp/ncu_reps; mkdir -p $T $O
python bench.py > $O/r2_bench_line.json 2> $O/bench.err
python bench.py --impl reference > $O/r2_bench_reference_arm.json 2>> $O/bench.err
python bench.py --workload c5 --no-cpu > $O/r2_bench_c5_1gpu_64streams.json 2>> $O/bench.err
python bench.py --streams 8 --no-cpu > $O/r2_bench_c5_8streams_per_gpu.json 2>> $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
bash tools/profile_bt.sh r2_bytetrack_step_ncu_full.txt
ncu --set full --clock-control none --import-source on -k regex:"boosttrack_step" -s 2 -c 1 -o $T/boost -f python tools/microbench.py --only engine_boosttrack > $O/ncu_boost.log 2>&1
{
  echo "# boosttrack_step_kernel<1536,512,4096>, 296 streams x 25 frames, steady state (python tools/microbench.py --only engine_boosttrack under ncu --set full)"
  python tools/ncu_summary.py $T/boost.ncu-rep
  echo "# per-source-line breakdown (tools/ncu_lines.py) of the same capture"
  python tools/ncu_lines.py $T/boost.ncu-rep motcpp_b200/libmotb200.so boosttrack_step_kernel 30
} > $O/r2_boosttrack_step_ncu_full.txt 2>&1
python tools/microbench.py > $O/r2_microbench.jsonl 2> $O/mb.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1
tail -c 300 $O/r2_bench_line.json; tail -2 $O/r2_smoke.txt; wc -l $O/r2_microbench.jsonl; head -12 $O/r2_bytetrack_step_ncu_full.txt
