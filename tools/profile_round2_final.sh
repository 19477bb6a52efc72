#!/bin/bash
# Round-2 evidence refresh (one gpurun call): bench lines (N = 1, reference arm, BASELINE configs[4] in both readings),
# the event-timed microbenchmarks, smoke().  Text / JSON only -> gpurun_out/
O=gpurun_out
python bench.py > $O/r2_bench_line.json 2> $O/bench.err
python bench.py --impl reference > $O/r2_bench_reference_arm.json 2>> $O/bench.err
python bench.py --workload c5 --no-cpu > $O/r2_bench_c5_1gpu_64streams.json 2>> $O/bench.err
python bench.py --streams 8 --no-cpu > $O/r2_bench_c5_8streams_per_gpu.json 2>> $O/bench.err
python tools/microbench.py > $O/r2_microbench.jsonl 2> $O/mb.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1
tail -c 300 $O/r2_bench_line.json; tail -2 $O/r2_smoke.txt; wc -l $O/r2_microbench.jsonl
