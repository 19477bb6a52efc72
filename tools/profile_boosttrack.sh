#!/bin/bash
# BoostTrack engine (second half): smoke, event-timed microbench line, ncu --set full of boosttrack_step_kernel,
# compute-sanitizer memcheck + racecheck over the BoostTrack part of tools/sanitize_smoke.py.  Text only -> gpurun_out/
O=gpurun_out; T=/tmp/ncu_reps; mkdir -p $T $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1
python tools/microbench.py --only engine_boosttrack > $O/r2_microbench_boosttrack.jsonl 2> $O/mb_boost.err
ncu --set full --clock-control none --import-source on -k regex:"boosttrack_step" -s 2 -c 1 -o $T/boost -f python tools/microbench.py --only engine_boosttrack > $O/ncu_boost.log 2>&1
{
  echo "# boosttrack_step_kernel<1536,512,4096>, 296 streams x 25 frames, steady state (python tools/microbench.py --only engine_boosttrack under ncu --set full)"
  python tools/ncu_summary.py $T/boost.ncu-rep
} > $O/r2_boosttrack_step_ncu_full.txt 2>&1
SAN_ONLY=boosttrack timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -5 > $O/r2_sanitizer_boosttrack_memcheck.txt
SAN_ONLY=boosttrack timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py 2>&1 | tail -12 > $O/r2_sanitizer_boosttrack_racecheck.txt
python -m pytest tests/test_abi.py -q -m gpu 2>&1 | tail -3
tail -2 $O/r2_smoke.txt; cat $O/r2_microbench_boosttrack.jsonl | cut -c1-400; tail -12 $O/r2_sanitizer_boosttrack_memcheck.txt $O/r2_sanitizer_boosttrack_racecheck.txt; head -30 $O/r2_boosttrack_step_ncu_full.txt
