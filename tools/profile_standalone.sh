#!/bin/bash
# ncu --set full summaries of every stand-alone kernel behind the C ABI (north_star: "each evidenced by a committed ncu
# capture"): tensor-core cosine GEMM (dense and nearest-neighbour modes) + split, IoU / OCM / gate / tlwh / variant / aw
# cost kernels, the sparse assignment kernel, both reference-order LAPJV kernels, the Kalman kernels, the row compaction.
# Text only (the .ncu-rep files stay on the box) -> gpurun_out/r2_standalone_kernels_ncu_full.txt
set -x
O=gpurun_out; T=/tmp/ncu_reps; mkdir -p $T
K='cosine_gemm|cosine_split|nn_fill|nn_decode|iou_cost|ocm_cost|gate_cost|iou_tlwh|iou_variant|aw_|lap_dense|lap_jv|kf_|pack_'
ncu --set full --clock-control none -k regex:"$K" -c 260 -o $T/micro -f python tools/microbench.py --quick > $O/ncu_micro.log 2>&1
ncu --set full --clock-control none -k regex:"cosine_gemm|cosine_split" -c 30 -o $T/cos -f python tools/microbench.py --only cos > $O/ncu_cos.log 2>&1
ncu --set full --clock-control none -k regex:"pack_" -c 4 -o $T/pack -f python bench.py --no-cpu --steps 1 --warmup 1 > $O/ncu_pack.log 2>&1
{
  echo "# python tools/microbench.py --quick under ncu --set full (last capture of every kernel name)"
  python tools/ncu_summary.py $T/micro.ncu-rep --last
  echo "# cosine_gemm_kernel / cosine_split_kernel at 1024x1024x512 and 4096x4096x512 (python tools/microbench.py --only cos; every launch listed)"
  python tools/ncu_summary.py $T/cos.ncu-rep
  echo "# row compaction of the packed host path (bench.py e2e leg)"
  python tools/ncu_summary.py $T/pack.ncu-rep --last
} > $O/r2_standalone_kernels_ncu_full.txt 2>&1
python tools/microbench.py > $O/r2_microbench.jsonl 2> $O/mb.err
grep -c "kernel:" $O/r2_standalone_kernels_ncu_full.txt
