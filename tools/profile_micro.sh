#!/bin/bash
# ncu --set full summaries of the stand-alone kernels (tensor-core cosine GEMM in both modes, cost kernels, LAP) -> gpurun_out/
# (tools/profile_round.sh covers the engines; its micro capture runs out of launches before it reaches these).
set -x
O=gpurun_out; T=/tmp/ncu_reps; mkdir -p $T
ncu --set full --clock-control none -k regex:"cosine_gemm|cosine_split" -c 52 -o $T/cos -f python tools/microbench.py --only cos > $O/ncu_cos.log 2>&1
ncu --set full --clock-control none -k regex:"cosine_gemm|nn_fill|nn_decode" -c 12 -o $T/nn -f python tools/microbench.py --only nn_cos --quick > $O/ncu_nn.log 2>&1
{
  echo "# cosine_gemm_kernel / cosine_split_kernel: python tools/microbench.py --only cos (1024x1024x512, then 4096x4096x512; every launch listed)"
  python tools/ncu_summary.py $T/cos.ncu-rep --last
  echo "# nearest-neighbour mode (mot_cost_nn_cosine, 256 targets x 100 samples x 512 detections x 512-d)"
  python tools/ncu_summary.py $T/nn.ncu-rep --last
} > $O/r1_standalone_kernels_ncu_full.txt 2>&1
grep -n "kernel:\|tensor_cycles\|time_duration\|dram_throughput" $O/r1_standalone_kernels_ncu_full.txt | head -80
