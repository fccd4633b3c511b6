#!/bin/bash
# One `ncu --set full` capture per hot kernel of the sampling step (BASELINE config 2 shapes, plain launches, warm caches:
# --cache-control none keeps the L2 state of the real loop).  Reports land in gpurun_out/<tag>_<kernel>.ncu-rep.
tag=${1:-r02}
mkdir -p gpurun_out
for k in qkv_attention_kernel gemm_resid_ln_cluster gemm2w_f16_tcgen05 "gemm_f16_tcgen05"; do
  timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:$k -s 4 -c 4 \
      -f -o gpurun_out/${tag}_$k python tools/profile_step.py 2 > gpurun_out/${tag}_$k.log 2>&1
  echo "$k exit $?"
done
