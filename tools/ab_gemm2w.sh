#!/bin/bash
# A/B of the W-resident pair GEMM (csrc/gemm2w.cuh, B200MDM_GEMM2W=0/1) inside ONE gpurun call: kernel tests, loop times of
# both settings interleaved (boxes differ in how hard the power cap bites: only same-box numbers compare), DiP chunk
# times (the loop checksum printed by time_loop.py must not change: both kernels accumulate in the same order), and one
# ncu --set full capture of the new kernel.
tag=${1:-r02m}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/${tag}_$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-3} gpurun_out/${tag}_$name.log | cut -c1-300; return $rc; }
TMO=240 TAILN=12 run kernel_tests python -m pytest tests/test_kernels_gpu.py -q -x -k gemm_tcgen05 || { echo "kernel tests failed: stop"; exit 1; }
for rep in 1 2; do
  for v in 0 1; do B200MDM_GEMM2W=$v TMO=120 TAILN=1 run time_loop_w${v}_$rep python tools/time_loop.py 7; done
done
for v in 0 1; do B200MDM_GEMM2W=$v TMO=120 TAILN=2 run time_dip_w$v python tools/time_dip.py; done
TMO=150 TAILN=2 run ncu_gemm2w ncu --set full --cache-control none --clock-control none --import-source on -k regex:gemm2w -s 4 -c 2 \
  -f -o gpurun_out/${tag}_gemm2w_f16_tcgen05 python tools/profile_step.py 2
