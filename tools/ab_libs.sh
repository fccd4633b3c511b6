#!/bin/bash
# A/B of alternative builds of the library (same ABI, B200MDM_LIB=<path>) inside ONE gpurun call: the GEMM kernel tests of
# each build, loop times interleaved twice (only same-box numbers compare), one ncu --set full capture of the FFN-up GEMM
# of the default build.   usage: tools/ab_libs.sh <tag> <lib name under lib/> [...]
tag=${1:-r02n}; shift
libdir=motion-diffusion-model_b200/lib
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/${tag}_$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-3} gpurun_out/${tag}_$name.log | cut -c1-300; return $rc; }
for l in "$@"; do
  B200MDM_LIB=$PWD/$libdir/$l TMO=240 TAILN=4 run kernel_tests_${l%.so} python -m pytest tests/test_kernels_gpu.py -q -x -k gemm_tcgen05
done
for rep in 1 2; do
  for l in "$@"; do B200MDM_LIB=$PWD/$libdir/$l TMO=120 TAILN=1 run time_loop_${l%.so}_$rep python tools/time_loop.py 7; done
done
TMO=150 TAILN=2 run ncu_gemm2w ncu --set full --cache-control none --clock-control none --import-source on -k regex:gemm2w -s 4 -c 2 \
  -f -o gpurun_out/${tag}_gemm2w_f16_tcgen05 python tools/profile_step.py 2
