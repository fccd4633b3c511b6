#!/bin/bash
# Final evidence run of a round inside ONE gpurun call, most important first (the call may be cut by the GPU budget):
# full GPU test suite, bench line of the headline config, DiP chunk times of the default build and of alternative builds
# given on the command line, ncu launch list, DiP bench line, smoke.   usage: tools/profile_final.sh <tag> [alt lib names]
tag=${1:-r02p}; shift
libdir=motion-diffusion-model_b200/lib
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/${tag}_$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-3} gpurun_out/${tag}_$name.log | cut -c1-400; return $rc; }
TMO=600 TAILN=40 run gpu_tests python -m pytest tests -q -m gpu -s --durations=12
TMO=200 TAILN=1 run bench_c2 python bench.py --steps 10 --warmup 3
TMO=100 TAILN=1 run time_dip python tools/time_dip.py
for l in "$@"; do B200MDM_LIB=$PWD/$libdir/$l TMO=100 TAILN=1 run time_dip_${l%.so} python tools/time_dip.py; done
TMO=100 TAILN=1 run time_loop python tools/time_loop.py 9
TMO=200 TAILN=2 run ncu_list ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -s 120 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 2
TMO=150 TAILN=1 run bench_dip python bench.py --config dip --steps 3 --warmup 3
TMO=100 TAILN=2 run smoke python __graft_entry__.py smoke
TMO=200 TAILN=1 run bench_a2m python bench.py --config a2m --steps 3 --warmup 3
TMO=200 TAILN=1 run bench_c3 python bench.py --config c3 --steps 3 --warmup 3
