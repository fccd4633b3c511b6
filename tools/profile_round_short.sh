#!/bin/bash
# tools/profile_round.sh without the ncu --set full captures and the phase traces (for a re-run after a change that
# leaves the hot kernels untouched): tests, smoke, loop timings, every bench line, reference arm, launch list.
tag=${1:-r02k}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/${tag}_$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-3} gpurun_out/${tag}_$name.log | cut -c1-400; }
TMO=900 TAILN=45 run gpu_tests python -m pytest tests -q -m gpu -s
run smoke python __graft_entry__.py smoke
TMO=120 run time_loop python tools/time_loop.py 9
TMO=120 run time_dip python tools/time_dip.py
TMO=300 run bench_c2 python bench.py --steps 10 --warmup 3
TMO=400 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
for c in c3 dip a2m; do TMO=300 run bench_$c python bench.py --config $c --steps 3 --warmup 3; done
TMO=300 run ncu_list ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -s 120 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 2
