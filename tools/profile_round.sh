#!/bin/bash
# Round-end evidence run (one B200): tests, bench line, launch list, full ncu capture of the top kernel.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run gpu_tests python -m pytest tests -q -m gpu
run smoke python __graft_entry__.py smoke
TMO=120 run time_loop python tools/time_loop.py 9
TMO=300 run bench python bench.py
TMO=300 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
TMO=300 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 300 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2
TMO=400 run ncu_full ncu --set full --clock-control none --import-source on -k regex:gemm_resid_ln_cluster -s 4 -c 2 -o gpurun_out/top_kernel python tools/profile_step.py 2
