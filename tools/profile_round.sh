#!/bin/bash
# Round-end evidence run (one B200): tests, smoke, bench lines of every BASELINE config, reference CPU arm, launch list,
# `ncu --set full` per hot kernel, phase traces (instrumented build).  Outputs under gpurun_out/<tag>_*; summaries are
# copied into profiles/ by hand (tools/launch_summary.py, tools/ncu_summary.py, tools/sass_listing.py).
tag=${1:-r02f}
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
tools/ncu_kernels.sh $tag
TMO=120 run trace_qkv_attn python tools/trace_qkv_attn.py
TMO=120 run trace_ln python tools/trace_ln.py
