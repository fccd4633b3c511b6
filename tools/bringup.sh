#!/bin/bash
# GPU bring-up: every stage in its own process under its own timeout so that a trapped kernel cannot hide the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
(nproc; cat /sys/fs/cgroup/cpu.max; lscpu | head -20; free -g | head -2) > gpurun_out/host.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run kernels python -m pytest tests/test_kernels_gpu.py -q
run parity python -m pytest tests/test_parity_gpu.py -q -s
TMO=300 run bench python bench.py --steps 3 --warmup 3
TMO=300 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 300 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2
