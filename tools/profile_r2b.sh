#!/bin/bash
# Final round-2 profiling session (GPU box, through gpurun; every ncu pass follows a plain run of the same command).
export BAMGPU_BENCH_CACHE=/dev/shm/bamgpu_cache
B="python bench.py --warmup 3 --no-e2e --no-cpu-baseline"
set -x
$B --steps 3 > gpurun_out/r2b_plain.json 2> gpurun_out/r2b_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_50m.csv $B --steps 2 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k5_|k_name_chunk|k_gather_bodies|k3_' -s 12 -c 14 -f -o gpurun_out/r2b_k345 $B --steps 1 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
