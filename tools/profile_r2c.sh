#!/bin/bash
# Last profiling session of round 2 (GPU box, through gpurun; every ncu pass follows a plain run of the same command).
export BAMGPU_BENCH_CACHE=/dev/shm/bamgpu_cache
set -x
python bench.py > gpurun_out/r2c_final.json 2> gpurun_out/r2c_final.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_reference.json 2> gpurun_out/r2c_reference.err
B="python bench.py --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches_50m.csv $B --steps 2 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_name_masks|k_name_group_key' -c 4 -f -o gpurun_out/r2c_name $B --steps 1 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
