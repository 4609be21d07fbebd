#!/bin/bash
# Round-2 profiling session (run on the GPU box through gpurun; every ncu pass follows a plain run of the same command).
# Outputs land in gpurun_out/; the summaries that are judged are copied to profiles/ afterwards (tools/ncu_summary.py).
export BAMGPU_BENCH_CACHE=/dev/shm/bamgpu_cache
B="python bench.py --warmup 3 --no-e2e --no-cpu-baseline"
set -x
$B --steps 3 > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
BAMGPU_INFLATE_IMPL=1 $B --steps 3 --no-sort > gpurun_out/r2_plain_v1.json 2> gpurun_out/r2_plain_v1.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_50m.csv $B --steps 2 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_inflate_pairs -s 3 -c 1 -f -o gpurun_out/r2_k1 $B --steps 1 --no-sort > gpurun_out/ncu_b.log 2>&1
BAMGPU_INFLATE_IMPL=1 ncu --set full --clock-control none --import-source on -k regex:k1_inflate_warp -s 3 -c 1 -f -o gpurun_out/r2_k1v1 $B --steps 1 --no-sort > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none -k regex:'k3_|k2_crc' -s 12 -c 5 -f -o gpurun_out/r2_k23 $B --steps 1 --no-sort > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
