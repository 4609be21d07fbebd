#!/bin/bash
# K6 profile: plain timing first, then one ncu --set full capture of the two compressor kernels (run under gpurun, one GPU).
set -e
PYTHONPATH=. python tools/k6_time.py 2000000 2 | tee gpurun_out/k6_time.txt
ncu --set full --clock-control none --import-source on -k regex:k6_deflate --launch-count 2 -o gpurun_out/k6 -f env PYTHONPATH=. python tools/k6_time.py 2000000 1 > gpurun_out/k6_ncu.log 2>&1
ncu -i gpurun_out/k6.ncu-rep --page raw --csv > gpurun_out/k6_raw.csv
ncu -i gpurun_out/k6.ncu-rep --page source --csv --kernel-name regex:k6_deflate_fixed > gpurun_out/k6_fixed_source.csv || true
