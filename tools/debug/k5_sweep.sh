#!/bin/bash
# usage: k5_sweep.sh "<EXTRA flags>" ...   (GPU box, repo root): K5 field-decode times for each build
for cfg in "$@"; do
  rm -f build/fields.cu.o
  make EXTRA="$cfg" bam_parallel_b200/lib/libbamgpu.so > /dev/null 2>&1
  timeout 300 python bench.py --warmup 1 --steps 1 --no-e2e --no-cpu-baseline 2> gpurun_out/sweep.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); f=d['fields']; print('CFG', '$cfg', f['ms_measure'], f['ms_decode'])"
done
