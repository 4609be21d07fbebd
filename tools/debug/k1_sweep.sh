#!/bin/bash
# usage: sweep.sh "<EXTRA flags>" ...   (run on the GPU box from the repo root)
for cfg in "$@"; do
  rm -f build/inflate.cu.o
  make EXTRA="$cfg" bam_parallel_b200/lib/libbamgpu.so > /dev/null 2>&1
  grep -A3 "k1_inflate" build/inflate.ptxas.log | grep -E "spill|Used" | tr '\n' ' '
  echo
  BAMGPU_DEBUG=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline 2> gpurun_out/sweep.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CFG', '$cfg', round(d['value'],1), d['roofline']['stage_ms_per_step'])"
  grep bamgpu gpurun_out/sweep.err | head -1
done
