#!/bin/bash
# usage: k1_diag.sh "<EXTRA flags>" ...   (GPU box, repo root): K1-only time of each build (tools/debug/k1_only.py)
for cfg in "$@"; do
  rm -f build/inflate.cu.o
  make EXTRA="$cfg" bam_parallel_b200/lib/libbamgpu.so > /dev/null 2>&1
  TAG="$cfg" timeout 300 python tools/debug/k1_only.py 2>&1 | tail -2
done
