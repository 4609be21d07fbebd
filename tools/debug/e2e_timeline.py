import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bam_parallel_b200 as bp
from bam_parallel_b200 import synth
bam = synth.generate("pe150", 50_000_000, level=6, header_block=True)
arr = bam.array
for it in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    if it==2: os.environ["BAMGPU_DEBUG"]="2"
    with bp.Reader.new(arr, 16, flags=bp.COPY_DATA|bp.COPY_OFFSETS, batch_bytes=256<<20) as r:
        t1=time.perf_counter(); r.read_header(); t2=time.perf_counter()
        n=0
        while True:
            b=r.next_batch()
            if b is None: break
            n+=b.n_records
        t3=time.perf_counter()
    t4=time.perf_counter()
    print(f"iter {it}: new {1e3*(t1-t0):.1f} header {1e3*(t2-t1):.1f} batches {1e3*(t3-t2):.1f} close {1e3*(t4-t3):.1f} total {1e3*(t4-t0):.1f} ms", file=sys.stderr)
