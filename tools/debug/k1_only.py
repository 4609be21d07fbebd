"""K1-only timing (no CRC, no records): python tools/debug/k1_only.py [records]  — for kernel experiments, incl. the
BG_DIAG_* builds whose output is wrong on purpose."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bam_parallel_b200 as bp
from bam_parallel_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cache = os.environ.get("BAMGPU_BENCH_CACHE", "/dev/shm/bamgpu_cache") + f".pe150.{n}.L6.bam"
if os.path.exists(cache):
    arr = np.fromfile(cache, dtype=np.uint8)
else:
    bam = synth.generate("pe150", n, level=6, header_block=True)
    arr = np.array(bam.array, copy=True)
    arr.tofile(cache)
eng = bp.Engine()
blocks, nb, consumed, total, rc = bp.scan_blocks(arr)
d = eng.upload(arr)
for it in range(6):
    eng.reset_stats()
    try:
        eng.inflate(d, blocks, nb, flags=bp.NO_CRC)
        eng.data()
    except Exception as e:
        print("  (error, expected for BG_DIAG builds):", str(e)[:80])
    print("K1_ONLY", os.environ.get("TAG", ""), nb, "blocks", round(eng.stats()["ms_inflate"], 2), "ms")
