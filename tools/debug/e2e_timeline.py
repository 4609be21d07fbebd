"""Stand-alone e2e loop over the Reader (page-locked source), for timelines and sweeps:
    BAMGPU_DEBUG=2 python tools/debug/e2e_timeline.py [records] [flags: bodies|cols|all] [devices]
BAMGPU_JOBS / BAMGPU_BATCH_MIB (read by the library) sweep the jobs per GPU and the batch size."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bam_parallel_b200 as bp
from bam_parallel_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "bodies"
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
flags = {"bodies": bp.COPY_DATA | bp.COPY_OFFSETS, "cols": bp.COPY_COLUMNS, "all": bp.COPY_DATA | bp.COPY_COLUMNS}[mode]
cache = os.environ.get("BAMGPU_BENCH_CACHE")
path = f"{cache}.pe150.{n}.L6.bam" if cache else None
if path and os.path.exists(path):
    arr = np.fromfile(path, dtype=np.uint8)
else:
    bam = synth.generate("pe150", n, level=6, header_block=True)
    arr = bam.array
    if path:
        arr.tofile(path)
L = bp.load()
p = L.bamgpu_pinned_alloc(arr.size)
C.memmove(p, arr.ctypes.data, arr.size)
parr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(arr.size,))
dbg = os.environ.pop("BAMGPU_DEBUG", None)
kw = {"devices": list(range(ndev))} if ndev > 1 else {}
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if it == 4 and dbg: os.environ["BAMGPU_DEBUG"] = dbg
    with bp.Reader.new(parr, 16, flags=flags, **kw) as r:
        t1 = time.perf_counter(); r.read_header(); t2 = time.perf_counter()
        nrec = 0; nb = 0
        while True:
            b = r.next_batch()
            if b is None: break
            nrec += b.n_records; nb += 1
        t3 = time.perf_counter()
        st = r.stats()
    t4 = time.perf_counter()
    print(f"iter {it}: new {1e3*(t1-t0):.1f} header {1e3*(t2-t1):.1f} batches {1e3*(t3-t2):.1f} close {1e3*(t4-t3):.1f} total {1e3*(t4-t0):.1f} ms"
          f" | {nb} batches, K1 {st['ms_inflate']:.0f} K2 {st['ms_crc']:.0f} K3 {st['ms_records']:.0f} D2H {st['ms_d2h']:.0f} ms (device, summed)"
          f" | {16.907 * n / 50e6 / (t4 - t0):.1f} GB/s", file=sys.stderr)
