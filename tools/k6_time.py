"""Times the BGZF writers (K6) on the name-sorted bodies of a synthetic BAM that is already in HBM.
usage: python tools/k6_time.py [records] [reps]   (run under ncu for the kernel profile: tools/profile_k6.sh)"""
import ctypes as C
import sys

import bam_parallel_b200 as bp
from bam_parallel_b200 import _lib, synth
from bam_parallel_b200.sorting import SortBy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
data = synth.generate("pe150", n, level=6)
L = _lib.load()
eng = bp.Engine()
blocks, nb, _, _, _ = bp.scan_blocks(data.array)
with bp.Reader.new(data.array[:1 << 20].copy(), 4, flags=bp.NO_CRC, batch_bytes=1 << 20) as r0:
    hdr, _ = r0.read_header()
d = eng.upload(data.array)
eng.decode(d, blocks, nb, 4 + len(hdr), True, 0)
_, _, srt = eng.sort(int(SortBy.Name))
nbytes = int(srt.bodies_len)
bound = int(L.bamgpu_bgzf_bound(nbytes))
d_out = L.bamgpu_device_alloc(eng._h, bound)
for level in (0, 1, 2):
    for rep in range(reps):
        eng.reset_stats()
        wrote = C.c_size_t()
        rc = L.bamgpu_engine_bgzf_write(eng._h, srt.bodies, nbytes, d_out, bound, C.byref(wrote), 1, level, None)
        assert rc == 0, rc
        ms = eng.stats()["ms_bgzf_write"]
    print(f"level {level}: {ms:8.2f} ms  {nbytes / ms / 1e6:7.1f} GB/s in  ratio {nbytes / wrote.value:.3f}  ({nbytes} -> {wrote.value} bytes)")
L.bamgpu_device_free(eng._h, d_out)
eng.close()
