import sys
sys.path.insert(0, '.')
import bam_parallel_b200 as bp
from bam_parallel_b200 import synth
shape, n, payload, level = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
data = synth.generate(shape, n, level=level, payload=payload).tobytes()
e = bp.Engine()
blocks, nb, consumed, total, rc = bp.scan_blocks(data)
d = e.upload(data)
batch, tail, rc = e.decode(d, blocks, nb, 0, True, bp.COPY_DATA | bp.NO_RECORDS)
print("ok", nb, batch.data_len)
