"""K6, the compressing BGZF writer (csrc/deflate.cu): what it writes must be read back, byte for byte, by zlib
(every member a gzip stream: CRC32 + ISIZE checked), by the oracle's restatement of the reference's framing + inflate
(util.rs:10-71), and by this repository's own K1 + K2; the output is the same bytes on every run; incompressible payloads fall
back to stored blocks and never grow beyond the stored bound."""
import gzip
import io
import os
import random

import numpy as np
import pytest

import bam_parallel_b200 as bp
from bam_parallel_b200 import synth
from bam_parallel_b200.sorting import SortBy, sort_bam
from tests import bamcraft as bc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    e = bp.Engine()
    yield e
    e.close()


def _payloads():
    rnd = random.Random(3)
    yield "empty", b""
    yield "one byte", b"Q"
    yield "three bytes", b"abc"
    yield "zeros", bytes(200000)                                              # dist-1 runs, 258-byte matches across chunk and payload edges
    yield "period 3", b"abc" * 50000
    yield "text", b"".join(b"read%07d\tchr%d\t%d\tFFFFFFFFFF:FFFFF,FFFF\n" % (i, i % 24, i * 37 % 100000) for i in range(6000))
    yield "random", bytes(rnd.randrange(256) for _ in range(150000))           # incompressible: stored fallback
    yield "high bytes", bytes(rnd.choice((200, 201, 255, 144)) for _ in range(70000))   # 9-bit literals
    yield "far matches", (bytes(rnd.randrange(256) for _ in range(40000)) * 4)[:0xFF00 * 2 + 17]   # period 40000 > the 32 KiB window
    for n in (0xFF00 - 1, 0xFF00, 0xFF00 + 1, 2 * 0xFF00):
        yield f"edge {n}", (b"ACGT" * 17 + b"\x00\x01" * 5) * (n // 78) + b"z" * (n % 78)


@pytest.mark.parametrize("level", [1, 2])
@pytest.mark.parametrize("name,payload", list(_payloads()), ids=[n for n, _ in _payloads()])
def test_bgzf_write_round_trips(oracle, engine, name, payload, level):
    """level 1: the fixed Huffman code; level 2: a code built per block (or the fixed one, or a stored block: the smallest)."""
    for eof in (True, False):
        z = engine.bgzf_write(payload, level=level, eof_marker=eof)
        assert z == engine.bgzf_write(payload, level=level, eof_marker=eof)    # deterministic
        assert len(z) <= len(engine.bgzf_store(payload, eof_marker=eof))       # never larger than the stored form
        assert (gzip.decompress(z) if z else b"") == payload
        assert z.endswith(bc.BGZF_EOF) == eof or (not eof and not payload)
        blocks, nb, consumed, total, rc = bp.scan_blocks(z)
        assert consumed == len(z) and total == len(payload)
        if z:
            u, sizes = oracle.inflate_all(z)
            assert u.tobytes() == payload and (sizes.size == 0 or int(sizes.max()) <= 0xFF00)
    if name in ("zeros", "period 3", "text"):
        assert len(z) < len(payload) // 3, (name, len(z), len(payload))
    if name == "random":
        assert len(z) > len(payload)                                           # stored blocks + framing
    if level == 2:
        assert len(z) <= len(engine.bgzf_write(payload, level=1, eof_marker=False))   # never worse than the fixed code
        if name == "high bytes":
            assert len(z) < 0.34 * len(payload)                                 # four symbols: two bits each


@pytest.mark.parametrize("shape,n,payload", [("pe150", 60000, 0xFF00), ("long10k", 300, 4096), ("longname", 6000, 0xFF00)])
def test_compressed_bam_output_of_the_sort(oracle, shape, n, payload):
    """bamgpu_sort_bam(sort_output = BAM, out_compr_level = 1): header + sorted records in compressed BGZF; this repository's Reader
    (K1 + CRC check on the GPU) and zlib read back what the stored-block writer would have held."""
    data = synth.generate(shape, n, level=6, payload=payload).tobytes()
    for sb in (SortBy.CoordinatesAndStrand, SortBy.Name):
        stored, packed, fixed = io.BytesIO(), io.BytesIO(), io.BytesIO()
        sort_bam(0, data, stored, sort_by=sb, output="bam")
        st = sort_bam(0, data, packed, 0, 2, sort_by=sb, output="bam")         # out_compr_level is the 5th argument (sort.rs:143)
        sort_bam(0, data, fixed, 0, 1, sort_by=sb, output="bam")
        want = gzip.decompress(stored.getvalue())
        z = packed.getvalue()
        assert gzip.decompress(z) == want and z.endswith(bc.BGZF_EOF)
        assert gzip.decompress(fixed.getvalue()) == want
        print(f"{shape} {sb.name}: stored {len(stored.getvalue())}, fixed code {len(fixed.getvalue())}, per-block codes {len(z)}, "
              f"zlib level 6 input {len(data)}")
        assert len(z) < 0.75 * len(stored.getvalue()) and len(z) < len(fixed.getvalue()), (len(z), len(fixed.getvalue()), len(stored.getvalue()))
        assert st["ms_bgzf_write"] > 0
        with bp.Reader.new(z, 4) as r:
            hdr, _ = r.read_header()
            got = b"".join(len(x).to_bytes(4, "little") + bytes(x) for x in r.records())
        assert b"BAM\x01" + hdr + got == want
        # the streamed form and the sort over several GPUs write the same file
        for kw in ({"device_mem_limit": 4 << 20}, {"devices": [0, 0, 0]}):
            other = io.BytesIO()
            sort_bam(0, data, other, 0, 2, sort_by=sb, output="bam", **kw)
            assert other.getvalue() == z, kw
