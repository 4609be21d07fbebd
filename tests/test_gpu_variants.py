"""K1 variant 1 (bamgpu_opts.inflate_impl = 1: one warp per BGZF block, the north-star's mapping, csrc/inflate_warp.cu)
must produce exactly what the default kernel and the oracle produce: the same parity cases are run through it."""
import numpy as np
import pytest

import bam_parallel_b200 as bp
from bam_parallel_b200 import synth
from tests import test_gpu_parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine_v1():
    e = bp.Engine(inflate_impl=1)
    yield e
    e.close()


@pytest.mark.parametrize("shape,n,payload", [("se150", 20000, 0xFF00), ("pe150", 20000, 0xFF00), ("longname", 8000, 0xFF00),
                                             ("long10k", 300, 4096), ("long10k", 300, 0xFF00), ("pe150", 1500, 300)])
@pytest.mark.parametrize("level", [1, 6, 9])
def test_variant1_parity_synth(oracle, engine_v1, shape, n, payload, level):
    data = synth.generate(shape, n, level=level, payload=payload).tobytes()
    P._assert_parity(oracle, engine_v1, data)


def test_variant1_block_kinds_overlaps_and_corruption(oracle, engine_v1):
    P.test_deflate_block_kinds_and_overlaps(oracle, engine_v1)
    P.test_many_small_unaligned_blocks(oracle, engine_v1)
    P.test_corrupt_blocks_are_reported(engine_v1)
    P.test_huge_record_spanning_many_tiles(oracle, engine_v1)


def test_variant1_through_the_reader(oracle):
    data = synth.generate("pe150", 60000).tobytes()
    u, hdr, off_refs, consumed, off, ln, cols, bf, bt = P._oracle_decode(oracle, data)
    import hashlib
    ub = u.tobytes()
    want = hashlib.md5(b"".join(ub[o:o + l] for o, l in zip(off.tolist(), ln.tolist()))).hexdigest()
    with bp.Reader.new(data, 4, batch_bytes=1 << 20, inflate_impl=1) as r:
        r.read_header()
        assert P._reader_md5(r) == (want, off.size)
