"""GPU parity at BASELINE.json's OWN sizes (configs[1], [3], [4]), through digests.

A 50M-record decode produces 16.9 GB of bodies and 3.7 GB of columns; instead of hauling them over PCIe both sides
fold every value into a few 64-bit words with the same published function (DIGEST SPEC: oracle/bam_oracle.c for the
CPU restatement of the reference, bam_parallel_b200/csrc/digest.cu for the CUDA path, which digests ITS columns and
buffers on the device).  Per record: index, stream offset of the body, block_size, the 11 fixed fields
(bamrawrecord.rs:86-97), the 4 variable-field offsets (:61-77), coord_key + strand (comparators.rs:41-59,173-180), HI
(tags.rs:113-129) and every body byte; for the sorts (sort.rs:138-178, comparators.rs:61-147) the permutation and the
gathered output bytes.  Any differing byte, field, boundary or position in the order changes a digest.

Sizes can be scaled for a quick run with BAMGPU_FULLSIZE_SCALE (default 1.0 = the BASELINE sizes).
"""
import os
import time

import numpy as np
import pytest

import bam_parallel_b200 as bp
from bam_parallel_b200 import synth

pytestmark = pytest.mark.gpu

SCALE = float(os.environ.get("BAMGPU_FULLSIZE_SCALE", "1.0"))
THREADS = min(os.cpu_count() or 4, 32)
FIELDS = {"cigar": 1, "seq": 2, "qual": 4}
MODES = {"coord": (bp.SortBy.CoordinatesAndStrand, 0), "name": (bp.SortBy.Name, 1), "name_mates": (bp.SortBy.NameAndMatchMates, 2)}


def _n(n):
    return max(1000, int(n * SCALE))


def _oracle_side(oracle, arr, sort_modes):
    """The reference restated on the CPU (several threads): digests of the decode and of each requested sort."""
    t0 = time.time()
    u = oracle.inflate_all_mt(arr, THREADS)
    _, _, start = oracle.read_header(u)
    off, ln = oracle.walk_records(u, start)
    cols, bad_flags, bad_tags = oracle.extract_mt(u, off, ln, True, THREADS)
    assert bad_flags == 0 and bad_tags == 0
    out = {"start": start, "ulen": int(u.size), "decode": oracle.digest(u, off, ln, cols, THREADS)}
    for name, fid in FIELDS.items():       # decode_cigar / decode_seq / RawQual over every record (bamrawrecord.rs:126-157,208-273)
        out["field_" + name] = oracle.field_digest(u, off, ln, fid, THREADS)
    for m in sort_modes:
        perm = oracle.sort_perm_mt(u, off, cols, MODES[m][1], THREADS)
        out[m] = oracle.sorted_digest(u, off, ln, perm, THREADS)
        del perm
    del cols, off, ln
    oracle.release(u)
    out["seconds"] = time.time() - t0
    return out


def _gpu_engine_side(arr, start, sort_modes):
    """Whole file as one batch through bamgpu_engine_decode, digests from the device; then each sort."""
    eng = bp.Engine()
    try:
        blocks, nb, consumed, total, rc = bp.scan_blocks(arr)
        assert consumed == arr.size
        d = eng.upload(arr)
        batch, tail, rc = eng.decode(d, blocks, nb, start, True, bp.WANT_HI)
        assert tail == total
        out = {"decode": eng.digest(), "n": batch.n_records, "ulen": total}
        fd = eng.decode_fields(7, 0)
        for name, fid in FIELDS.items():
            out["field_" + name] = dict(eng.field_digest(fid), bad=fd[name]["bad"])
        for m in sort_modes:
            eng.sort(int(MODES[m][0]), flags=0)
            out[m] = eng.sorted_digest()
        return out
    finally:
        eng.close()


def _gpu_reader_side(arr, **kw):
    """The drop-in Reader (batches, carry between batches, jobs in flight): sum of the per-batch digests."""
    tot = {"n_records": 0, "body_bytes": 0, "fields": 0, "bodies": 0}
    with bp.Reader.new(arr, 8, flags=bp.WANT_HI, **kw) as r:
        r.read_header()
        nb = 0
        while True:
            b = r.next_batch()
            if b is None:
                break
            d = r.batch_digest(b)
            for k in tot:
                tot[k] = (tot[k] + d[k]) & 0xFFFFFFFFFFFFFFFF
            nb += 1
    return tot, nb


def _check(oracle, shape, n, level, payload, sort_modes, reader_kw=None):
    bam = synth.generate(shape, n, level=level, payload=payload)
    arr = bam.array
    ref = _oracle_side(oracle, arr, sort_modes)
    got = _gpu_engine_side(arr, ref["start"], sort_modes)
    assert got["ulen"] == ref["ulen"]
    assert got["decode"] == ref["decode"], "decode digest differs from the oracle"
    assert got["n"] == n
    for name in FIELDS:
        assert got["field_" + name] == ref["field_" + name], f"{name} field digest differs from the oracle"
    for m in sort_modes:
        assert got[m] == ref[m], f"{m} sort digest differs from the oracle"
    if reader_kw is not None:
        tot, nb = _gpu_reader_side(arr, **reader_kw)
        assert tot == ref["decode"], "Reader (batched) digest differs from the oracle"
    bam.close()
    return ref


def test_config2_and_4_pe150_50m(oracle):
    """configs[1]: 50M-record 2x150 bp decode; configs[3]: its coordinate+strand sort, and both name sorts.
    Mates are emitted in shuffled order, so -n and -n -M must differ."""
    ref = _check(oracle, "pe150", _n(50_000_000), 6, 0xFF00, ("coord", "name", "name_mates"), reader_kw={})
    assert ref["name"]["perm"] != ref["name_mates"]["perm"]
    assert ref["name"]["bodies"] != ref["name_mates"]["bodies"]


@pytest.mark.parametrize("level", [1, 6, 9])
def test_config5_longname_5m(oracle, level):
    """configs[4]: name sort with mate matching on 200-222-byte names with HI on every hit, levels 1/6/9."""
    ref = _check(oracle, "longname", _n(5_000_000), level, 0xFF00, ("name", "name_mates"), reader_kw={"batch_bytes": 64 << 20})
    assert ref["name"]["perm"] != ref["name_mates"]["perm"]


@pytest.mark.parametrize("n,level,payload", [(500_000, 6, 0xFF00), (500_000, 1, 4096), (100_000, 9, 4096), (100_000, 1, 0xFF00),
                                             (100_000, 9, 0xFF00), (100_000, 6, 4096)])
def test_config5_long10k(oracle, n, level, payload):
    """configs[4]: ~15 KB records spanning 1-4+ BGZF blocks (payload 4096 and 0xff00), levels 1/6/9; -n -M sort."""
    _check(oracle, "long10k", _n(n), level, payload, ("name_mates",), reader_kw={"batch_bytes": 128 << 20})


class _CrcSink:
    """io.Write for sort_bam that keeps a CRC-32 and a length instead of 17 GB of output."""
    def __init__(self):
        import zlib
        self._crc32, self.crc, self.n = zlib.crc32, 0, 0

    def write(self, b):
        self.crc = self._crc32(b, self.crc)
        self.n += len(b)


def test_sort_over_several_gpus_50m():
    """configs[3] through bamgpu_sort_bam: the sample sort over the box's GPUs (on one GPU: two block ranges and two buckets on it)
    must write the bytes of the single-GPU sort, whose order and bytes test_config2_and_4_pe150_50m compares with the oracle."""
    import torch
    from bam_parallel_b200.sorting import SortBy, sort_bam
    bam = synth.generate("pe150", _n(50_000_000), level=6)
    ng = torch.cuda.device_count()
    devices = list(range(ng)) if ng > 1 else [0, 0]
    for sb in (SortBy.CoordinatesAndStrand, SortBy.NameAndMatchMates):
        one, many = _CrcSink(), _CrcSink()
        sort_bam(0, bam.array, one, sort_by=sb)
        st = sort_bam(0, bam.array, many, sort_by=sb, devices=devices)
        assert st["batches"] >= len(devices), "the multi-GPU path did not run"
        assert (one.n, one.crc) == (many.n, many.crc), sb
    bam.close()


def test_compressed_bam_at_scale():
    """K6 on 10 M records: the name-sorted BAM written with per-block Huffman codes reads back (K1 + CRC check, batched Reader) to
    the digest of the same BAM written with stored blocks, and is about a third of its size."""
    from bam_parallel_b200.sorting import SortBy, sort_bam
    import io
    bam = synth.generate("pe150", _n(10_000_000), level=6)
    files = {}
    for level in (0, 2):
        sink = io.BytesIO()
        sort_bam(0, bam.array, sink, None, level, sort_by=SortBy.Name, output="bam")
        files[level] = np.frombuffer(sink.getbuffer(), dtype=np.uint8)
    bam.close()
    assert files[2].size < 0.4 * files[0].size, (files[2].size, files[0].size)
    d0, _ = _gpu_reader_side(files[0])
    d2, nb = _gpu_reader_side(files[2], batch_bytes=64 << 20)
    assert d0["n_records"] == _n(10_000_000) and d2 == d0 and nb > 3
