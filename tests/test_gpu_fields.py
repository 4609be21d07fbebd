"""GPU parity of K5 (variable-length field decode) against the oracle's restatement of the reference's
accessors: decode_cigar over get_cigar (bamrawrecord.rs:126-157,223-236), decode_seq (:259-270), RawQual (:99) — through
the C ABI (bamgpu_engine_decode_fields / bamgpu_engine_field_digest)."""
import random
import struct

import numpy as np
import pytest

import bam_parallel_b200 as bp
from bam_parallel_b200 import synth
from tests import bamcraft as bc

pytestmark = pytest.mark.gpu
FIELDS = {"cigar": 1, "seq": 2, "qual": 4}


@pytest.fixture(scope="module")
def engine():
    e = bp.Engine()
    yield e
    e.close()


def _decode(engine, data, start):
    blocks, nb, consumed, total, rc = bp.scan_blocks(data)
    d = engine.upload(data)
    try:
        engine.decode(d, blocks, nb, start, True, 0)
        f = engine.decode_fields(7, bp.COPY_DATA)
        out = {k: dict(off=f[k]["off"].copy(), bytes=f[k]["bytes"].copy(), total=f[k]["total"], bad=f[k]["bad"]) for k in FIELDS}
        dig = {k: engine.field_digest(v) for k, v in FIELDS.items()}
    finally:
        engine.free(d)
    return out, dig, f["n_records"]


def _assert_fields(oracle, engine, data, expect_bad=None):
    u, _ = oracle.inflate_all(data)
    _, _, start = oracle.read_header(u)
    off, ln = oracle.walk_records(u, start)
    g, dig, n = _decode(engine, data, start)
    assert n == off.size
    for name, fid in FIELDS.items():
        ooff, obytes, obad = oracle.decode_field(u, off, ln, fid)
        assert np.array_equal(g[name]["off"], ooff), name
        assert g[name]["bytes"].tobytes() == obytes.tobytes(), name
        assert g[name]["bad"] == obad and g[name]["total"] == int(ooff[-1]), name
        od = oracle.field_digest(u, off, ln, fid, 3)
        assert dig[name] == {"n_records": off.size, "body_bytes": int(ooff[-1]), "bodies": od["bodies"]}, name
        assert od["bad"] == obad
    if expect_bad is not None:
        assert {k: g[k]["bad"] for k in FIELDS} == expect_bad
    return g


@pytest.mark.parametrize("shape,n,level", [("pe150", 20000, 6), ("se150", 5000, 1), ("longname", 4000, 9), ("long10k", 200, 6)])
def test_synthetic_shapes(oracle, engine, shape, n, level):
    g = _assert_fields(oracle, engine, synth.generate(shape, n, level=level).tobytes(), expect_bad={"cigar": 0, "seq": 0, "qual": 0})
    assert g["seq"]["total"] > 0 and g["cigar"]["total"] > 0


def _seq_bytes(codes):
    b = bytearray()
    for i in range(0, len(codes), 2):
        b.append((codes[i] << 4) | (codes[i + 1] if i + 1 < len(codes) else 0))
    return bytes(b)


def _raw_record(ref_id=0, pos=5, name=b"r", cigar=b"", n_cigar=None, seq=b"", l_seq=0, qual=b"", tags=b"", flag=0):
    nb = name + b"\0"
    n_cigar = len(cigar) // 4 if n_cigar is None else n_cigar
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(nb) & 0xFF, 30, 4681, n_cigar, flag, l_seq, -1, -1, 0) + nb + cigar + seq + qual + tags
    return struct.pack("<I", len(body)) + body


def _ops(*ops):
    return b"".join(struct.pack("<I", (n << 4) | op) for n, op in ops)


def test_cigar_seq_qual_edge_cases(oracle, engine):
    rnd = random.Random(4)
    recs = []
    # every operation, every digit count up to the 28-bit maximum
    recs.append(_raw_record(cigar=_ops(*[(10 ** k % (1 << 28), k % 9) for k in range(9)], ((1 << 28) - 1, 8), (0, 0)), seq=_seq_bytes([1] * 4), l_seq=4, qual=b"\x01\x02\x03\x04"))
    # unmapped: refID < 0 / pos < 0 -> "" although the field holds operations
    recs.append(_raw_record(ref_id=-1, cigar=_ops((5, 0))))
    recs.append(_raw_record(pos=-1, cigar=_ops((5, 0))))
    recs.append(_raw_record(cigar=b""))
    # sequences: every base code in both nibbles; the reference drops EVERY zero low nibble, odd lengths, empty
    recs.append(_raw_record(cigar=_ops((32, 0)), seq=_seq_bytes(list(range(16)) * 2), l_seq=32, qual=bytes(range(32))))
    recs.append(_raw_record(cigar=_ops((7, 0)), seq=_seq_bytes([8, 0, 0, 4, 2, 0, 1]), l_seq=7, qual=bytes([0xFF] * 7)))
    recs.append(_raw_record(cigar=_ops((1, 0)), seq=_seq_bytes([15]), l_seq=1, qual=b"\x00"))
    # long read with a many-operation CIGAR (several warp rounds)
    ops = [(rnd.randrange(1, 3000), rnd.randrange(9)) for _ in range(1000)]
    L = 1001
    recs.append(_raw_record(name=b"long", cigar=_ops(*ops), seq=_seq_bytes([rnd.randrange(16) for _ in range(L)]), l_seq=L, qual=bytes(rnd.randrange(64) for _ in range(L))))
    # long-CIGAR convention: two ops, first `<x>S`; the reference compares x with (l_seq + 1) / 2, so x = 3 matches l_seq = 5 or 6
    real = _ops((70000, 0), (3, 1), (9, 2))
    cg = b"CGBI" + struct.pack("<I", 3) + real
    recs.append(_raw_record(name=b"cg5", cigar=_ops((3, 4), (100, 3)), seq=_seq_bytes([1, 2, 4, 8, 1]), l_seq=5, qual=b"!!!!!", tags=b"NMC\x05" + cg + b"XZZhello\0"))
    recs.append(_raw_record(name=b"cg6", cigar=_ops((3, 4), (100, 3)), seq=_seq_bytes([1, 2, 4, 8, 1, 2]), l_seq=6, qual=b"!!!!!!", tags=cg))
    recs.append(_raw_record(name=b"lseq-form", cigar=_ops((5, 4), (100, 3)), seq=_seq_bytes([1, 2, 4, 8, 1]), l_seq=5, qual=b"!!!!!", tags=cg))      # x == l_seq: NOT replaced
    recs.append(_raw_record(name=b"no-tag", cigar=_ops((3, 4), (100, 3)), seq=_seq_bytes([1] * 5), l_seq=5, qual=b"!!!!!", tags=b"NMC\x05"))               # falls back to the field
    recs.append(_raw_record(name=b"cg-z", cigar=_ops((3, 4), (100, 3)), seq=_seq_bytes([1] * 5), l_seq=5, qual=b"!!!!!", tags=b"CGZ" + _ops((4, 0), (2, 7)) .replace(b"\0", b"\x10") + b"\0"))
    # names long enough for the u8 wrap of `32 + l_read_name` (l_read_name >= 224): the slices start 256 bytes early
    for k in (222, 223, 229, 254):
        recs.append(_raw_record(name=bytes(rnd.choice(b"ACGT0123") for _ in range(k)), cigar=_ops((9, 0)), seq=_seq_bytes([2] * 9), l_seq=9, qual=bytes(range(9))))
    stream = bc.bam_header() + b"".join(recs)
    for payload in (0xFF00, 97):
        _assert_fields(oracle, engine, bc.bgzf(stream, payload=payload), expect_bad={"cigar": 0, "seq": 0, "qual": 0})


def test_reference_panics_become_counted_errors(oracle, engine):
    recs = [
        _raw_record(name=b"ok", cigar=_ops((4, 0)), seq=_seq_bytes([1] * 4), l_seq=4, qual=b"abcd"),
        _raw_record(name=b"op9", cigar=_ops((4, 9)), seq=_seq_bytes([1] * 4), l_seq=4, qual=b"abcd"),                    # get_cig_op panics
        _raw_record(name=b"short", cigar=_ops((4, 0)), n_cigar=3, seq=_seq_bytes([1] * 4), l_seq=4000, qual=b"abcd"),    # slices beyond the record
        _raw_record(name=b"cgodd", cigar=_ops((3, 4), (1, 3)), seq=_seq_bytes([1] * 5), l_seq=5, qual=b"!!!!!", tags=b"CGBC" + struct.pack("<I", 6) + b"\x10\0\0\0\x20\0"),   # 6 bytes: short chunk
        _raw_record(name=b"badtag", cigar=_ops((3, 4), (1, 3)), seq=_seq_bytes([1] * 5), l_seq=5, qual=b"!!!!!", tags=b"XX?\x01"),                                          # unknown tag type
        _raw_record(name=b"ok2", cigar=_ops((2, 0), (2, 4)), seq=_seq_bytes([4] * 4), l_seq=4, qual=b"wxyz"),
    ]
    stream = bc.bam_header() + b"".join(recs) + struct.pack("<I", 20) + b"\x01" * 20   # a 20-byte body: shorter than the fixed fields
    g = _assert_fields(oracle, engine, bc.bgzf(stream), expect_bad={"cigar": 5, "seq": 2, "qual": 2})
    off = g["cigar"]["off"]
    assert g["cigar"]["bytes"][int(off[0]):int(off[1])].tobytes() == b"4M" and g["cigar"]["bytes"][int(off[5]):int(off[6])].tobytes() == b"2M2S"


def test_decode_fields_needs_a_batch_and_subsets_work(oracle):
    e = bp.Engine()
    try:
        with pytest.raises(bp.BamGpuError):
            e.decode_fields(7)
        data = synth.generate("pe150", 3000, level=6).tobytes()
        u, _ = oracle.inflate_all(data)
        _, _, start = oracle.read_header(u)
        off, ln = oracle.walk_records(u, start)
        blocks, nb, *_ = bp.scan_blocks(data)
        d = e.upload(data)
        e.decode(d, blocks, nb, start, True, 0)
        f = e.decode_fields(2, bp.COPY_DATA)
        assert "cigar" not in f and f["seq"]["bytes"].tobytes() == oracle.decode_field(u, off, ln, 2)[1].tobytes()
        with pytest.raises(bp.BamGpuError):
            e.field_digest(1)                       # not decoded in this call
        f = e.decode_fields(5, 0)                   # device only
        assert f["qual"]["bytes"] is None and f["qual"]["total"] == int(ln.size) * 150
        assert e.field_digest(4)["bodies"] == oracle.field_digest(u, off, ln, 4, 2)["bodies"]
    finally:
        e.close()
