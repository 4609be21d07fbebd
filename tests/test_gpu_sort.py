"""GPU parity for the sort path (BASELINE.json configs[3], configs[4]): bamgpu_engine_sort /
bamgpu_sort_bam against the oracle's restatement of the reference's comparators (sorting/comparators.rs:61-147) and
of its output format (sort.rs:643: record bodies, stable order).  Bit-exact: the permutation and the output bytes."""
import hashlib
import io
import random
import struct

import numpy as np
import pytest

import bam_parallel_b200 as bp
from bam_parallel_b200 import synth
from bam_parallel_b200.sorting import SortBy, sort_bam
from tests import bamcraft as bc

pytestmark = pytest.mark.gpu

# oracle mode numbers (oracle/bam_oracle.c orc_sort_perm): 0 coord+strand, 1 name, 2 name+mates
ORACLE_MODE = {SortBy.CoordinatesAndStrand: 0, SortBy.Name: 1, SortBy.NameAndMatchMates: 2}


@pytest.fixture(scope="module")
def engine():
    e = bp.Engine()
    yield e
    e.close()


def _oracle_sorted(oracle, data, sort_by):
    u, _ = oracle.inflate_all(data)
    _, _, consumed = oracle.read_header(u)
    off, ln = oracle.walk_records(u, consumed)
    cols, bf, bt = oracle.extract(u, off, ln, True)
    perm = oracle.sort_perm(u, off, cols, ORACLE_MODE[sort_by])
    ub = u.tobytes()
    out = b"".join(ub[int(off[i]):int(off[i]) + int(ln[i])] for i in perm.tolist())
    return consumed, perm, out


def _engine_sorted(engine, data, start, sort_by):
    blocks, nb, _, _, _ = bp.scan_blocks(data)
    d = engine.upload(data)
    try:
        engine.decode(d, blocks, nb, start, True, bp.WANT_HI)
        perm, bodies, srt = engine.sort(int(sort_by))
        return perm.copy(), bodies.tobytes(), int(srt.n_records)
    finally:
        engine.free(d)


@pytest.mark.parametrize("sort_by", list(SortBy))
@pytest.mark.parametrize("shape,n,payload,level", [("pe150", 30000, 0xFF00, 6), ("se150", 20000, 0xFF00, 6),
                                                   ("longname", 8000, 0xFF00, 1), ("longname", 8000, 0xFF00, 9),
                                                   ("long10k", 300, 4096, 6), ("long10k", 300, 0xFF00, 1)])
def test_sort_matches_oracle(oracle, engine, shape, n, payload, level, sort_by):
    data = synth.generate(shape, n, level=level, payload=payload).tobytes()
    start, perm, out = _oracle_sorted(oracle, data, sort_by)
    g_perm, g_out, g_n = _engine_sorted(engine, data, start, sort_by)
    assert g_n == perm.size
    assert np.array_equal(g_perm, perm)
    assert g_out == out


def _rec(name: bytes, ref_id=0, pos=0, flag=0, tags=b"", seq_len=4):
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(name) + 1, 30, 4680, 1, flag, seq_len, -1, -1, 0)
    body += name + b"\0" + struct.pack("<I", (seq_len << 4) | 0) + bytes((seq_len + 1) // 2) + bytes([30] * seq_len) + tags
    return struct.pack("<I", len(body)) + body


def _bam(recs):
    return bc.bgzf(bc.bam_header() + b"".join(recs), payload=700)


def test_name_order_prefixes_ties_and_stability(oracle, engine):
    """comparators.rs:61-65: bytewise, a proper prefix first; equal names keep the input order (stable sort, merge tie ->
    lower chunk index, sort.rs:487-510)."""
    names = [b"b", b"ab", b"a", b"abc", b"ab", b"a" * 40, b"a" * 39 + b"b", b"a" * 40, b"", b"zz", b"ab", b"a" * 9, b"a" * 8, b"a" * 17, b"a" * 16]
    recs = [_rec(nm, pos=i, flag=(i * 7) & 0x7ff) for i, nm in enumerate(names)]
    data = _bam(recs)
    for sb in (SortBy.Name, SortBy.NameAndMatchMates):
        start, perm, out = _oracle_sorted(oracle, data, sb)
        g_perm, g_out, _ = _engine_sorted(engine, data, start, sb)
        assert g_perm.tolist() == perm.tolist(), sb
        assert g_out == out
    exp = sorted(range(len(names)), key=lambda i: (names[i] + b"\0", i))
    assert _engine_sorted(engine, data, start, SortBy.Name)[0].tolist() == exp


def test_match_mates_hi_then_flag(oracle, engine):
    """comparators.rs:78-93: equal names -> HI (i32), then the full u16 flag."""
    hi = lambda v: b"HIi" + struct.pack("<i", v)
    recs = [_rec(b"q1", flag=147, tags=hi(2)), _rec(b"q1", flag=99, tags=hi(2)), _rec(b"q1", flag=163, tags=hi(1)),
            _rec(b"q1", flag=83, tags=hi(-5)), _rec(b"q0", flag=16), _rec(b"q0", flag=0), _rec(b"q2", flag=1, tags=b"NMC\x01" + hi(7))]
    data = _bam(recs)
    start, perm, out = _oracle_sorted(oracle, data, SortBy.NameAndMatchMates)
    g_perm, g_out, _ = _engine_sorted(engine, data, start, SortBy.NameAndMatchMates)
    assert g_perm.tolist() == perm.tolist() == [5, 4, 3, 2, 1, 0, 6]
    assert g_out == out


def test_reference_panics_are_errors(engine):
    hi = b"HIi" + struct.pack("<i", 1)
    data = _bam([_rec(b"q1", flag=99, tags=hi), _rec(b"q1", flag=147)])            # HI on one mate only: comparators.rs:90
    start = len(bc.bam_header())
    with pytest.raises(bp.BamGpuError) as ei:
        _engine_sorted(engine, data, start, SortBy.NameAndMatchMates)
    assert ei.value.code == -30
    _engine_sorted(engine, data, start, SortBy.Name)                                # plain name sort never looks at HI
    data = _bam([_rec(b"q1", flag=0x1000), _rec(b"q2", flag=0)])                   # comparators.rs:178 from_bits().unwrap()
    with pytest.raises(bp.BamGpuError) as ei:
        _engine_sorted(engine, data, start, SortBy.CoordinatesAndStrand)
    assert ei.value.code == -31


def test_coordinate_order_unmapped_last_and_strand(oracle, engine):
    """comparators.rs:51-54,117-147: refID -1 sorts as i32::MAX; forward strand before reverse at equal (refID, pos)."""
    recs = [_rec(b"a", -1, -1, 4), _rec(b"b", 1, 5, 16), _rec(b"c", 1, 5, 0), _rec(b"d", 0, 2**31 - 1, 0), _rec(b"e", 0, 0, 16),
            _rec(b"f", 1, 5, 16), _rec(b"g", 1, 5, 0), _rec(b"h", -1, 7, 0)]
    data = _bam(recs)
    start, perm, out = _oracle_sorted(oracle, data, SortBy.CoordinatesAndStrand)
    g_perm, g_out, _ = _engine_sorted(engine, data, start, SortBy.CoordinatesAndStrand)
    assert g_perm.tolist() == perm.tolist() == [4, 3, 2, 6, 1, 5, 0, 7]
    assert g_out == out


@pytest.mark.parametrize("sort_by", list(SortBy))
def test_sort_bam_entry_point(oracle, sort_by):
    """The sort_bam mirror (sort.rs:138-178) end to end: host bytes in, sorted bodies out through a sink."""
    data = synth.generate("pe150", 40000, level=6).tobytes()
    _, _, out = _oracle_sorted(oracle, data, sort_by)
    sink = io.BytesIO()
    st = sort_bam(2000 << 20, io.BytesIO(data), sink, None, 0, 5, None, None, sort_by, None)
    assert hashlib.md5(sink.getvalue()).hexdigest() == hashlib.md5(out).hexdigest()
    assert st["records"] == 40000 and st["ms_sort"] > 0
    sink2 = io.BytesIO()
    sort_bam(2000 << 20, data, sink2, sort_by=sort_by)                               # bytes-like source
    assert sink2.getvalue() == out


def test_sort_empty_and_single(oracle, engine):
    for recs in ([], [_rec(b"only", 3, 9, 16)]):
        data = _bam(recs)
        for sb in SortBy:
            sink = io.BytesIO()
            sort_bam(0, data, sink, sort_by=sb)
            assert sink.getvalue() == b"".join(r[4:] for r in recs)


def test_sort_full_size_properties(engine):
    """BASELINE.json configs[3] shape at a size the oracle would take minutes for: size-independent properties of the
    coordinate sort of 2M paired-end records — perm is a permutation, keys are non-decreasing, equal keys keep the
    input order, and the gathered bytes are exactly the bodies in that order (checksum of checksums)."""
    n = 2_000_000
    data = synth.generate("pe150", n, level=6).tobytes()
    blocks, nb, _, _, _ = bp.scan_blocks(data)
    with bp.Reader.new(data, 8) as r:
        hdr, _ = r.read_header()
    start = 4 + len(hdr)
    d = engine.upload(data)
    try:
        batch, _, _ = engine.decode(d, blocks, nb, start, True, bp.COPY_DATA | bp.COPY_COLUMNS)
        u = batch.data.copy()
        cols = {k: batch.columns[k].copy() for k in ("rec_off", "rec_len", "coord_key", "strand")}
        perm, bodies, srt = engine.sort(int(SortBy.CoordinatesAndStrand))
        perm, bodies = perm.copy(), bodies.copy()
    finally:
        engine.free(d)
    assert int(srt.n_records) == n and np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    k, st = cols["coord_key"][perm], cols["strand"][perm].astype(np.int64)
    assert np.all(k[1:] >= k[:-1])
    same = k[1:] == k[:-1]
    assert np.all(st[1:][same] >= st[:-1][same])
    tie = same & (st[1:] == st[:-1])
    assert np.all(perm[1:][tie] > perm[:-1][tie])                                  # stability
    lens = cols["rec_len"][perm].astype(np.int64)
    assert bodies.size == int(lens.sum())
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    for i in np.random.default_rng(1).integers(0, n, 2000).tolist():
        o, l, src = int(offs[i]), int(lens[i]), int(cols["rec_off"][perm[i]])
        assert bodies[o:o + l].tobytes() == u[src:src + l].tobytes()
    # every byte: per-record additive checksums, input side taken through the permutation
    def body_sums(buf, off, ln):
        cs = np.concatenate([[0], np.cumsum(buf, dtype=np.uint64)])
        return cs[off + ln] - cs[off]
    src_off = cols["rec_off"].astype(np.int64)[perm]
    assert np.array_equal(body_sums(u, src_off, lens), body_sums(bodies, offs, lens))


# ---- output wire formats and the BGZF writer (sort.rs:310-347, :643) -------------------------------

def _records_stream(out_bodies: bytes, lens) -> bytes:
    """(u32 block_size, body)* from the bodies-only output and the record lengths in output order."""
    parts, o = [], 0
    for l in lens:
        parts.append(struct.pack("<I", l) + out_bodies[o:o + l])
        o += l
    return b"".join(parts)


@pytest.mark.parametrize("sort_by", [SortBy.CoordinatesAndStrand, SortBy.NameAndMatchMates])
@pytest.mark.parametrize("shape,n,payload", [("pe150", 40000, 0xFF00), ("long10k", 300, 4096), ("longname", 5000, 0xFF00)])
def test_sort_output_formats_and_bgzf_writer(oracle, shape, n, payload, sort_by):
    import gzip
    data = synth.generate(shape, n, level=6, payload=payload).tobytes()
    start, perm, out = _oracle_sorted(oracle, data, sort_by)
    u, _ = oracle.inflate_all(data)
    off, ln = oracle.walk_records(u, start)
    want_records = _records_stream(out, [int(ln[i]) for i in perm.tolist()])
    # (u32 block_size, body)*: the reference's spill wire format (sort.rs:340-347)
    sink = io.BytesIO()
    sort_bam(0, data, sink, sort_by=sort_by, output="records")
    assert sink.getvalue() == want_records
    # a complete BAM: header + sorted records in stored-block BGZF + EOF marker
    sink = io.BytesIO()
    st = sort_bam(0, data, sink, sort_by=sort_by, output="bam")
    bam = sink.getvalue()
    want_stream = u.tobytes()[:start] + want_records
    assert gzip.decompress(bam) == want_stream                       # every member a valid gzip: CRC32 + ISIZE checked by zlib
    assert bam.endswith(bc.BGZF_EOF)
    u2, sizes = oracle.inflate_all(bam)                              # the reference's framing rules (util.rs:18-71) read it
    assert u2.tobytes() == want_stream and int(sizes.max()) <= 0xFF00
    with bp.Reader.new(bam, 4) as r:                                  # and so does this repository's reader (CRC verified on the GPU)
        hdr, _ = r.read_header()
        assert b"BAM\x01" + hdr == u.tobytes()[:start]
        got = b"".join(bytes(x) for x in r.records())
    assert got == out
    assert st["ms_bgzf_write"] > 0


@pytest.mark.parametrize("sort_by", list(SortBy))
@pytest.mark.parametrize("shape,n,payload,limit", [("pe150", 60000, 0xFF00, 8 << 20), ("long10k", 400, 4096, 4 << 20), ("longname", 6000, 0xFF00, 1 << 20)])
def test_streamed_sort_matches_the_resident_one(oracle, shape, n, payload, limit, sort_by):
    """bamgpu_opts.sort_mem_limit (sort.rs:138 mem_limit): the file goes through the Reader in many small batches, keys and
    names stay on the device, bodies on the host — the output bytes must be those of the one-batch sort and of the oracle,
    in all three output formats."""
    import gzip
    data = synth.generate(shape, n, level=6, payload=payload).tobytes()
    start, perm, out = _oracle_sorted(oracle, data, sort_by)
    u, _ = oracle.inflate_all(data)
    off, ln = oracle.walk_records(u, start)
    sink = io.BytesIO()
    st = sort_bam(0, io.BytesIO(data), sink, sort_by=sort_by, device_mem_limit=limit)
    assert st["batches"] > 3, st["batches"]                          # really streamed
    assert hashlib.md5(sink.getvalue()).hexdigest() == hashlib.md5(out).hexdigest()
    assert st["records"] == n and st["ms_sort"] > 0
    want_records = _records_stream(out, [int(ln[i]) for i in perm.tolist()])
    sink = io.BytesIO()
    sort_bam(0, data, sink, sort_by=sort_by, output="records", device_mem_limit=limit)
    assert sink.getvalue() == want_records
    sink = io.BytesIO()
    sort_bam(0, data, sink, sort_by=sort_by, output="bam", device_mem_limit=limit)
    bam = sink.getvalue()
    assert gzip.decompress(bam) == u.tobytes()[:start] + want_records and bam.endswith(bc.BGZF_EOF)
    resident = io.BytesIO()
    sort_bam(0, data, resident, sort_by=sort_by, output="bam")
    assert resident.getvalue() == bam                                 # byte-identical files either way


def test_streamed_sort_edge_cases(oracle):
    for recs in ([], [_rec(b"only", 3, 9, 16)]):
        data = _bam(recs)
        for sb in SortBy:
            sink = io.BytesIO()
            sort_bam(0, data, sink, sort_by=sb, device_mem_limit=1 << 20)
            assert sink.getvalue() == b"".join(r[4:] for r in recs)
            sink = io.BytesIO()
            sort_bam(0, data, sink, sort_by=sb, device_mem_limit=1 << 20, output="bam")
            u, _ = oracle.inflate_all(sink.getvalue())
            _, _, start = oracle.read_header(u)
            assert u.tobytes()[start:] == b"".join(recs)
    # the reference's panics are errors here too
    data = _bam([_rec(b"q", 0, 1, 0, tags=b"HIi" + struct.pack("<i", 1)), _rec(b"q", 0, 2, 0)])
    with pytest.raises(bp.BamGpuError) as e:
        sort_bam(0, data, io.BytesIO(), sort_by=SortBy.NameAndMatchMates, device_mem_limit=1 << 20)
    assert e.value.code == -30
    data = _bam([_rec(b"q", 0, 1, 0x1000)])
    with pytest.raises(bp.BamGpuError) as e:
        sort_bam(0, data, io.BytesIO(), sort_by=SortBy.CoordinatesAndStrand, device_mem_limit=1 << 20)
    assert e.value.code == -31


def test_bgzf_store_edge_sizes(oracle, engine):
    import gzip
    rnd = random.Random(5)
    for n in (0, 1, 0xFF00 - 1, 0xFF00, 0xFF00 + 1, 3 * 0xFF00, 3 * 0xFF00 + 77):
        payload = bytes(rnd.randrange(256) for _ in range(n))
        for eof in (True, False):
            z = engine.bgzf_store(payload, eof_marker=eof)
            assert gzip.decompress(z) == payload if z else n == 0
            assert z.endswith(bc.BGZF_EOF) == eof or (not eof and n == 0)
            blocks, nb, consumed, total, rc = bp.scan_blocks(z)
            assert consumed == len(z) and total == n
