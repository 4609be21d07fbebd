"""CPU-only: pins the oracle.  (1) the one reference-held known-answer vector on this
path, BAMRawRecord::default() (bam_tools/src/record/bamrawrecord.rs:180-197); (2) the C
restatement against the independent pure-Python twin; (3) every reference quirk listed in
SURVEY.md §8(c)."""
import gzip
import hashlib
import struct
import zlib

import numpy as np
import pytest

from bam_parallel_b200 import synth
from oracle import oracle as orc
from tests import bamcraft as bc

# bamrawrecord.rs:180-197, bytes copied as DATA (a golden vector), comments give the expected fields.
DEFAULT_RECORD = bytes([
    0xff, 0xff, 0xff, 0xff,  # ref_id = -1
    0xff, 0xff, 0xff, 0xff,  # pos = -1
    0x02,                    # l_read_name = 2
    0xff,                    # mapq = 255
    0x48, 0x12,              # bin = 4680
    0x00, 0x00,              # n_cigar_op = 0
    0x04, 0x00,              # flag = 4
    0x00, 0x00, 0x00, 0x00,  # l_seq = 0
    0xff, 0xff, 0xff, 0xff,  # next_ref_id = -1
    0xff, 0xff, 0xff, 0xff,  # next_pos = -1
    0x00, 0x00, 0x00, 0x00,  # tlen = 0
    0x2a, 0x00,              # read_name = "*\x00"
])


def _decode(oracle, data, want_hi=True):
    u, sizes = oracle.inflate_all(data)
    hdr, off_refs, consumed = oracle.read_header(u)
    off, ln = oracle.walk_records(u, consumed)
    cols, bad_flags, bad_tags = oracle.extract(u, off, ln, want_hi)
    return u, sizes, hdr, off_refs, consumed, off, ln, cols, bad_flags, bad_tags


def test_reference_default_record_vector(oracle):
    stream = bc.bam_header() + struct.pack("<I", len(DEFAULT_RECORD)) + DEFAULT_RECORD
    u, sizes, hdr, off_refs, consumed, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(stream))
    assert ln.tolist() == [34]
    exp = dict(ref_id=-1, pos=-1, l_read_name=2, mapq=255, bin=4680, n_cigar_op=0, flag=4, l_seq=0, next_ref_id=-1,
               next_pos=-1, tlen=0, cigar_off=34, seq_off=34, qual_off=34, tags_off=34)
    for k, v in exp.items():
        assert int(cols[k][0]) == v, k
    # comparators.rs:51-54: refID -1 sorts as i32::MAX
    assert int(cols["coord_key"][0]) == ((0x7FFFFFFF ^ 0x80000000) << 32) | (0xFFFFFFFF ^ 0x80000000)
    assert bytes(u[off[0] + 32: off[0] + 34]) == b"*\0"
    rows = orc.py_extract(bytes(u), off.tolist(), ln.tolist())
    for k, v in exp.items():
        assert rows[0][k] == v


@pytest.mark.parametrize("shape,n,payload", [("se150", 3000, 0xFF00), ("pe150", 3000, 0xFF00), ("longname", 2000, 0xFF00),
                                             ("long10k", 60, 4096), ("long10k", 40, 0xFF00), ("pe150", 500, 300)])
@pytest.mark.parametrize("level", [1, 6, 9])
def test_c_oracle_matches_python_twin(oracle, shape, n, payload, level):
    bam = synth.generate(shape, n, level=level, payload=payload, threads=2)
    data = bam.tobytes()
    u, sizes, hdr, off_refs, consumed, off, ln, cols, bf, bt = _decode(oracle, data)
    # independent decoders agree on the byte stream (python gzip handles multi-member files)
    assert bytes(u) == gzip.decompress(data)
    pu, psizes = orc.py_inflate_all(data)
    assert pu == bytes(u) and psizes == sizes.tolist()
    assert len(u) == bam.ustream_len and consumed == bam.header_len
    phdr, poff_refs, pconsumed = orc.py_read_header(pu)
    assert (phdr, poff_refs, pconsumed) == (hdr, off_refs, consumed)
    refs = orc.py_parse_reference_sequences(hdr[off_refs:])
    assert len(refs) == 24 and refs[0] == ("chr1", 249250621)
    poff, pln = orc.py_walk_records(pu, pconsumed)
    assert poff == off.tolist() and pln == ln.tolist() and len(poff) == n
    rows = orc.py_extract(pu, poff, pln)
    for k in orc.COLUMN_DTYPES:
        assert [r[k] for r in rows] == cols[k].tolist(), k
    assert [r["strand"] for r in rows] == cols["strand"].tolist()
    assert bf == 0 and bt == 0
    for mode in (0, 1, 2):
        assert orc.py_sort_perm(rows, mode) == oracle.sort_perm(u, off, cols, mode).tolist(), mode


def test_generator_is_thread_count_invariant():
    a = synth.generate("pe150", 20000, threads=1).tobytes()
    b = synth.generate("pe150", 20000, threads=5).tobytes()
    assert hashlib.md5(a).hexdigest() == hashlib.md5(b).hexdigest()


def test_longname_has_hi_and_mates(oracle):
    bam = synth.generate("longname", 400, threads=1)
    *_, off, ln, cols, bf, bt = _decode(oracle, bam.tobytes())
    assert cols["hi_present"].all() and set(cols["hi_value"].tolist()) == {1, 2}
    assert cols["l_read_name"].min() >= 200 and cols["l_read_name"].max() <= 222


# ---- quirks --------------------------------------------------------------

def _one_rec_stream(**kw):
    return bc.bam_header() + bc.bam_record(**kw)


def test_quirk_crc_and_isize_never_read(oracle):
    stream = _one_rec_stream()
    data = bc.bgzf_block(stream, crc=0xDEADBEEF, isize=7) + bc.BGZF_EOF
    u, _ = oracle.inflate_all(data)
    assert bytes(u) == stream                       # util.rs:44,68-71
    assert orc.py_inflate_all(data)[0] == stream


def test_quirk_bsize_ffff_is_silent_eof(oracle):
    stream = _one_rec_stream()
    good = bc.bgzf_block(stream)
    evil = bytearray(bc.bgzf_block(b"x" * 10))
    evil[16:18] = b"\xff\xff"                        # BSIZE+1 wraps to 0 (util.rs:65)
    data = good + bytes(evil) + bc.bgzf_block(b"never reached")
    u, sizes = oracle.inflate_all(data)
    assert bytes(u) == stream and len(sizes) == 1
    assert orc.py_inflate_all(data)[0] == stream


def test_quirk_short_trailing_header_is_silent_eof(oracle):
    stream = _one_rec_stream()
    data = bc.bgzf_block(stream) + b"\x1f\x8b\x08\x04\x00"   # 5 stray bytes (util.rs:58-60)
    assert bytes(oracle.inflate_all(data)[0]) == stream
    assert orc.py_inflate_all(data)[0] == stream


def test_quirk_no_magic_check_on_bgzf_header(oracle):
    stream = _one_rec_stream()
    blk = bytearray(bc.bgzf_block(stream))
    blk[0:4] = b"ABCD"                               # util.rs:55-66 looks only at bytes 16..18
    assert bytes(oracle.inflate_all(bytes(blk))[0]) == stream


def test_quirk_eof_marker_midstream_is_an_empty_block(oracle):
    s = bc.bam_header() + bc.bam_record(name=b"a") + bc.bam_record(name=b"b")
    cut = len(bc.bam_header()) + 20
    data = bc.bgzf_block(s[:cut]) + bc.BGZF_EOF + bc.bgzf_block(s[cut:]) + bc.BGZF_EOF
    u, sizes = oracle.inflate_all(data)
    assert bytes(u) == s and sizes.tolist() == [cut, 0, len(s) - cut, 0]
    _, _, consumed = oracle.read_header(u)
    assert oracle.walk_records(u, consumed)[0].size == 2


def test_error_block_too_small_and_truncated(oracle):
    tiny = bytearray(bc.BGZF_EOF)
    tiny[16:18] = struct.pack("<H", 20 - 1)
    with pytest.raises(orc.OracleError) as e:
        oracle.inflate_all(bytes(tiny))
    assert e.value.code == -1                        # util.rs:29-38
    blk = bc.bgzf_block(_one_rec_stream())
    with pytest.raises(orc.OracleError) as e:
        oracle.inflate_all(blk[:-3])
    assert e.value.code == -2
    with pytest.raises(orc.OracleError):
        orc.py_inflate_all(blk[:-3])


def test_error_corrupt_deflate(oracle):
    blk = bytearray(bc.bgzf_block(bytes(range(256)) * 20, level=6))
    blk[18] |= 0x06                                  # BTYPE=3 (reserved)
    with pytest.raises(orc.OracleError) as e:
        oracle.inflate_all(bytes(blk))
    assert e.value.code == -3


def test_quirk_block_size_zero_ends_iteration(oracle):
    s = bc.bam_header() + bc.bam_record(name=b"a") + struct.pack("<I", 0) + bc.bam_record(name=b"ghost")
    u, _ = oracle.inflate_all(bc.bgzf(s))
    _, _, consumed = oracle.read_header(u)
    off, ln = oracle.walk_records(u, consumed)       # reader.rs:64-72, records.rs:25
    assert off.size == 1
    assert orc.py_walk_records(bytes(u), consumed)[0] == off.tolist()


@pytest.mark.parametrize("extra", [0, 1, 2, 3])
def test_quirk_trailing_bytes_lt4_are_silent_eof(oracle, extra):
    s = bc.bam_header() + bc.bam_record() + b"\x07" * extra
    u, _ = oracle.inflate_all(bc.bgzf(s))
    _, _, consumed = oracle.read_header(u)
    assert oracle.walk_records(u, consumed)[0].size == 1   # reader.rs:78


def test_error_short_record_body(oracle):
    s = bc.bam_header() + bc.bam_record()[:-5]
    u, _ = oracle.inflate_all(bc.bgzf(s))
    _, _, consumed = oracle.read_header(u)
    with pytest.raises(orc.OracleError) as e:
        oracle.walk_records(u, consumed)
    assert e.value.code == -7                        # reader.rs:69-70 expect("Failed to read.")


def test_error_bad_magic_and_short_header(oracle):
    u = np.frombuffer(b"BAX\x01" + b"\0" * 20, dtype=np.uint8)
    with pytest.raises(orc.OracleError) as e:
        oracle.read_header(u)
    assert e.value.code == -5
    h = bc.bam_header()
    with pytest.raises(orc.OracleError) as e:
        oracle.read_header(np.frombuffer(h[:-2], dtype=np.uint8))
    assert e.value.code == -6


def test_quirk_name_key_includes_nul_and_order(oracle):
    names = [b"b", b"a\x01", b"a", b"ab", b"a"]
    s = bc.bam_header() + b"".join(bc.bam_record(name=n, pos=i) for i, n in enumerate(names))
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(s))
    perm = oracle.sort_perm(u, off, cols, 1).tolist()
    assert [names[i] for i in perm] == [b"a", b"a", b"a\x01", b"ab", b"b"]
    assert perm[:2] == [2, 4]                        # stable


def test_quirk_tag_offset_wraps_for_long_names(oracle):
    # bamrawrecord.rs:81: `32 + l_read_name` in u8. l_read_name = 230 -> tags slice starts 256 bytes early.
    name = b"N" * 229
    tags = b"HIi" + struct.pack("<i", 5)
    rec = bc.bam_record(name=name, tags=tags, seq_len=10)
    s = bc.bam_header() + rec
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(s))
    assert int(cols["l_read_name"][0]) == 230
    assert int(cols["tags_off"][0]) == 32 + 230 + 4 + 5 + 10   # the usize form (bamrawrecord.rs:61-77) is right
    # ... but the HI lookup walks from the wrapped offset: lands inside the name ('NN' + type 'N') => panic
    assert bt == 1 and int(cols["hi_present"][0]) == 0
    with pytest.raises(orc.OracleError):
        orc.py_hit_count(rec[4:])
    # l_read_name = 223 is the largest safe value
    rec2 = bc.bam_record(name=b"N" * 222, tags=tags, seq_len=10)
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(bc.bam_header() + rec2))
    assert bt == 0 and int(cols["hi_present"][0]) == 1 and int(cols["hi_value"][0]) == 5


@pytest.mark.parametrize("t,raw,val", [("A", b"\x07", 7), ("c", b"\xfe", -2), ("C", b"\xfe", 254), ("s", b"\xfe\xff", -2),
                                       ("S", b"\xfe\xff", 65534), ("i", b"\xfe\xff\xff\xff", -2),
                                       ("I", b"\xfe\xff\xff\xff", -2)])
def test_hi_tag_types(oracle, t, raw, val):
    tags = b"NMC\x01" + b"XZZab\0" + b"XBBs\x02\0\0\0\x01\0\x02\0" + b"HI" + t.encode() + raw
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(bc.bam_header() + bc.bam_record(tags=tags)))
    assert bt == 0 and int(cols["hi_present"][0]) == 1 and int(cols["hi_value"][0]) == val   # tags.rs:113-129
    assert orc.py_hit_count(bc.bam_record(tags=tags)[4:]) == val


def test_hi_bad_types_are_errors(oracle):
    for tags in (b"HIf\0\0\0\0", b"HIZabc\0", b"XX?\0"):
        u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(bc.bam_header() + bc.bam_record(tags=tags)))
        assert bt == 1


def test_quirk_flag_bits_ge_0x1000_counted(oracle):
    s = bc.bam_header() + bc.bam_record(flag=0x1010) + bc.bam_record(flag=0x10)
    *_r, cols, bf, bt = _decode(oracle, bc.bgzf(s))
    assert bf == 1                                   # comparators.rs:178 from_bits(..).unwrap()


def test_name_mates_order_and_mixed_hi_error(oracle):
    hi = lambda v: b"HIi" + struct.pack("<i", v)
    recs = [bc.bam_record(name=b"q", flag=0x80, tags=hi(1)), bc.bam_record(name=b"q", flag=0x40, tags=hi(2)),
            bc.bam_record(name=b"q", flag=0x40, tags=hi(1)), bc.bam_record(name=b"p", flag=0x40, tags=hi(9))]
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(bc.bam_header() + b"".join(recs)))
    assert oracle.sort_perm(u, off, cols, 2).tolist() == [3, 2, 0, 1]   # name, then HI, then flag
    recs2 = [bc.bam_record(name=b"q", tags=hi(1)), bc.bam_record(name=b"q")]
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(bc.bam_header() + b"".join(recs2)))
    with pytest.raises(orc.OracleError) as e:
        oracle.sort_perm(u, off, cols, 2)
    assert e.value.code == -9                        # comparators.rs:90


def test_coord_order_unmapped_last_and_strand(oracle):
    recs = [bc.bam_record(ref_id=-1, pos=-1, flag=4), bc.bam_record(ref_id=1, pos=5, flag=16),
            bc.bam_record(ref_id=1, pos=5, flag=0), bc.bam_record(ref_id=0, pos=900, flag=0),
            bc.bam_record(ref_id=1, pos=-1, flag=0)]
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bc.bgzf(bc.bam_header() + b"".join(recs)))
    assert oracle.sort_perm(u, off, cols, 0).tolist() == [3, 4, 2, 1, 0]   # comparators.rs:51-57,117-147


def test_cpu_reader_topology_counts_and_early_stop(oracle):
    bam = synth.generate("pe150", 30000, threads=2)
    data = bam.array
    for th in (1, 3, 5):
        r = oracle.cpu_reader_run(data, th)
        assert r["records"] == 30000 and r["ubytes"] == bam.ustream_len
    r = oracle.cpu_reader_run(data, 4, max_records=1234)
    assert r["records"] == 1234


def test_md5_of_bodies_is_stable(oracle):
    """main.rs:83-99 `-h`: md5 over every record body.  Self-golden for the seeded generator:
    guards generator + oracle drift between rounds (not a reference-published value)."""
    bam = synth.generate("pe150", 5000, threads=1)
    u, *_r, off, ln, cols, bf, bt = _decode(oracle, bam.tobytes())
    h = hashlib.md5()
    ub = bytes(u)
    for o, l in zip(off.tolist(), ln.tolist()):
        h.update(ub[o:o + l])
    import json, os
    gpath = os.path.join(os.path.dirname(__file__), "golden", "selfgolden.json")
    g = json.load(open(gpath))
    assert h.hexdigest() == g["pe150_5000_L6_md5_bodies"]
    assert orc.zlib.crc32(ub) == oracle.crc32(u)


@pytest.mark.parametrize("shape,n,payload,level", [("pe150", 4000, 0xFF00, 6), ("longname", 2000, 0xFF00, 1), ("long10k", 40, 4096, 9)])
def test_inflate_against_pythons_gzip_reader(oracle, shape, n, payload, level):
    """An implementation that shares nothing with the oracle's framing code: CPython's gzip module reads a BGZF file as
    a multi-member gzip stream (and, unlike the reference, checks every CRC32 and ISIZE).  Same bytes."""
    import gzip
    data = synth.generate(shape, n, level=level, payload=payload, threads=2).tobytes()
    u, sizes = oracle.inflate_all(data)
    assert gzip.decompress(data) == u.tobytes()
    assert int(sizes.sum()) == u.size


def test_sorted_output_format_matches_sort_rs(oracle):
    """sort.rs:643 writes rec bodies only (no block_size); a stable sort by the coordinate comparator of a file that is
    already sorted must reproduce the hash-mode byte stream (main.rs:83-99) exactly."""
    recs = [bc.bam_record(ref_id=r, pos=p, flag=f, name=b"n%d" % i) for i, (r, p, f) in
            enumerate([(0, 1, 0), (0, 1, 16), (0, 5, 0), (1, 0, 0), (1, 0, 0), (-1, -1, 4)])]
    data = bc.bgzf(bc.bam_header() + b"".join(recs), payload=100)
    u, _ = oracle.inflate_all(data)
    _, _, start = oracle.read_header(u)
    off, ln = oracle.walk_records(u, start)
    cols, _, _ = oracle.extract(u, off, ln, True)
    assert oracle.sort_perm(u, off, cols, 0).tolist() == list(range(len(recs)))
    ub = u.tobytes()
    assert b"".join(ub[int(o):int(o) + int(l)] for o, l in zip(off, ln)) == b"".join(r[4:] for r in recs)


# ---- variable-length field decode: C restatement vs the independent pure-Python one --------------------

def test_field_decoders_c_vs_python(oracle):
    from oracle.oracle import OracleError, py_decode_field
    for shape, n in (("pe150", 1500), ("longname", 600), ("long10k", 30)):
        data = synth.generate(shape, n, level=6, threads=2).tobytes()
        u, _ = oracle.inflate_all(data)
        _, _, start = oracle.read_header(u)
        off, ln = oracle.walk_records(u, start)
        ub = u.tobytes()
        for field in (1, 2, 4):
            fo, fb, bad = oracle.decode_field(u, off, ln, field)
            assert bad == 0
            for i in range(off.size):
                assert fb[int(fo[i]):int(fo[i + 1])].tobytes() == py_decode_field(ub[int(off[i]):int(off[i]) + int(ln[i])], field)
            d = oracle.field_digest(u, off, ln, field, 3)
            want = sum(orc.py_mix64(orc.py_body_hash(fb[int(fo[i]):int(fo[i + 1])].tobytes()) + i * 0x9E3779B97F4A7C15) for i in range(off.size)) & (2 ** 64 - 1)
            assert d["bodies"] == want and d["body_bytes"] == int(fo[-1]) and d["bad"] == 0


def test_field_decoder_known_answers(oracle):
    """Values worked out by hand from bamrawrecord.rs: decode_cigar :223-236, decode_seq :259-270 (zero low nibbles dropped),
    get_cigar :126-157 (CG substitution keyed on (l_seq + 1) / 2), the u8 wrap of `32 + l_read_name` (:80)."""
    from oracle.oracle import OracleError, py_decode_field
    import struct as st

    def rec(ref=0, pos=1, name=b"r", cigar=b"", seq=b"", l_seq=0, qual=b"", tags=b""):
        nb = name + b"\0"
        return st.pack("<iiBBHHHIiii", ref, pos, len(nb) & 0xFF, 0, 0, len(cigar) // 4, 0, l_seq, -1, -1, 0) + nb + cigar + seq + qual + tags

    ops = lambda *o: b"".join(st.pack("<I", (n << 4) | op) for n, op in o)
    r = rec(cigar=ops((10, 0), (2, 1), (268435455, 8)), seq=bytes([0x12, 0x40, 0x8F]), l_seq=6, qual=b"abcdef")
    assert py_decode_field(r, 1) == b"10M2I268435455X"
    assert py_decode_field(r, 2) == b"ACGTN"            # 0x40: the low nibble 0 is dropped although it is not padding
    assert py_decode_field(r, 4) == b"abcdef"
    assert py_decode_field(rec(ref=-1, cigar=ops((5, 0))), 1) == b""
    cg = b"CGBI" + st.pack("<I", 2) + ops((7, 0), (9, 2))
    assert py_decode_field(rec(cigar=ops((3, 4), (1, 3)), seq=b"\x11\x11\x10", l_seq=5, qual=b"!" * 5, tags=cg), 1) == b"7M9D"
    assert py_decode_field(rec(cigar=ops((5, 4), (1, 3)), seq=b"\x11\x11\x10", l_seq=5, qual=b"!" * 5, tags=cg), 1) == b"5S1N"
    with pytest.raises(OracleError):
        py_decode_field(rec(cigar=ops((1, 9))), 1)
    # the C restatement gives the same answers
    body = rec(cigar=ops((3, 4), (1, 3)), seq=b"\x11\x11\x10", l_seq=5, qual=b"!" * 5, tags=cg)
    u = np.frombuffer(st.pack("<I", len(body)) + body + bytes(64), dtype=np.uint8)
    fo, fb, bad = oracle.decode_field(u, np.array([4], np.uint64), np.array([len(body)], np.uint32), 1)
    assert fb.tobytes() == b"7M9D" and bad == 0


def test_reference_flags_known_answers(oracle):
    """The reference's own unit tests for Flags (sorting/flags.rs:207-251: READ_1 == 0x40, twelve flags, default empty) pin
    what the coordinate comparator reads from the flag word (comparators.rs:173-180): strand = REVERSE_COMPLEMENTED = 0x10,
    and Flags::from_bits(..).unwrap() panics exactly for bits >= 0x1000 (flags.rs:2-31 defines bits 0..11)."""
    flags = [0x0, 0x40, 0x10, 0x10 | 0x40, 0x1, 0x2, 0x4, 0x8, 0x20, 0x80, 0x100, 0x200, 0x400, 0x800, 0xFFF, 0x1000, 0x8010]
    recs = [bc.bam_record(name=b"f%d" % i, flag=f) for i, f in enumerate(flags)]
    u, _ = oracle.inflate_all(bc.bgzf(bc.bam_header() + b"".join(recs)))
    _, _, start = oracle.read_header(u)
    off, ln = oracle.walk_records(u, start)
    cols, bad_flags, _ = oracle.extract(u, off, ln, True)
    assert cols["flag"].tolist() == flags
    assert cols["strand"].tolist() == [1 if f & 0x10 else 0 for f in flags]        # is_reverse_complemented
    assert bad_flags == 2                                                          # 0x1000 and 0x8010 only
