"""CPU-only: the K1 per-lane DEFLATE logic (inflate_core.cuh), compiled for the host, against the oracle.
The same header is what the CUDA kernel instantiates; the -m gpu tests then check the kernel itself."""
import random
import struct
import zlib

import numpy as np
import pytest

from bam_parallel_b200 import synth
from tests import bamcraft as bc
from tests import emul


def _check(oracle, data, lanes=32):
    out, status, (lead, tail) = emul.inflate(data, lanes)
    u, sizes = oracle.inflate_all(data)
    assert status.tolist() == [0] * len(status)
    assert out.tobytes() == u.tobytes()
    assert (lead == 0xEE).all() and (tail == 0xEE).all()      # nothing written outside the stream
    return out


@pytest.mark.parametrize("shape,n,payload", [("se150", 4000, 0xFF00), ("pe150", 4000, 0xFF00), ("longname", 2000, 0xFF00),
                                             ("long10k", 60, 4096), ("pe150", 300, 333), ("pe150", 2000, 7001)])
@pytest.mark.parametrize("level", [1, 6, 9])
def test_synth_shapes(oracle, shape, n, payload, level):
    _check(oracle, synth.generate(shape, n, level=level, payload=payload, threads=2).tobytes())


@pytest.mark.parametrize("lanes", [1, 3, 32])
def test_lane_counts(oracle, lanes):
    _check(oracle, synth.generate("pe150", 3000, threads=2).tobytes(), lanes)


def _blocks(payloads, **kw):
    return b"".join(bc.bgzf_block(p, **kw) for p in payloads) + bc.BGZF_EOF


def test_stored_fixed_and_dynamic_blocks(oracle):
    rnd = random.Random(1)
    noise = bytes(rnd.randrange(256) for _ in range(40000))
    text = b"".join(b"read%05d\tACGTACGTTTGA\t!!!!FFFFFFFF\n" % i for i in range(1500))
    data = (_blocks([noise], level=0) + _blocks([noise], level=6)        # stored blocks (level 0 / incompressible)
            + _blocks([b"abcabcabcabc" * 3, b"a", b"", b"xyz"], level=6)  # tiny -> fixed Huffman
            + _blocks([text[:60000]], level=9)
            + _blocks([text[:50000]], level=6, strategy=zlib.Z_FIXED)     # long fixed-Huffman block with matches
            + _blocks([text[:50000]], level=6, strategy=zlib.Z_HUFFMAN_ONLY)
            + _blocks([b"\0" * 65000, b"ab" * 30000, b"abc" * 20000], level=6, strategy=zlib.Z_RLE))
    _check(oracle, data)


@pytest.mark.parametrize("period", [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 258, 259, 1000, 32768])
def test_overlapping_matches_every_small_distance(oracle, period):
    rnd = random.Random(period)
    pat = bytes(rnd.randrange(1, 256) for _ in range(period))
    for lead in range(0, 5):                     # every output alignment
        body = bytes(rnd.randrange(256) for _ in range(lead)) + (pat * (65000 // period + 1))[: 65000 - lead]
        _check(oracle, _blocks([body], level=6))


def test_many_deflate_blocks_per_bgzf_block(oracle):
    # Z_FULL_FLUSH every few hundred bytes -> dozens of DEFLATE blocks (incl. empty stored) in one BGZF block
    rnd = random.Random(7)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts, raw = [], b""
    for i in range(60):
        chunk = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randrange(1, 900)))
        parts.append(chunk)
        raw += c.compress(chunk) + c.flush(zlib.Z_FULL_FLUSH if i % 3 else zlib.Z_SYNC_FLUSH)
    raw += c.flush()
    payload = b"".join(parts)
    _check(oracle, bc.bgzf_block(payload, raw_deflate=raw) + bc.BGZF_EOF)


def test_unaligned_everything(oracle):
    rnd = random.Random(3)
    payloads = [bytes(rnd.choice(b"ACGTN!F#") for _ in range(rnd.randrange(0, 700))) for _ in range(200)]
    _check(oracle, _blocks(payloads, level=6))
    _check(oracle, _blocks(payloads, level=1), lanes=5)


def test_max_size_block(oracle):
    rnd = random.Random(5)
    _check(oracle, _blocks([bytes(rnd.randrange(256) for _ in range(0xFF00))], level=6))
    _check(oracle, _blocks([b"Q" * 65535], level=9))      # ISIZE 65535 in a tiny payload


def test_empty_file_and_eof_only(oracle):
    out, status, _ = emul.inflate(b"")
    assert out.size == 0 and status.size == 0
    _check(oracle, bc.BGZF_EOF)
    _check(oracle, bc.BGZF_EOF * 3)


def _status_of(data, table=None):
    out, status, (lead, tail) = emul.inflate(data, 4, table)
    assert (lead == 0xEE).all() and (tail == 0xEE).all()
    return status.tolist()


def test_corrupt_streams_report_status_and_stay_in_bounds():
    rnd = random.Random(11)
    good = bytes(rnd.choice(b"ACGT") for _ in range(30000))
    blk = bytearray(bc.bgzf_block(good, level=6))
    bad = bytearray(blk); bad[18] |= 0x06                       # BTYPE 3
    assert _status_of(bytes(bad) + bc.BGZF_EOF)[0] == 1
    for trial in range(60):                                      # random bit flips: any status, never OOB / hang
        b2 = bytearray(blk)
        for _ in range(rnd.randrange(1, 4)):
            i = rnd.randrange(18, len(b2) - 8)
            b2[i] ^= 1 << rnd.randrange(8)
        st = _status_of(bytes(blk) + bytes(b2) + bytes(blk) + bc.BGZF_EOF)
        assert st[0] == 0 and st[2] == 0 and st[3] == 0        # neighbours unaffected
    # ISIZE smaller / larger than the real size
    assert _status_of(bc.bgzf_block(good, isize=len(good) - 1) + bc.BGZF_EOF)[0] == 9
    assert _status_of(bc.bgzf_block(good, isize=len(good) + 1) + bc.BGZF_EOF)[0] == 11
    # truncated DEFLATE payload (stream runs past its bytes)
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    rawb = raw.compress(good) + raw.flush()
    st = _status_of(bc.bgzf_block(good, raw_deflate=rawb[: len(rawb) // 2]) + bc.BGZF_EOF)
    assert st[0] != 0
    # stored block with bad NLEN
    st = _status_of(bc.bgzf_block(b"abcd", raw_deflate=b"\x01\x04\x00\x00\x00abcd") + bc.BGZF_EOF)
    assert st[0] == 2
    # distance before start of block
    st = _status_of(bc.bgzf_block(b"aaaa", raw_deflate=bytes([0x4b, 0x04, 0x03, 0x00])[:0] + zlib.compress(b"a" * 4)[2:-4]) + bc.BGZF_EOF)
    assert st[0] == 0


def test_distance_too_far_is_rejected():
    # fixed block: literal 'a' then match len 3 dist 2 (reaches before the block start)
    bits, nb = 0, 0
    def put(v, n):
        nonlocal bits, nb
        bits |= v << nb; nb += n
    def put_code(code, n):            # Huffman codes are sent MSB first
        for i in range(n - 1, -1, -1):
            put((code >> i) & 1, 1)
    put(1, 1); put(1, 2)              # BFINAL, BTYPE=01
    put_code(0x30 + ord("a"), 8)      # literal 'a'
    put_code(1, 7)                    # length symbol 257 (len 3)
    put_code(1, 5)                    # distance symbol 1 (dist 2)
    put_code(0, 7)                    # EOB
    raw = bits.to_bytes((nb + 7) // 8, "little")
    assert _status_of(bc.bgzf_block(b"aaaa", raw_deflate=raw) + bc.BGZF_EOF)[0] == 8


@pytest.mark.parametrize("defines", [("BG_META_REGS=1",), ("BG_LITS_PER_TOKEN=2",), ("BG_RING_SHIFT=1",), ("BG_RING_SHIFT=5", "BG_LITS_PER_TOKEN=3")])
def test_kernel_variants_behind_macros(oracle, defines):
    """The switches DESIGN.md's experiments used (META in registers, literals per token, ring depth) stay correct."""
    L = emul.variant(defines)
    for shape, n, payload, level in [("pe150", 3000, 0xFF00, 6), ("long10k", 40, 4096, 1), ("longname", 1500, 7001, 9)]:
        data = synth.generate(shape, n, level=level, payload=payload, threads=2).tobytes()
        out, status, (lead, tail) = emul.inflate(data, 7, library=L)
        u, _ = oracle.inflate_all(data)
        assert status.tolist() == [0] * len(status) and out.tobytes() == u.tobytes()
        assert (lead == 0xEE).all() and (tail == 0xEE).all()
