"""Hand-crafted BGZF/BAM byte strings for edge-case tests (pure Python + zlib)."""
from __future__ import annotations

import struct
import zlib

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_block(payload: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY, crc: int | None = None,
               isize: int | None = None, raw_deflate: bytes | None = None) -> bytes:
    if raw_deflate is None:
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        raw_deflate = c.compress(payload) + c.flush()
    total = 18 + len(raw_deflate) + 8
    assert total <= 65536
    hdr = bytes.fromhex("1f8b08040000000000ff060042430200") + struct.pack("<H", (total - 1) & 0xFFFF)
    crc = zlib.crc32(payload) if crc is None else crc
    isize = len(payload) if isize is None else isize
    return hdr + raw_deflate + struct.pack("<II", crc & 0xFFFFFFFF, isize & 0xFFFFFFFF)


def bgzf(stream: bytes, payload: int = 0xFF00, level: int = 6, eof: bool = True, **kw) -> bytes:
    out = [bgzf_block(stream[i:i + payload], level, **kw) for i in range(0, len(stream), payload)]
    if eof:
        out.append(BGZF_EOF)
    return b"".join(out)


def bam_header(text: bytes = b"@HD\tVN:1.6\n", refs=(("chr1", 1000000), ("chr2", 500000))) -> bytes:
    h = b"BAM\x01" + struct.pack("<I", len(text)) + text + struct.pack("<I", len(refs))
    for name, ln in refs:
        nb = name.encode() + b"\0"
        h += struct.pack("<I", len(nb)) + nb + struct.pack("<I", ln)
    return h


def bam_record(ref_id=0, pos=100, mapq=60, bin_=4681, flag=0, next_ref=-1, next_pos=-1, tlen=0, name=b"read1",
               cigar=((10, 0),), seq_len=10, qual=None, tags=b"", l_seq_override=None) -> bytes:
    nb = name + b"\0"
    cig = b"".join(struct.pack("<I", (ln << 4) | op) for ln, op in cigar)
    seq = bytes(((1 << 4) | 2) for _ in range((seq_len + 1) // 2))
    q = bytes([30] * seq_len) if qual is None else qual
    l_seq = seq_len if l_seq_override is None else l_seq_override
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(nb) & 0xFF, mapq, bin_, len(cigar), flag, l_seq, next_ref, next_pos,
                       tlen) + nb + cig + seq + q + tags
    return struct.pack("<I", len(body)) + body
