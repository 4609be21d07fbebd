"""CPU oracle for the BGZF -> BAM-record hot path.  TEST INFRASTRUCTURE ONLY.

Two independent restatements of the reference reader live here:

* ``COracle``  — ctypes binding of ``oracle/bam_oracle.c`` (C + zlib), used at every size.
* ``py_*``     — a second, pure-Python/``zlib`` twin written directly from the reference
                 sources; small inputs only.  tests/test_oracle.py checks the two agree.

PARITY UNPINNED (see bam_oracle.c header and DESIGN.md): the Rust reference cannot be
built in this image and its golden md5s need network files.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  File:line citations are relative to
/root/reference/bam_tools/src/.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import struct
import subprocess
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ORC_ERRORS = {
    -1: "BLOCK_TOO_SMALL", -2: "TRUNCATED_BLOCK", -3: "INFLATE", -4: "NOMEM", -5: "BAD_MAGIC",
    -6: "SHORT_HEADER", -7: "SHORT_RECORD", -8: "BAD_FLAGS", -9: "HI_MIXED", -10: "BAD_TAG",
}


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(f"oracle error {code} ({ORC_ERRORS.get(code, '?')})")
        self.code = code


def build(force: bool = False) -> str:
    """Compile oracle/bam_oracle.c -> oracle/_build/liboracle.so (gcc + system zlib)."""
    src = os.path.join(_HERE, "bam_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Columns(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ref_id", "pos", "l_read_name", "mapq", "bin", "n_cigar_op", "flag", "l_seq", "next_ref_id",
        "next_pos", "tlen", "cigar_off", "seq_off", "qual_off", "tags_off", "coord_key", "hi_present", "hi_value")]


class _Digest(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("body_bytes", C.c_uint64), ("fields", C.c_uint64), ("bodies", C.c_uint64)]

    def as_dict(self):
        return {"n_records": self.n_records, "body_bytes": self.body_bytes, "fields": self.fields, "bodies": self.bodies}


COLUMN_DTYPES = {
    "ref_id": np.int32, "pos": np.int32, "l_read_name": np.uint8, "mapq": np.uint8, "bin": np.uint16,
    "n_cigar_op": np.uint16, "flag": np.uint16, "l_seq": np.uint32, "next_ref_id": np.int32,
    "next_pos": np.int32, "tlen": np.int32, "cigar_off": np.uint32, "seq_off": np.uint32,
    "qual_off": np.uint32, "tags_off": np.uint32, "coord_key": np.uint64, "hi_present": np.uint8,
    "hi_value": np.int32,
}


def _u8ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class COracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_inflate_all.restype = C.c_long
        L.orc_inflate_all.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                      C.c_void_p, C.c_size_t]
        L.orc_scan_blocks.restype = C.c_long
        L.orc_scan_blocks.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_read_header.restype = C.c_int
        L.orc_read_header.argtypes = [C.c_void_p, C.c_uint64] + [C.POINTER(C.c_uint64)] * 4
        L.orc_walk_records.restype = C.c_long
        L.orc_walk_records.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_extract.restype = C.c_int
        L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(_Columns), C.c_void_p,
                                  C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_sort_perm.restype = C.c_int
        L.orc_sort_perm.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Columns), C.c_void_p, C.c_size_t, C.c_int,
                                    C.c_void_p]
        L.orc_cpu_reader_run.restype = C.c_int
        L.orc_cpu_reader_run.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_inflate_all_mt.restype = C.c_long
        L.orc_inflate_all_mt.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int]
        L.orc_extract_mt.restype = C.c_int
        L.orc_extract_mt.argtypes = L.orc_extract.argtypes + [C.c_int]
        L.orc_sort_perm_mt.restype = C.c_int
        L.orc_sort_perm_mt.argtypes = L.orc_sort_perm.argtypes + [C.c_int]
        L.orc_digest.restype = C.c_int
        L.orc_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Columns), C.c_void_p, C.c_size_t, C.c_uint64,
                                 C.c_uint64, C.c_int, C.POINTER(_Digest)]
        L.orc_sorted_digest.restype = C.c_int
        L.orc_sorted_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(_Digest)]
        L.orc_decode_field.restype = C.c_int
        L.orc_decode_field.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_field_digest.restype = C.c_int
        L.orc_field_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(_Digest)]

    # -- util.rs:18-71 + readahead.rs:145-150 over the whole file --------------------------
    def inflate_all(self, data) -> tuple[np.ndarray, np.ndarray]:
        """Returns (decompressed stream u8[], per-block inflated sizes u32[])."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        nb = self.lib.orc_scan_blocks(buf.ctypes.data, buf.size, None, 0)
        if nb < 0:
            raise OracleError(nb)
        block_ulen = np.zeros(max(nb, 1), dtype=np.uint32)
        out = C.c_void_p()
        ulen = C.c_uint64()
        rc = self.lib.orc_inflate_all(buf.ctypes.data, buf.size, C.byref(out), C.byref(ulen), block_ulen.ctypes.data, nb)
        if rc < 0:
            raise OracleError(rc)
        u = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(max(ulen.value, 1),))[: ulen.value].copy()
        self.lib.orc_free(out)
        return u, block_ulen[:nb]

    def scan_blocks(self, data) -> np.ndarray:
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        nb = self.lib.orc_scan_blocks(buf.ctypes.data, buf.size, None, 0)
        if nb < 0:
            raise OracleError(nb)
        blocks = np.zeros(max(nb, 1), dtype=np.dtype([("coff", np.uint64), ("bsize", np.uint32), ("_pad", np.uint32)]))
        self.lib.orc_scan_blocks(buf.ctypes.data, buf.size, blocks.ctypes.data, nb)
        return blocks[:nb]

    def read_header(self, u: np.ndarray):
        """Returns (header bytes without magic, offset_to_ref_seqs, stream bytes consumed)."""
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = self.lib.orc_read_header(u.ctypes.data, u.size, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        if rc < 0:
            raise OracleError(rc)
        return bytes(u[a.value: a.value + b.value]), c.value, d.value

    def walk_records(self, u: np.ndarray, start: int):
        n = self.lib.orc_walk_records(u.ctypes.data, u.size, start, None, None, 0)
        if n < 0:
            raise OracleError(n)
        off = np.zeros(max(n, 1), dtype=np.uint64)
        ln = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.orc_walk_records(u.ctypes.data, u.size, start, off.ctypes.data, ln.ctypes.data, n)
        return off[:n], ln[:n]

    def extract(self, u: np.ndarray, rec_off: np.ndarray, rec_len: np.ndarray, want_hi: bool = True):
        n = rec_off.size
        cols = {k: np.zeros(max(n, 1), dtype=dt) for k, dt in COLUMN_DTYPES.items()}
        strand = np.zeros(max(n, 1), dtype=np.uint8)
        cs = _Columns(**{k: v.ctypes.data for k, v in cols.items()})
        bf, bt = C.c_uint64(), C.c_uint64()
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        rec_len = np.ascontiguousarray(rec_len, dtype=np.uint32)
        rc = self.lib.orc_extract(u.ctypes.data, rec_off.ctypes.data, rec_len.ctypes.data, n, C.byref(cs),
                                  strand.ctypes.data, int(want_hi), C.byref(bf), C.byref(bt))
        if rc < 0:
            raise OracleError(rc)
        cols = {k: v[:n] for k, v in cols.items()}
        cols["strand"] = strand[:n]
        return cols, bf.value, bt.value

    def sort_perm(self, u, rec_off, cols, mode: int) -> np.ndarray:
        """mode 0 coord+strand, 1 name, 2 name+mates. Stable order == reference output order."""
        n = rec_off.size
        perm = np.zeros(max(n, 1), dtype=np.uint32)
        cs = _Columns(**{k: np.ascontiguousarray(cols[k]).ctypes.data for k in COLUMN_DTYPES})
        strand = np.ascontiguousarray(cols["strand"])
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        rc = self.lib.orc_sort_perm(u.ctypes.data, rec_off.ctypes.data, C.byref(cs), strand.ctypes.data, n, mode,
                                    perm.ctypes.data)
        if rc < 0:
            raise OracleError(rc)
        return perm[:n]

    def cpu_reader_run(self, data, threads: int, max_records: int = 2**63):
        """Times the reference-topology CPU reader. Returns dict(records, ubytes, seconds, checksum)."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        r, ub, s, ck = C.c_uint64(), C.c_uint64(), C.c_double(), C.c_uint64()
        rc = self.lib.orc_cpu_reader_run(buf.ctypes.data, buf.size, threads, max_records, C.byref(r), C.byref(ub),
                                         C.byref(s), C.byref(ck))
        if rc < 0:
            raise OracleError(rc)
        return {"records": r.value, "ubytes": ub.value, "seconds": s.value, "checksum": ck.value}

    # -- full-size helpers: the same restatement on several threads + the DIGEST SPEC of bam_oracle.c ------------
    def inflate_all_mt(self, data, threads: int) -> np.ndarray:
        """Whole-file inflate on `threads` threads; the returned array owns a C buffer (freed with the array)."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out, ulen = C.c_void_p(), C.c_uint64()
        rc = self.lib.orc_inflate_all_mt(buf.ctypes.data, buf.size, C.byref(out), C.byref(ulen), threads)
        if rc < 0:
            raise OracleError(rc)
        lib = self.lib

        class _Owner:
            def __init__(self, p):
                self.p = p

            def __del__(self):
                lib.orc_free(self.p)

        arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(ulen.value + 64,))
        own = _Owner(out)
        view = arr[: ulen.value]
        self._owners = getattr(self, "_owners", {})
        self._owners[view.ctypes.data] = own   # dropped by release()
        return view

    def release(self, u: np.ndarray):
        getattr(self, "_owners", {}).pop(u.ctypes.data, None)

    def extract_mt(self, u, rec_off, rec_len, want_hi: bool, threads: int):
        n = rec_off.size
        cols = {k: np.empty(max(n, 1), dtype=dt) for k, dt in COLUMN_DTYPES.items()}
        strand = np.empty(max(n, 1), dtype=np.uint8)
        cs = _Columns(**{k: v.ctypes.data for k, v in cols.items()})
        bf, bt = C.c_uint64(), C.c_uint64()
        rc = self.lib.orc_extract_mt(u.ctypes.data, rec_off.ctypes.data, rec_len.ctypes.data, n, C.byref(cs), strand.ctypes.data,
                                     int(want_hi), C.byref(bf), C.byref(bt), threads)
        if rc < 0:
            raise OracleError(rc)
        cols = {k: v[:n] for k, v in cols.items()}
        cols["strand"] = strand[:n]
        return cols, bf.value, bt.value

    def sort_perm_mt(self, u, rec_off, cols, mode: int, threads: int) -> np.ndarray:
        n = rec_off.size
        perm = np.empty(max(n, 1), dtype=np.uint32)
        cs = _Columns(**{k: cols[k].ctypes.data for k in COLUMN_DTYPES})
        rc = self.lib.orc_sort_perm_mt(u.ctypes.data, rec_off.ctypes.data, C.byref(cs), cols["strand"].ctypes.data, n, mode,
                                       perm.ctypes.data, threads)
        if rc < 0:
            raise OracleError(rc)
        return perm[:n]

    def digest(self, u, rec_off, rec_len, cols, threads: int, index_base: int = 0, stream_base: int = 0) -> dict:
        d = _Digest()
        cs = _Columns(**{k: cols[k].ctypes.data for k in COLUMN_DTYPES})
        self.lib.orc_digest(u.ctypes.data, rec_off.ctypes.data, rec_len.ctypes.data, C.byref(cs), cols["strand"].ctypes.data,
                            rec_off.size, index_base, stream_base, threads, C.byref(d))
        return d.as_dict()

    def sorted_digest(self, u, rec_off, rec_len, perm, threads: int) -> dict:
        d = _Digest()
        self.lib.orc_sorted_digest(u.ctypes.data, rec_off.ctypes.data, rec_len.ctypes.data, perm.ctypes.data, perm.size, threads,
                                   C.byref(d))
        r = d.as_dict()
        r["perm"] = r.pop("fields")
        return r

    # -- bamrawrecord.rs:126-157 get_cigar, :223-236 decode_cigar, :259-270 decode_seq, :99 RawQual ----------
    def decode_field(self, u, rec_off, rec_len, field: int):
        """field: 1 CIGAR string, 2 sequence string, 4 raw quality.  Returns (offsets u64[n+1], bytes u8[], records the
        reference would panic on)."""
        n = rec_off.size
        off = np.zeros(n + 1, dtype=np.uint64)
        bad = C.c_uint64()
        self.lib.orc_decode_field(u.ctypes.data, rec_off.ctypes.data, rec_len.ctypes.data, n, field, off.ctypes.data, None, C.byref(bad))
        out = np.zeros(max(int(off[n]), 1), dtype=np.uint8)
        self.lib.orc_decode_field(u.ctypes.data, rec_off.ctypes.data, rec_len.ctypes.data, n, field, off.ctypes.data, out.ctypes.data, C.byref(bad))
        return off, out[: int(off[n])], bad.value

    def field_digest(self, u, rec_off, rec_len, field: int, threads: int) -> dict:
        d = _Digest()
        self.lib.orc_field_digest(u.ctypes.data, rec_off.ctypes.data, rec_len.ctypes.data, rec_off.size, field, threads, C.byref(d))
        r = d.as_dict()
        r["bad"] = r.pop("fields")
        return r

    def crc32(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        return self.lib.orc_crc32(a.ctypes.data, a.size)


# ======================================================================================
# Pure-Python twin (small inputs).  Written from the reference sources, independently of
# bam_oracle.c, so that a mistake in one restatement shows up as a disagreement.
# ======================================================================================

def py_inflate_all(data: bytes) -> tuple[bytes, list[int]]:
    """util.rs:18-71 framing + util.rs:10-16 raw inflate; trailer ignored."""
    out, sizes, pos, n = [], [], 0, len(data)
    while True:
        if n - pos < 18:                                  # read_exact(header) fails -> Ok(0) (util.rs:58-60)
            break
        bs = (struct.unpack_from("<H", data, pos + 16)[0] + 1) & 0xFFFF   # u16 wrap (util.rs:65)
        if bs == 0:
            break
        if bs < 26:
            raise OracleError(-1)
        if n - pos < bs:
            raise OracleError(-2)
        cdata = data[pos + 18: pos + bs - 8]
        d = zlib.decompressobj(-15)
        try:
            u = d.decompress(cdata)
        except zlib.error:
            raise OracleError(-3)
        if not d.eof:
            raise OracleError(-3)
        out.append(u)
        sizes.append(len(u))
        pos += bs
    return b"".join(out), sizes


def py_read_header(u: bytes):
    """reader.rs:87-98,154-200."""
    if len(u) < 4:
        raise OracleError(-6)
    if u[:4] != b"BAM\x01":
        raise OracleError(-5)
    p = 4
    def need(k):
        if len(u) - p < k:
            raise OracleError(-6)
    need(4); l_text = struct.unpack_from("<I", u, p)[0]; p += 4
    need(l_text); p += l_text
    off_refs = p - 4
    need(4); n_ref = struct.unpack_from("<I", u, p)[0]; p += 4
    for _ in range(n_ref):
        need(4); l = struct.unpack_from("<I", u, p)[0]; p += 4
        need(l); p += l
        need(4); p += 4
    return u[4:p], off_refs, p


def py_parse_reference_sequences(b: bytes):
    """reader.rs:101-112,213-228: names must be NUL-terminated UTF-8; NUL stripped."""
    n = struct.unpack_from("<I", b, 0)[0]
    p, out = 4, []
    for _ in range(n):
        l = struct.unpack_from("<I", b, p)[0]; p += 4
        raw = b[p:p + l]; p += l
        if len(raw) != l or not raw.endswith(b"\0") or b"\0" in raw[:-1]:
            raise OracleError(-6)
        out.append((raw[:-1].decode("utf-8"), struct.unpack_from("<I", b, p)[0])); p += 4
    return out


def py_walk_records(u: bytes, start: int):
    """reader.rs:57-81, records.rs:23-29."""
    p, offs, lens = start, [], []
    while True:
        if len(u) - p < 4:
            break
        bs = struct.unpack_from("<I", u, p)[0]
        if bs == 0:
            break
        if len(u) - (p + 4) < bs:
            raise OracleError(-7)
        offs.append(p + 4); lens.append(bs)
        p += 4 + bs
    return offs, lens


_TAG_SIZE = {ord("C"): 1, ord("c"): 1, ord("A"): 1, ord("S"): 2, ord("s"): 2, ord("I"): 4, ord("i"): 4, ord("f"): 4}
_TAG_TYPES = set(b"ABCcfHiISsZ")


def py_get_tag(data: bytes, tag: bytes):
    """tags.rs:62-111. Returns (value bytes, type) or None; raises OracleError(-10) for panics."""
    idx = 0
    while idx < len(data):
        d = data[idx + 2:]
        if idx + 2 > len(data) or len(d) < 1 or d[0] not in _TAG_TYPES:
            raise OracleError(-10)
        t = d[0]
        if t == ord("B"):
            if len(d) < 2 or d[1] not in _TAG_TYPES or d[1] not in _TAG_SIZE or len(d) < 6:
                raise OracleError(-10)
            nbytes = struct.unpack_from("<I", d, 2)[0] * _TAG_SIZE[d[1]]
            if len(d) < 6 + nbytes:
                raise OracleError(-10)
            val, used = d[6:6 + nbytes], 6 + nbytes
        elif t in (ord("Z"), ord("H")):
            z = d.find(b"\0", 1)
            if z < 0:
                raise OracleError(-10)
            val, used = d[1:z], z + 1
        else:
            sz = _TAG_SIZE[t]
            if len(d) < 1 + sz:
                raise OracleError(-10)
            val, used = d[1:1 + sz], 1 + sz
        if data[idx:idx + 2] == tag:
            return val, t
        idx += 2 + used
    return None


def py_hit_count(body: bytes):
    """tags.rs:113-129 via bamrawrecord.rs:80-104 (u8-wrapping tags offset)."""
    l_name = body[8]
    cig = (32 + l_name) & 0xFF
    n_cig = struct.unpack_from("<H", body, 12)[0]
    l_seq = struct.unpack_from("<I", body, 16)[0]
    tags = cig + 4 * n_cig + (((l_seq + 1) & 0xFFFFFFFF) // 2) + l_seq
    if tags > len(body):
        raise OracleError(-10)
    r = py_get_tag(body[tags:], b"HI")
    if r is None:
        return None
    val, t = r
    fmt = {ord("A"): "<B", ord("C"): "<B", ord("c"): "<b", ord("s"): "<h", ord("S"): "<H", ord("i"): "<i", ord("I"): "<i"}.get(t)
    if fmt is None:
        raise OracleError(-10)
    return struct.unpack_from(fmt, val, 0)[0]


_CIG_OPS = "MIDNSHP=X"
_SEQ_BASES = "=ACMGRSVTWYHKDBN"


def py_decode_field(body: bytes, field: int) -> bytes:
    """One record's field the way the reference's accessors produce it: 1 decode_cigar(get_bytes(RawCigar))
    (bamrawrecord.rs:223-236 over :126-157), 2 decode_seq(get_bytes(RawSequence)) (:259-270), 4 get_bytes(RawQual) (:99).
    Raises OracleError(-40) where the reference panics."""
    if len(body) < 32:
        raise OracleError(-40)
    ref, pos = struct.unpack_from("<ii", body, 0)
    n_cig = struct.unpack_from("<H", body, 12)[0]
    l_seq = struct.unpack_from("<I", body, 16)[0]
    nseq = ((l_seq + 1) & 0xFFFFFFFF) // 2
    cig = (32 + body[8]) & 0xFF                      # u8 arithmetic in the get_bytes closures (:80)
    seq = cig + 4 * n_cig
    qual = seq + nseq
    tags = qual + l_seq
    if field == 2:
        if seq + nseq > len(body):
            raise OracleError(-40)
        out = []
        for b in body[seq:seq + nseq]:
            out.append(_SEQ_BASES[b >> 4])
            if b & 15:
                out.append(_SEQ_BASES[b & 15])
        return "".join(out).encode()
    if field == 4:
        if qual + l_seq > len(body):
            raise OracleError(-40)
        return body[qual:qual + l_seq]
    # get_cigar
    if ref < 0 or pos < 0 or n_cig == 0:
        data = b""
    else:
        if cig + 4 * n_cig > len(body):
            raise OracleError(-40)
        data = body[cig:cig + 4 * n_cig]
        first = struct.unpack_from("<I", data, 0)[0]
        if (first & 0xF) == 4 and (first >> 4) == nseq and n_cig == 2:
            if tags > len(body):
                raise OracleError(-40)
            try:
                r = py_get_tag(body[tags:], b"CG")
            except OracleError:
                raise OracleError(-40)
            if r is not None:
                data = r[0]
    out = []
    for i in range(0, len(data), 4):
        if len(data) - i < 4:
            raise OracleError(-40)
        v = struct.unpack_from("<I", data, i)[0]
        if (v & 15) > 8:
            raise OracleError(-40)
        out.append(str(v >> 4) + _CIG_OPS[v & 15])
    return "".join(out).encode()


def py_extract(u: bytes, offs, lens, want_hi=True):
    """bamrawrecord.rs:49-105; comparators.rs:41-59."""
    rows = []
    for o, ln in zip(offs, lens):
        b = u[o:o + ln]
        ref_id, pos, l_name, mapq, bin_, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", b, 0)
        cig = 32 + l_name
        seq = cig + 4 * n_cig
        qual = seq + (l_seq + 1) // 2
        tags = qual + l_seq
        r = 0x7FFFFFFF if ref_id == -1 else ref_id
        key = (((r & 0xFFFFFFFF) ^ 0x80000000) << 32) | ((pos & 0xFFFFFFFF) ^ 0x80000000)
        hi = None
        bad_tag = False
        if want_hi:
            try:
                hi = py_hit_count(b)
            except OracleError:
                bad_tag = True
        rows.append(dict(ref_id=ref_id, pos=pos, l_read_name=l_name, mapq=mapq, bin=bin_, n_cigar_op=n_cig, flag=flag,
                         l_seq=l_seq, next_ref_id=nref, next_pos=npos, tlen=tlen, cigar_off=cig, seq_off=seq,
                         qual_off=qual, tags_off=tags, coord_key=key, strand=1 if flag & 0x10 else 0,
                         hi_present=0 if hi is None else 1, hi_value=0 if hi is None else hi, bad_tag=bad_tag,
                         name=b[32:32 + l_name]))
    return rows


def py_sort_perm(rows, mode: int):
    """comparators.rs:61-147 as a stable sort (sort.rs:366 par_sort_by is stable; merge tie -> chunk index :487-510)."""
    def cmp(a, b):
        ra, rb = rows[a], rows[b]
        if mode == 0:
            ka, kb = (ra["coord_key"], ra["strand"]), (rb["coord_key"], rb["strand"])
            return (ka > kb) - (ka < kb)
        c = (ra["name"] > rb["name"]) - (ra["name"] < rb["name"])
        if c or mode == 1:
            return c
        ha = ra["hi_value"] if ra["hi_present"] else None
        hb = rb["hi_value"] if rb["hi_present"] else None
        if ha == hb:
            return (ra["flag"] > rb["flag"]) - (ra["flag"] < rb["flag"])
        if ha is None or hb is None:
            raise OracleError(-9)
        return (ha > hb) - (ha < hb)
    return sorted(range(len(rows)), key=functools.cmp_to_key(cmp))


# ---- DIGEST SPEC, pure-Python twin (small inputs): the third statement of the same function, next to
# oracle/bam_oracle.c (orc_digest / orc_sorted_digest) and bam_parallel_b200/csrc/digest.cu ----
_M64 = (1 << 64) - 1


def py_mix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def py_body_hash(body: bytes) -> int:
    h = 0
    for k in range(0, (len(body) + 3) // 4):
        w = int.from_bytes(body[4 * k:4 * k + 4].ljust(4, b"\0"), "little")
        h = (h + py_mix64(w | (k << 32))) & _M64
    return h


def py_digest(u: bytes, rec_off, rec_len, cols, index_base: int = 0, stream_base: int = 0) -> dict:
    order = ["ref_id", "pos", "l_read_name", "mapq", "bin", "n_cigar_op", "flag", "l_seq", "next_ref_id", "next_pos", "tlen",
             "cigar_off", "seq_off", "qual_off", "tags_off"]
    d = {"n_records": len(rec_off), "body_bytes": 0, "fields": 0, "bodies": 0}
    for i, (o, l) in enumerate(zip(rec_off, rec_len)):
        o, l = int(o), int(l)
        idx = index_base + i
        h = 0
        vals = [idx, stream_base + o, l] + [int(cols[k][i]) & 0xFFFFFFFF for k in order]
        vals += [int(cols["coord_key"][i]), int(cols["strand"][i])]
        hp = int(cols["hi_present"][i]) == 1
        vals += [1 if hp else 0, (int(cols["hi_value"][i]) & 0xFFFFFFFF) if hp else 0]
        for v in vals:
            h = py_mix64(h ^ (v & _M64))
        d["fields"] = (d["fields"] + h) & _M64
        d["bodies"] = (d["bodies"] + py_mix64((py_body_hash(u[o:o + l]) + idx * 0x9E3779B97F4A7C15) & _M64)) & _M64
        d["body_bytes"] += l
    return d


def py_sorted_digest(u: bytes, rec_off, rec_len, perm) -> dict:
    d = {"n_records": len(perm), "body_bytes": 0, "perm": 0, "bodies": 0}
    for i, r in enumerate(perm):
        r = int(r)
        o, l = int(rec_off[r]), int(rec_len[r])
        d["perm"] = (d["perm"] + py_mix64(r | (i << 32))) & _M64
        d["bodies"] = (d["bodies"] + py_mix64((py_body_hash(u[o:o + l]) + i * 0x9E3779B97F4A7C15) & _M64)) & _M64
        d["body_bytes"] += l
    return d
