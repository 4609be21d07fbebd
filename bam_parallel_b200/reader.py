"""Host-side mirror of the reference's reader API over the C ABI (include/bamgpu.h).

Names, argument meaning and error behaviour follow ``bam_tools::Reader``
(/root/reference/bam_tools/src/reader.rs, reader/records.rs):

    reader = Reader.new(inner, thread_num, track_progress=None)   # reader.rs:28-55
    header_bytes, offset_to_ref_seqs = reader.read_header()       # reader.rs:87-98
    records = reader.records()                                    # reader.rs:83
    rec = records.next_rec()            # Option<&Vec<u8>> -> memoryview | None   records.rs:23-29
    for rec in reader.records(): ...    # cloning iterator -> bytes               records.rs:32-42
    n = reader.read_record(buf)         # clears then appends; 0 = EOF            reader.rs:57-61
    n = reader.append_record(buf)       # reader.rs:63-73
    data = reader.read(n)               # impl Read                               reader.rs:114-152
    parse_reference_sequences(header_bytes[offset_to_ref_seqs:])  # reader.rs:101-112

Where the reference panics (corrupt block, short record body) this raises BamGpuError; where it
ends silently (short tail, block_size 0, BSIZE wrap) this ends silently too.

The record bytes, the columns and the keys all come from the CUDA kernels in libbamgpu.so.
There is no CPU decode path here; importing works without a GPU but constructing a Reader does not.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (COLUMN_FIELDS, COPY_COLUMNS, COPY_DATA, NO_CRC, NO_RECORDS, WANT_HI, BamGpuError, Batch, Block, Digest,
                   EdgeHandle, Opts, Sorted, Stats, make_opts)

_NP = {C.c_uint64: np.uint64, C.c_uint32: np.uint32, C.c_int32: np.int32, C.c_uint16: np.uint16, C.c_uint8: np.uint8}


def _view(ptr, n, ctype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=_NP[ctype])
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,))


class RecordBatch:
    """All records of one pipeline batch: device pointers plus (when copied) host numpy views.
    Views are valid until the next batch is requested from the same Reader / Engine."""

    def __init__(self, b: Batch):
        self.n_records = int(b.n_records)
        self.data_len = int(b.data_len)
        self.first_record_index = int(b.first_record_index)
        self.bad_flag_records = int(b.bad_flag_records)
        self.chain_start = int(b.chain_start)
        self.chain_repairs = int(b.chain_repairs)
        self.data_start = int(b.data_start)        # Reader batches: bytes [data_start, data_len) are the inflated stream
        self.stream_offset = int(b.stream_offset)  # position of data[data_start] in the whole inflated stream
        self.dev_data = b.data
        self.dev = {n: getattr(b.dev, n) for n, _ in COLUMN_FIELDS}
        self.data = _view(b.h_data, self.data_len, C.c_uint8) if b.h_data else None
        self.columns = None
        if self.n_records == 0:
            self.columns = {n: np.zeros(0, dtype=_NP[ct]) for n, ct in COLUMN_FIELDS}
        elif b.host.rec_off:   # BAMGPU_COPY_OFFSETS alone fills only rec_off / rec_len
            self.columns = {n: _view(getattr(b.host, n), self.n_records, ct) for n, ct in COLUMN_FIELDS if getattr(b.host, n)}

    def body(self, i: int) -> memoryview:
        o = int(self.columns["rec_off"][i])
        return memoryview(self.data)[o:o + int(self.columns["rec_len"][i])]


class Records:
    """reader/records.rs:10-42."""

    def __init__(self, reader: "Reader"):
        self._r = reader

    def next_rec(self) -> Optional[memoryview]:
        """Lends the next record body (valid until the next call); None at EOF."""
        L, r = self._r._L, self._r
        body, n = C.c_void_p(), C.c_size_t()
        rc = L.bamgpu_next_rec(r._h, C.byref(body), C.byref(n))
        if rc == _lib.EOF:
            return None
        if rc < 0:
            raise BamGpuError(rc, L.bamgpu_last_error(r._h).decode())
        return memoryview((C.c_uint8 * n.value).from_address(body.value))

    def __iter__(self):
        return self

    def __next__(self) -> bytes:
        rec = self.next_rec()
        if rec is None:
            raise StopIteration
        return bytes(rec)


class Reader:
    def __init__(self, inner, thread_num: int = 4, track_progress: Optional[int] = None, *, flags: int = 0,
                 batch_bytes: int = 0, device: int = -1, devices=None, jobs_per_device: int = 0, inflate_impl: int = 0):
        """`devices`: list of CUDA ordinals - batches go to them round robin and the partial record at a batch edge moves
        GPU to GPU (bamgpu_opts.devices).  `inner`: bytes-like / numpy array (a page-locked array is read
        by the GPUs directly, without a staging copy) or an object with read()/readinto()."""
        self._L = L = _lib.load()
        self._inner = inner
        opts = make_opts(device, flags, batch_bytes, inflate_impl, jobs_per_device, devices)
        tp = C.c_uint64(track_progress) if track_progress is not None else None
        self._keep = None
        if isinstance(inner, (bytes, bytearray, memoryview, np.ndarray)):
            arr = np.frombuffer(inner, dtype=np.uint8) if not isinstance(inner, np.ndarray) else inner
            self._keep = arr
            self._h = L.bamgpu_reader_from_memory(arr.ctypes.data, arr.size, thread_num, C.byref(opts))
        else:
            readinto = getattr(inner, "readinto", None)

            def _read(_ctx, dst, cap):
                try:
                    if readinto is not None:
                        buf = (C.c_uint8 * cap).from_address(C.addressof(dst.contents))
                        n = readinto(buf)
                        return int(n or 0)
                    chunk = inner.read(cap)
                    C.memmove(dst, chunk, len(chunk))
                    return len(chunk)
                except Exception:
                    return -1

            self._cb = _lib.READ_FN(_read)
            self._h = L.bamgpu_reader_new(self._cb, None, thread_num, C.byref(tp) if tp is not None else None, C.byref(opts))
        if not self._h:
            raise BamGpuError(-2, "could not create reader: no usable CUDA device (there is no CPU fallback)")

    # Reader::new(inner, thread_num, track_progress)
    @classmethod
    def new(cls, inner, thread_num: int = 4, track_progress: Optional[int] = None, **kw) -> "Reader":
        return cls(inner, thread_num, track_progress, **kw)

    def _check(self, rc):
        if rc < 0:
            raise BamGpuError(rc, self._L.bamgpu_last_error(self._h).decode())
        return rc

    def read_header(self):
        p, n, off = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(self._L.bamgpu_read_header(self._h, C.byref(p), C.byref(n), C.byref(off)))
        return C.string_at(p.value, n.value), off.value

    def records(self) -> Records:
        return Records(self)

    def append_record(self, buf: bytearray) -> int:
        need = C.c_size_t()
        rc = self._L.bamgpu_append_record(self._h, None, 0, C.byref(need))
        if rc == _lib.OK and need.value == 0:
            return 0
        if rc != -23:
            self._check(rc)
        n = need.value
        tmp = (C.c_uint8 * n)()
        self._check(self._L.bamgpu_append_record(self._h, tmp, n, C.byref(need)))
        buf += bytes(tmp)
        return n

    def read_record(self, buf: bytearray) -> int:
        del buf[:]
        return self.append_record(buf)

    def read(self, n: int = -1) -> bytes:
        out = bytearray()
        chunk = 1 << 20 if n < 0 else n
        tmp = (C.c_uint8 * chunk)()
        while n < 0 or len(out) < n:
            want = chunk if n < 0 else min(chunk, n - len(out))
            got = self._L.bamgpu_read(self._h, tmp, want)
            if got < 0:
                raise BamGpuError(got, self._L.bamgpu_last_error(self._h).decode())
            if got == 0:
                break
            out += bytes(tmp[:got]) if got < chunk else bytes(tmp)
        return bytes(out)

    def next_batch(self) -> Optional[RecordBatch]:
        """GPU-native interface: every record of the next batch as structure-of-arrays columns."""
        b = Batch()
        rc = self._L.bamgpu_next_batch(self._h, C.byref(b))
        if rc == _lib.EOF:
            return None
        self._check(rc)
        return RecordBatch(b)

    def stats(self) -> dict:
        s = Stats()
        self._L.bamgpu_reader_stats(self._h, C.byref(s))
        return s.asdict()

    def batch_digest(self, b: "RecordBatch") -> dict:
        """DIGEST SPEC (csrc/digest.cu) of the batch next_batch() returned last, with global record indices and stream offsets."""
        e = self._L.bamgpu_reader_current_engine(self._h)
        d = Digest()
        base = (b.stream_offset - b.data_start) & 0xFFFFFFFFFFFFFFFF
        rc = self._L.bamgpu_engine_digest(e, b.first_record_index, base, None, C.byref(d))
        if rc < 0:
            raise BamGpuError(rc, self._L.bamgpu_engine_last_error(e).decode())
        return d.asdict()

    def close(self):
        if getattr(self, "_h", None):
            self._L.bamgpu_reader_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def parse_reference_sequences(b: bytes):
    """reader.rs:101-112: [(name without NUL, l_ref)] from header bytes starting at n_ref."""
    L = _lib.load()
    out = []

    @_lib.REFSEQ_CB
    def cb(name, n, l_ref, _u):
        out.append((C.string_at(name, n).decode("utf-8"), l_ref))

    buf = (C.c_uint8 * len(b)).from_buffer_copy(b) if len(b) else None
    rc = L.bamgpu_parse_reference_sequences(buf, len(b), cb, None)
    if rc < 0:
        raise BamGpuError(rc, "parse_reference_sequences")
    return out


def scan_blocks(data, final: bool = True):
    """bamgpu_scan_blocks over a bytes-like object -> (ctypes Block array, n, consumed, total_isize, rc)."""
    L = _lib.load()
    arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n, consumed, tot = C.c_size_t(), C.c_size_t(), C.c_uint64()
    rc = L.bamgpu_scan_blocks(arr.ctypes.data, arr.size, int(final), None, 0, C.byref(n), C.byref(consumed), C.byref(tot))
    blocks = (Block * max(n.value, 1))()
    if n.value:
        L.bamgpu_scan_blocks(arr.ctypes.data, arr.size, int(final), blocks, n.value, C.byref(n), C.byref(consumed), C.byref(tot))
    return blocks, n.value, consumed.value, tot.value, rc


class Engine:
    """Device-resident decode (bamgpu_engine_*): what Reader runs per batch and what bench.py times."""

    def __init__(self, device: int = -1, flags: int = 0, inflate_impl: int = 0):
        self._L = L = _lib.load()
        opts = make_opts(device, flags, 0, inflate_impl)
        self._h = L.bamgpu_engine_new(C.byref(opts))
        if not self._h:
            raise BamGpuError(-2, "could not create engine: no usable CUDA device (there is no CPU fallback)")
        self._dev_bufs = []

    def upload(self, data) -> int:
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        d = self._L.bamgpu_device_alloc(self._h, max(arr.size, 1))
        if not d:
            raise BamGpuError(-3, "device alloc")
        self._dev_bufs.append(d)
        if arr.size:
            rc = self._L.bamgpu_memcpy_h2d(self._h, d, arr.ctypes.data, arr.size)
            if rc < 0:
                raise BamGpuError(rc, self._L.bamgpu_engine_last_error(self._h).decode())
        return d

    def free(self, d):
        self._dev_bufs.remove(d)
        self._L.bamgpu_device_free(self._h, d)

    def decode(self, d_comp, blocks, n_blocks, stream_start, final=True, flags=COPY_DATA | COPY_COLUMNS, cuda_stream=None):
        b, tail = Batch(), C.c_uint64()
        rc = self._L.bamgpu_engine_decode(self._h, d_comp, blocks, n_blocks, stream_start, int(final), flags,
                                          cuda_stream, C.byref(b), C.byref(tail))
        if rc < 0:
            raise BamGpuError(rc, self._L.bamgpu_engine_last_error(self._h).decode())
        return RecordBatch(b), tail.value, rc

    # ---- the two halves of decode(), for block-range sharding over several GPUs (include/bamgpu.h) ----
    def _eck(self, rc):
        if rc < 0:
            raise BamGpuError(rc, self._L.bamgpu_engine_last_error(self._h).decode())
        return rc

    def inflate(self, d_comp, blocks, n_blocks, flags=0, cuda_stream=None):
        self._eck(self._L.bamgpu_engine_inflate(self._h, d_comp, blocks, n_blocks, flags, cuda_stream))

    def data(self):
        """(device pointer, length) of the engine's own inflated bytes."""
        p, n = C.c_void_p(), C.c_uint64()
        self._eck(self._L.bamgpu_engine_data(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def append_halo(self, d_src, nbytes, cuda_stream=None):
        self._eck(self._L.bamgpu_engine_append_halo(self._h, d_src, nbytes, cuda_stream))

    def records(self, stream_start, final=True, flags=0, cuda_stream=None):
        b, tail = Batch(), C.c_uint64()
        rc = self._eck(self._L.bamgpu_engine_records(self._h, stream_start, int(final), flags, cuda_stream, C.byref(b), C.byref(tail)))
        return RecordBatch(b), tail.value, rc

    def sort(self, sort_by: int, flags=COPY_DATA | COPY_COLUMNS, cuda_stream=None):
        """bamgpu_engine_sort: stable sort of the last decoded batch by the reference's comparator (sort.rs:92-96 SortBy,
        comparators.rs:61-147).  Returns (perm as uint32 numpy or None, sorted bodies as uint8 numpy or None, Sorted struct):
        the bodies are the bytes bam_binary writes (sort.rs:643)."""
        srt = Sorted()
        self._eck(self._L.bamgpu_engine_sort(self._h, sort_by, flags, cuda_stream, C.byref(srt)))
        n, nb = int(srt.n_records), int(srt.bodies_len)
        perm = (_view(srt.h_perm, n, C.c_uint32) if srt.h_perm else np.zeros(0, np.uint32)) if flags & COPY_COLUMNS else None
        bodies = (_view(srt.h_bodies, nb, C.c_uint8) if srt.h_bodies else np.zeros(0, np.uint8)) if flags & COPY_DATA else None
        return perm, bodies, srt

    def digest(self, index_base: int = 0, stream_base: int = 0, cuda_stream=None) -> dict:
        """DIGEST SPEC of the last decoded batch, from the engine's columns and inflated bytes on the device."""
        d = Digest()
        self._eck(self._L.bamgpu_engine_digest(self._h, index_base, stream_base & 0xFFFFFFFFFFFFFFFF, cuda_stream, C.byref(d)))
        return d.asdict()

    def sorted_digest(self, cuda_stream=None) -> dict:
        """DIGEST SPEC of the last sort: {'perm': permutation digest, 'bodies': digest of the gathered output bytes}."""
        d = Digest()
        self._eck(self._L.bamgpu_engine_sorted_digest(self._h, cuda_stream, C.byref(d)))
        r = d.asdict()
        r["perm"] = r.pop("fields")
        return r

    def decode_fields(self, which: int = 7, flags: int = COPY_DATA, cuda_stream=None) -> dict:
        """bamgpu_engine_decode_fields: CIGAR strings (decode_cigar over get_cigar, bamrawrecord.rs:126-157,223-236), sequence
        strings (decode_seq, :259-270) and raw qualities (:99) of the last decoded batch.  Returns {"cigar"|"seq"|"qual":
        {"off": uint64 numpy n+1 (host, with COPY_DATA), "bytes": uint8 numpy, "total", "bad"}, "n_records", "ms_measure", "ms_decode"}."""
        from ._lib import Fields
        f = Fields()
        self._eck(self._L.bamgpu_engine_decode_fields(self._h, which, flags, cuda_stream, C.byref(f)))
        n = int(f.n_records)
        out = {"n_records": n, "ms_measure": f.ms_measure, "ms_decode": f.ms_decode}
        for k, name in ((1, "cigar"), (2, "seq"), (4, "qual")):
            if not which & k:
                continue
            fl = getattr(f, name)
            d = {"total": int(fl.total), "bad": int(fl.bad), "off": None, "bytes": None}
            if flags & COPY_DATA:
                d["off"] = _view(fl.h_off, n + 1, C.c_uint64)
                d["bytes"] = _view(fl.h_bytes, int(fl.total), C.c_uint8) if fl.total else np.zeros(0, np.uint8)
            out[name] = d
        return out

    def field_digest(self, field: int, cuda_stream=None) -> dict:
        """bodies-digest (DIGEST SPEC) of one field decode_fields produced, computed on the device."""
        d = Digest()
        self._eck(self._L.bamgpu_engine_field_digest(self._h, field, cuda_stream, C.byref(d)))
        r = d.asdict()
        r.pop("fields")
        return r

    # ---- shard-edge exchange over peer memory (bamgpu_engine_edge_*) ----
    def edge_open(self) -> bytes:
        h = EdgeHandle()
        self._eck(self._L.bamgpu_engine_edge_open(self._h, C.byref(h)))
        return bytes(h)

    def edge_connect(self, left: Optional[bytes], right: Optional[bytes]):
        hl = EdgeHandle.from_buffer_copy(left) if left else None
        hr = EdgeHandle.from_buffer_copy(right) if right else None
        self._eck(self._L.bamgpu_engine_edge_connect(self._h, C.byref(hl) if hl else None, C.byref(hr) if hr else None))

    def edge_exchange(self, step: int, cuda_stream=None):
        self._eck(self._L.bamgpu_engine_edge_exchange(self._h, step, cuda_stream))

    def bgzf_store(self, data, eof_marker: bool = True) -> bytes:
        """bamgpu_engine_bgzf_store: `data` (bytes-like, uploaded here) wrapped into stored-block BGZF; returns the BGZF bytes."""
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        d_src = self.upload(arr)
        bound = self._L.bamgpu_bgzf_bound(arr.size)
        d_out = self._L.bamgpu_device_alloc(self._h, bound)
        try:
            n = C.c_size_t()
            self._eck(self._L.bamgpu_engine_bgzf_store(self._h, d_src, arr.size, d_out, bound, C.byref(n), int(eof_marker), None))
            out = np.empty(n.value, dtype=np.uint8)
            if n.value:
                self._eck(self._L.bamgpu_memcpy_d2h(self._h, out.ctypes.data, d_out, n.value))
            return out.tobytes()
        finally:
            self._L.bamgpu_device_free(self._h, d_out)
            self.free(d_src)

    def bgzf_write(self, data, level: int = 1, eof_marker: bool = True) -> bytes:
        """bamgpu_engine_bgzf_write: `data` wrapped into BGZF with compressed payloads (K6; level 0 = stored); returns the BGZF bytes."""
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        d_src = self.upload(arr)
        bound = self._L.bamgpu_bgzf_bound(arr.size)
        d_out = self._L.bamgpu_device_alloc(self._h, bound)
        try:
            n = C.c_size_t()
            self._eck(self._L.bamgpu_engine_bgzf_write(self._h, d_src, arr.size, d_out, bound, C.byref(n), int(eof_marker), int(level), None))
            out = np.empty(n.value, dtype=np.uint8)
            if n.value:
                self._eck(self._L.bamgpu_memcpy_d2h(self._h, out.ctypes.data, d_out, n.value))
            return out.tobytes()
        finally:
            self._L.bamgpu_device_free(self._h, d_out)
            self.free(d_src)

    def set_n_ref(self, n_ref: int):
        """Optional hint (header n_ref) for K3's guess of record starts; results never depend on it."""
        self._L.bamgpu_engine_set_n_ref(self._h, n_ref)

    def stats(self) -> dict:
        s = Stats()
        self._L.bamgpu_engine_stats(self._h, C.byref(s))
        return s.asdict()

    def reset_stats(self):
        self._L.bamgpu_engine_reset_stats(self._h)

    def close(self):
        if getattr(self, "_h", None):
            for d in self._dev_bufs:
                self._L.bamgpu_device_free(self._h, d)
            self._dev_bufs = []
            self._L.bamgpu_engine_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
