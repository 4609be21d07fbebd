"""ctypes binding of libbamgpu.so (include/bamgpu.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("BAMGPU_LIB") or os.path.join(_ROOT, "bam_parallel_b200", "lib", "libbamgpu.so")  # BAMGPU_LIB: kernel-variant experiments only

OK, EOF = 0, 1
ERR_NAMES = {
    -1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_NOMEM", -4: "ERR_IO", -10: "ERR_BLOCK_TOO_SMALL", -11: "ERR_TRUNCATED_BLOCK",
    -12: "ERR_INFLATE", -13: "ERR_ISIZE", -14: "ERR_CRC", -20: "ERR_BAD_MAGIC", -21: "ERR_SHORT_HEADER",
    -22: "ERR_SHORT_RECORD", -23: "ERR_BUFFER_TOO_SMALL", -24: "ERR_STATE", -30: "ERR_SORT_HI_MIXED",
    -31: "ERR_SORT_BAD_FLAGS", -32: "ERR_SORT_BAD_TAGS", -33: "ERR_TOO_LARGE",
}
SORT_NAME, SORT_NAME_AND_MATCH_MATES, SORT_COORDINATES_AND_STRAND = 0, 1, 2   # sorting/sort.rs:92-96 SortBy
COPY_DATA, COPY_COLUMNS, WANT_HI, NO_CRC, NO_RECORDS, SPECULATIVE_START, COPY_OFFSETS = 1, 2, 4, 8, 16, 32, 64
DEFER_CRC = 256
MAX_HALO = 1 << 20


class BamGpuError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"bamgpu {ERR_NAMES.get(code, code)} ({code}): {msg}")
        self.code = code


class Block(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("out_off", C.c_uint64), ("in_len", C.c_uint32), ("isize", C.c_uint32),
                ("crc32", C.c_uint32), ("reserved", C.c_uint32)]


COLUMN_FIELDS = [
    ("rec_off", C.c_uint64), ("rec_len", C.c_uint32), ("ref_id", C.c_int32), ("pos", C.c_int32),
    ("l_read_name", C.c_uint8), ("mapq", C.c_uint8), ("bin", C.c_uint16), ("n_cigar_op", C.c_uint16),
    ("flag", C.c_uint16), ("l_seq", C.c_uint32), ("next_ref_id", C.c_int32), ("next_pos", C.c_int32),
    ("tlen", C.c_int32), ("cigar_off", C.c_uint32), ("seq_off", C.c_uint32), ("qual_off", C.c_uint32),
    ("tags_off", C.c_uint32), ("coord_key", C.c_uint64), ("strand", C.c_uint8), ("hi_present", C.c_uint8),
    ("hi_value", C.c_int32),
]


class Columns(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in COLUMN_FIELDS]


class Batch(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("data_len", C.c_uint64), ("data", C.c_void_p), ("h_data", C.c_void_p),
                ("dev", Columns), ("host", Columns), ("first_record_index", C.c_uint64), ("bad_flag_records", C.c_uint64),
                ("chain_start", C.c_uint64), ("chain_repairs", C.c_uint64), ("data_start", C.c_uint64),
                ("stream_offset", C.c_uint64)]


MAX_DEVICES = 16


class Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("batch_bytes", C.c_uint64), ("inflate_impl", C.c_int32),
                ("jobs_per_device", C.c_int32), ("n_devices", C.c_int32), ("devices", C.c_int32 * MAX_DEVICES),
                ("sort_output", C.c_int32), ("sort_mem_limit", C.c_uint64), ("out_compr_level", C.c_int32), ("reserved", C.c_int32)]


OUT_BODIES, OUT_RECORDS, OUT_BAM = 0, 1, 2          # bamgpu_opts.sort_output
SORT_EMIT_BLOCK_SIZE = 128


def make_opts(device: int = -1, flags: int = 0, batch_bytes: int = 0, inflate_impl: int = 0, jobs_per_device: int = 0,
              devices=None, sort_output: int = 0, sort_mem_limit: int = 0, out_compr_level: int = 0) -> Opts:
    o = Opts()
    o.out_compr_level = out_compr_level
    o.sort_output = sort_output
    o.sort_mem_limit = sort_mem_limit
    o.device, o.flags, o.batch_bytes, o.inflate_impl, o.jobs_per_device = device, flags, batch_bytes, inflate_impl, jobs_per_device
    devices = list(devices or [])
    if len(devices) > MAX_DEVICES:
        raise ValueError(f"at most {MAX_DEVICES} devices")
    o.n_devices = len(devices)
    for i, d in enumerate(devices):
        o.devices[i] = d
    return o


class Digest(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("body_bytes", C.c_uint64), ("fields", C.c_uint64), ("bodies", C.c_uint64)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EdgeHandle(C.Structure):
    _fields_ = [("ipc", C.c_uint8 * 64), ("ptr", C.c_uint64), ("device", C.c_int32), ("pid", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("blocks", "bytes_in", "bytes_out", "records", "batches", "kernel_launches")] + \
               [(n, C.c_double) for n in ("ms_h2d", "ms_inflate", "ms_crc", "ms_records", "ms_d2h", "ms_total", "ms_sort",
                                          "ms_gather")] + [("p2p_bytes", C.c_uint64), ("ms_bgzf_write", C.c_double)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Sorted(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("bodies_len", C.c_uint64), ("perm", C.c_void_p), ("body_off", C.c_void_p),
                ("bodies", C.c_void_p), ("h_perm", C.c_void_p), ("h_bodies", C.c_void_p)]


FIELD_CIGAR, FIELD_SEQ, FIELD_QUAL = 1, 2, 4       # bamgpu_engine_decode_fields


class Field(C.Structure):
    _fields_ = [("total", C.c_uint64), ("bad", C.c_uint64), ("off", C.c_void_p), ("bytes", C.c_void_p), ("h_off", C.c_void_p),
                ("h_bytes", C.c_void_p)]


class Fields(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("cigar", Field), ("seq", Field), ("qual", Field), ("ms_measure", C.c_double),
                ("ms_decode", C.c_double)]


READ_FN = C.CFUNCTYPE(C.c_long, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)
WRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)
REFSEQ_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p)

# every symbol include/bamgpu.h declares
EXPORTS = [
    "bamgpu_scan_blocks", "bamgpu_engine_new", "bamgpu_engine_free", "bamgpu_engine_last_error", "bamgpu_device_alloc",
    "bamgpu_device_free", "bamgpu_pinned_alloc", "bamgpu_pinned_free", "bamgpu_memcpy_h2d", "bamgpu_memcpy_d2h",
    "bamgpu_engine_decode", "bamgpu_engine_inflate", "bamgpu_engine_data", "bamgpu_engine_append_halo",
    "bamgpu_engine_records", "bamgpu_engine_set_n_ref", "bamgpu_engine_stats", "bamgpu_engine_reset_stats", "bamgpu_reader_new",
    "bamgpu_reader_from_memory", "bamgpu_reader_free", "bamgpu_last_error", "bamgpu_read_header", "bamgpu_next_rec",
    "bamgpu_append_record", "bamgpu_read", "bamgpu_next_batch", "bamgpu_reader_stats",
    "bamgpu_parse_reference_sequences", "bamgpu_status_str", "bamgpu_build_info", "bamgpu_engine_sort", "bamgpu_sort_bam", "bamgpu_sort_bam_from_memory",
    "bamgpu_engine_digest", "bamgpu_engine_sorted_digest", "bamgpu_reader_current_engine", "bamgpu_engine_edge_open",
    "bamgpu_engine_edge_connect", "bamgpu_engine_edge_exchange", "bamgpu_engine_edge_close", "bamgpu_bgzf_bound",
    "bamgpu_engine_bgzf_store", "bamgpu_engine_bgzf_write", "bamgpu_engine_decode_fields", "bamgpu_engine_field_digest",
]

_lib = None


def load() -> C.CDLL:
    """Loads libbamgpu.so. Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a). "
                          "bam_parallel_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)
    L.bamgpu_scan_blocks.restype = C.c_int
    L.bamgpu_scan_blocks.argtypes = [vp, sz, C.c_int, vp, sz, C.POINTER(sz), C.POINTER(sz), u64p]
    L.bamgpu_engine_new.restype = vp
    L.bamgpu_engine_new.argtypes = [C.POINTER(Opts)]
    L.bamgpu_engine_free.argtypes = [vp]
    L.bamgpu_engine_last_error.restype = C.c_char_p
    L.bamgpu_engine_last_error.argtypes = [vp]
    L.bamgpu_device_alloc.restype = vp
    L.bamgpu_device_alloc.argtypes = [vp, sz]
    L.bamgpu_device_free.argtypes = [vp, vp]
    L.bamgpu_pinned_alloc.restype = vp
    L.bamgpu_pinned_alloc.argtypes = [sz]
    L.bamgpu_pinned_free.argtypes = [vp]
    L.bamgpu_memcpy_h2d.restype = C.c_int
    L.bamgpu_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.bamgpu_memcpy_d2h.restype = C.c_int
    L.bamgpu_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.bamgpu_engine_decode.restype = C.c_int
    L.bamgpu_engine_decode.argtypes = [vp, vp, vp, sz, C.c_uint64, C.c_int, C.c_uint32, vp, C.POINTER(Batch), u64p]
    L.bamgpu_engine_inflate.restype = C.c_int
    L.bamgpu_engine_inflate.argtypes = [vp, vp, vp, sz, C.c_uint32, vp]
    L.bamgpu_engine_data.restype = C.c_int
    L.bamgpu_engine_data.argtypes = [vp, C.POINTER(vp), u64p]
    L.bamgpu_engine_append_halo.restype = C.c_int
    L.bamgpu_engine_append_halo.argtypes = [vp, vp, sz, vp]
    L.bamgpu_engine_records.restype = C.c_int
    L.bamgpu_engine_records.argtypes = [vp, C.c_uint64, C.c_int, C.c_uint32, vp, C.POINTER(Batch), u64p]
    L.bamgpu_engine_set_n_ref.restype = None
    L.bamgpu_engine_set_n_ref.argtypes = [vp, C.c_int32]
    L.bamgpu_engine_stats.restype = C.c_int
    L.bamgpu_engine_stats.argtypes = [vp, C.POINTER(Stats)]
    L.bamgpu_engine_reset_stats.argtypes = [vp]
    L.bamgpu_reader_new.restype = vp
    L.bamgpu_reader_new.argtypes = [READ_FN, vp, sz, u64p, C.POINTER(Opts)]
    L.bamgpu_reader_from_memory.restype = vp
    L.bamgpu_reader_from_memory.argtypes = [vp, sz, sz, C.POINTER(Opts)]
    L.bamgpu_reader_free.argtypes = [vp]
    L.bamgpu_last_error.restype = C.c_char_p
    L.bamgpu_last_error.argtypes = [vp]
    L.bamgpu_read_header.restype = C.c_int
    L.bamgpu_read_header.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)]
    L.bamgpu_next_rec.restype = C.c_int
    L.bamgpu_next_rec.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.bamgpu_append_record.restype = C.c_int
    L.bamgpu_append_record.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.bamgpu_read.restype = C.c_long
    L.bamgpu_read.argtypes = [vp, vp, sz]
    L.bamgpu_next_batch.restype = C.c_int
    L.bamgpu_next_batch.argtypes = [vp, C.POINTER(Batch)]
    L.bamgpu_reader_stats.restype = C.c_int
    L.bamgpu_reader_stats.argtypes = [vp, C.POINTER(Stats)]
    L.bamgpu_parse_reference_sequences.restype = C.c_int
    L.bamgpu_parse_reference_sequences.argtypes = [vp, sz, REFSEQ_CB, vp]
    L.bamgpu_status_str.restype = C.c_char_p
    L.bamgpu_status_str.argtypes = [C.c_int]
    L.bamgpu_build_info.restype = C.c_char_p
    L.bamgpu_engine_sort.restype = C.c_int
    L.bamgpu_engine_sort.argtypes = [vp, C.c_int, C.c_uint32, vp, C.POINTER(Sorted)]
    L.bamgpu_sort_bam.restype = C.c_int
    L.bamgpu_sort_bam.argtypes = [READ_FN, vp, WRITE_FN, vp, sz, C.c_int, C.POINTER(Opts), C.POINTER(Stats), C.c_char_p, sz]
    L.bamgpu_sort_bam_from_memory.restype = C.c_int
    L.bamgpu_sort_bam_from_memory.argtypes = [vp, sz, WRITE_FN, vp, sz, C.c_int, C.POINTER(Opts), C.POINTER(Stats), C.c_char_p, sz]
    L.bamgpu_engine_digest.restype = C.c_int
    L.bamgpu_engine_digest.argtypes = [vp, C.c_uint64, C.c_uint64, vp, C.POINTER(Digest)]
    L.bamgpu_engine_sorted_digest.restype = C.c_int
    L.bamgpu_engine_sorted_digest.argtypes = [vp, vp, C.POINTER(Digest)]
    L.bamgpu_reader_current_engine.restype = vp
    L.bamgpu_reader_current_engine.argtypes = [vp]
    L.bamgpu_engine_edge_open.restype = C.c_int
    L.bamgpu_engine_edge_open.argtypes = [vp, C.POINTER(EdgeHandle)]
    L.bamgpu_engine_edge_connect.restype = C.c_int
    L.bamgpu_engine_edge_connect.argtypes = [vp, C.POINTER(EdgeHandle), C.POINTER(EdgeHandle)]
    L.bamgpu_engine_edge_exchange.restype = C.c_int
    L.bamgpu_engine_edge_exchange.argtypes = [vp, C.c_uint64, vp]
    L.bamgpu_bgzf_bound.restype = sz
    L.bamgpu_bgzf_bound.argtypes = [sz]
    L.bamgpu_engine_bgzf_store.restype = C.c_int
    L.bamgpu_engine_bgzf_store.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), C.c_int, vp]
    L.bamgpu_engine_bgzf_write.restype = C.c_int
    L.bamgpu_engine_bgzf_write.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), C.c_int, C.c_int, vp]
    L.bamgpu_engine_decode_fields.restype = C.c_int
    L.bamgpu_engine_decode_fields.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.POINTER(Fields)]
    L.bamgpu_engine_field_digest.restype = C.c_int
    L.bamgpu_engine_field_digest.argtypes = [vp, C.c_uint32, vp, C.POINTER(Digest)]
    L.bamgpu_engine_edge_close.restype = None
    L.bamgpu_engine_edge_close.argtypes = [vp]
    _lib = L
    return L
