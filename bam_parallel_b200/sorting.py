"""Host-side mirror of the reference's sort entry point over the C ABI (bamgpu_sort_bam).

    bam_tools::sorting::sort::sort_bam(mem_limit, reader, sorted_sink, tmp_dir, out_compr_level, reader_thread_num,
                                       temp_files_mode, index_file_to_create, sort_by, bam_file_size)   sort.rs:138-149
    SortBy::{Name, NameAndMatchMates, CoordinatesAndStrand}                                             sort.rs:92-96

The reference reads with its Reader, sorts chunks of records with rayon, spills them and merges; the output is the
record bodies in globally stable sorted order (sort.rs:643).  Here the Reader's CUDA path decodes the whole file into
HBM, the keys K3 extracted are radix-sorted on the device and the bodies gathered in that order: same bytes, no chunks,
no temp files.  A file that does not fit in HBM — or any file, with device_mem_limit — takes the streamed form
(sort.rs:185-308,590-652 without the chunk files): batches through the Reader, sort keys and read names on the device, bodies
in host memory, gathered there in sorted order.  Arguments that only steer the reference's external merge (mem_limit — a
HOST memory bound of 2 GB in bam_binary, main.rs:72 —, tmp_dir, temp_files_mode, out_compr_level) are accepted for signature
parity and ignored; the unreachable GBAM index mode (main.rs:74 passes None) is not built.
Where the reference panics (HI on one mate only, flag bits >= 0x1000 under coordinate sort) this raises BamGpuError.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional

import numpy as np

from . import _lib
from ._lib import BamGpuError, Opts, Stats, make_opts


class SortBy(enum.IntEnum):
    """sort.rs:92-96."""
    Name = _lib.SORT_NAME
    NameAndMatchMates = _lib.SORT_NAME_AND_MATCH_MATES
    CoordinatesAndStrand = _lib.SORT_COORDINATES_AND_STRAND


def sort_bam(mem_limit: int, reader, sorted_sink, tmp_dir=None, out_compr_level: int = 0, reader_thread_num: int = 5,
             temp_files_mode=None, index_file_to_create=None, sort_by: SortBy = SortBy.CoordinatesAndStrand,
             bam_file_size: Optional[int] = None, *, device: int = -1, flags: int = 0, output: str = "bodies",
             device_mem_limit: int = 0, batch_bytes: int = 0, devices=None) -> dict:
    """Sorts the BAM read from `reader` (bytes-like, or an object with read()/readinto()) and writes the sorted record
    bodies to `sorted_sink` (an object with write()).  Returns the stage times (bamgpu_stats)."""
    if index_file_to_create is not None:
        raise NotImplementedError("GBAM index sort (sort.rs:106-134) is unreachable from bam_binary and not built")
    L = _lib.load()
    arr = None
    if isinstance(reader, (bytes, bytearray, memoryview, np.ndarray)):
        # already in memory: bamgpu_sort_bam_from_memory reads it in place (no copy through the read callback)
        arr = np.ascontiguousarray(np.frombuffer(reader, dtype=np.uint8) if not isinstance(reader, np.ndarray) else reader)
        _read = None
    else:
        readinto = getattr(reader, "readinto", None)

        def _read(_ctx, dst, cap):
            try:
                if readinto is not None:
                    return int(readinto((C.c_uint8 * cap).from_address(C.addressof(dst.contents))) or 0)
                chunk = reader.read(cap)
                C.memmove(dst, chunk, len(chunk))
                return len(chunk)
            except Exception:
                return -1

    def _write(_ctx, src, n):
        try:
            if sorted_sink is not None:          # None: the output is produced and dropped (timing runs)
                sorted_sink.write(C.string_at(src, n))
            return 0
        except Exception:
            return -1

    wcb = _lib.WRITE_FN(_write)
    # output: "bodies" = the reference's (sort.rs:643); "records" = (u32 block_size, body)*, its spill wire format
    # (sort.rs:340-347); "bam" = a complete BAM (header + sorted records in BGZF + EOF marker): stored blocks, or with
    # out_compr_level >= 1 (sort.rs:143, unused by the reference) blocks compressed on the GPU (K6)
    fmt = {"bodies": _lib.OUT_BODIES, "records": _lib.OUT_RECORDS, "bam": _lib.OUT_BAM}[output]
    # device_mem_limit > 0 (bamgpu_opts.sort_mem_limit): the streamed form, with batches of about device_mem_limit / 8 compressed bytes
    # devices (bamgpu_opts.devices[]): without a memory limit the sample sort over those GPUs (contiguous block ranges decoded side by
    # side, buckets exchanged with peer copies, each GPU sorts one bucket); the streamed form decodes its batches on all of them and
    # collects the keys on the first one used
    opts, st = make_opts(device, flags, sort_output=fmt, sort_mem_limit=device_mem_limit, batch_bytes=batch_bytes, devices=devices,
                          out_compr_level=out_compr_level), Stats()
    err = C.create_string_buffer(512)
    if arr is not None:
        rc = L.bamgpu_sort_bam_from_memory(arr.ctypes.data, arr.size, wcb, None, reader_thread_num, int(sort_by), C.byref(opts), C.byref(st),
                                           err, len(err))
    else:
        rcb = _lib.READ_FN(_read)
        rc = L.bamgpu_sort_bam(rcb, None, wcb, None, reader_thread_num, int(sort_by), C.byref(opts), C.byref(st), err, len(err))
    if rc < 0:
        raise BamGpuError(rc, err.value.decode(errors="replace"))
    return st.asdict()
