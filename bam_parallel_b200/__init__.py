"""bam_parallel_b200 — B200-native BGZF inflate + BAM record decode behind the bam_tools::Reader API.

Only the hot path of NickRoz1/bam_parallel is here: CUDA kernels + C ABI in
``csrc/`` (built into ``lib/libbamgpu.so``), and this host-side mirror of the reference interface.
"""
from ._lib import DEFER_CRC  # noqa: F401
from ._lib import BamGpuError, COPY_COLUMNS, COPY_DATA, COPY_OFFSETS, FIELD_CIGAR, FIELD_QUAL, FIELD_SEQ, NO_CRC, NO_RECORDS, SPECULATIVE_START, WANT_HI, load  # noqa: F401
from .reader import Engine, Reader, RecordBatch, Records, parse_reference_sequences, scan_blocks  # noqa: F401
from .sorting import SortBy, sort_bam  # noqa: F401

__all__ = ["Reader", "Records", "RecordBatch", "Engine", "parse_reference_sequences", "scan_blocks", "BamGpuError",
           "COPY_DATA", "COPY_COLUMNS", "COPY_OFFSETS", "WANT_HI", "NO_CRC", "NO_RECORDS", "SPECULATIVE_START", "load", "SortBy", "sort_bam", "FIELD_CIGAR", "FIELD_SEQ", "FIELD_QUAL"]
