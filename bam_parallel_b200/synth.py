"""Seeded synthetic BAM inputs (ctypes binding of tools/bamgen.cpp).

There is no network on the build or GPU boxes, so the BASELINE.json workloads are
manufactured by this generator.  It only makes *inputs*; it is not on the decode path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(_ROOT, "bam_parallel_b200", "lib", "libbamgen.so")

SHAPES = {"se150": 0, "pe150": 1, "longname": 2, "long10k": 3}
DEFAULT_SEED = 0xBA5EBA11


class _Params(C.Structure):
    _fields_ = [("shape", C.c_int32), ("level", C.c_int32), ("n_records", C.c_uint64), ("seed", C.c_uint64),
                ("payload", C.c_uint32), ("threads", C.c_int32), ("no_eof", C.c_int32), ("header_block", C.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-C", _ROOT, "bam_parallel_b200/lib/libbamgen.so"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB)
        _lib.bamgen_generate.restype = C.c_int
        _lib.bamgen_generate.argtypes = [C.POINTER(_Params), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib.bamgen_free.argtypes = [C.c_void_p]
    return _lib


class SynthBam:
    """A generated BAM held in native memory; ``.array`` is a zero-copy uint8 view."""

    def __init__(self, ptr, nbytes, ustream_len, header_len, params):
        self._ptr = ptr
        self.nbytes = nbytes
        self.ustream_len = ustream_len
        self.header_len = header_len
        self.params = params
        self.array = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def tobytes(self) -> bytes:
        return self.array.tobytes()

    def close(self):
        if self._ptr:
            self.array = None
            _load().bamgen_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate(shape: str, n_records: int, level: int = 6, payload: int = 0xFF00, seed: int = DEFAULT_SEED,
             threads: int = 0, no_eof: bool = False, header_block: bool = False) -> SynthBam:
    lib = _load()
    p = _Params(SHAPES[shape], level, n_records, seed, payload, threads, int(no_eof), int(header_block))
    out, n, u, h = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib.bamgen_generate(C.byref(p), C.byref(out), C.byref(n), C.byref(u), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"bamgen_generate failed rc={rc}")
    return SynthBam(out.value, n.value, u.value, h.value,
                    dict(shape=shape, n_records=n_records, level=level, payload=payload, seed=seed))
