"""ctypes driver for tests/host_emul.cpp (CPU emulation of the K1 per-lane logic). TEST ONLY."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SRC = os.path.join(ROOT, "tests", "host_emul.cpp")
_CORE = os.path.join(ROOT, "bam_parallel_b200", "csrc", "inflate_core.cuh")


def _build(asan: bool, defines=()) -> str:
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    tag = "".join("_" + d.replace("=", "") for d in defines)
    out = os.path.join(ROOT, "build", ("libhostemul_asan" if asan else "libhostemul") + tag + ".so")
    newest = max(os.path.getmtime(_SRC), os.path.getmtime(_CORE))
    if not os.path.exists(out) or os.path.getmtime(out) < newest:
        flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"] if asan else ["-O2"]
        subprocess.check_call(["g++", *flags, *["-D" + d for d in defines], "-fPIC", "-shared", "-std=c++17", "-o", out, _SRC])
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build(False))
        _lib.host_emul_inflate.restype = C.c_int
        _lib.host_emul_inflate.argtypes = [C.c_void_p] * 7 + [C.c_uint32, C.c_int]
    return _lib


def py_scan(data: bytes):
    """Block table the way bamgpu_scan_blocks builds it (payload offset/len, ISIZE scan, CRC)."""
    in_off, in_len, out_off, isize, crc = [], [], [], [], []
    pos, o, n = 0, 0, len(data)
    while n - pos >= 18:
        bs = (struct.unpack_from("<H", data, pos + 16)[0] + 1) & 0xFFFF
        if bs == 0:
            break
        assert bs >= 26 and n - pos >= bs
        c, isz = struct.unpack_from("<II", data, pos + bs - 8)
        in_off.append(pos + 18); in_len.append(bs - 26); out_off.append(o); isize.append(isz); crc.append(c)
        o += isz
        pos += bs
    return (np.array(in_off, np.uint64), np.array(in_len, np.uint32), np.array(out_off, np.uint64),
            np.array(isize, np.uint32), np.array(crc, np.uint32), o)


def variant(defines):
    """The emulation built with extra -D switches (kernel variants kept behind macros)."""
    L = C.CDLL(_build(False, tuple(defines)))
    L.host_emul_inflate.restype = C.c_int
    L.host_emul_inflate.argtypes = [C.c_void_p] * 7 + [C.c_uint32, C.c_int]
    return L


def inflate(data: bytes, lanes: int = 32, table=None, library=None):
    in_off, in_len, out_off, isize, crc, total = table if table is not None else py_scan(data)
    comp = np.zeros(len(data) + 64, np.uint8)           # K1 prefetches up to 48 bytes past a payload
    comp[: len(data)] = np.frombuffer(data, np.uint8)
    outbuf = np.full(total + 64, 0xEE, np.uint8)       # 32 B lead pad (>=16), tail pad
    lead = 32 + (-outbuf.ctypes.data) % 4              # keep out[0] 4-byte aligned
    status = np.full(max(len(in_off), 1), 0xFFFFFFFF, np.uint32)
    (library or lib()).host_emul_inflate(comp.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, out_off.ctypes.data,
                            isize.ctypes.data, outbuf.ctypes.data + lead, status.ctypes.data, len(in_off), lanes)
    return outbuf[lead: lead + total], status[: len(in_off)], (outbuf[:lead], outbuf[lead + total:])
