"""CPU-only: the C-ABI library loads without a GPU, exports every symbol include/bamgpu.h declares, and its
host-side pieces (BGZF scan, reference-sequence parse) match the oracle.  No kernel is launched here."""
import os
import re
import struct

import numpy as np
import pytest

import bam_parallel_b200 as bp
from bam_parallel_b200 import _lib, synth
from oracle import oracle as orc
from tests import bamcraft as bc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_header_symbols():
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "bamgpu.h")).read()
    declared = set(re.findall(r"\b(bamgpu_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"bamgpu_read_fn", "bamgpu_refseq_cb"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(L, name) is not None
    assert L.bamgpu_build_info().startswith(b"sm_100a")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bp.BamGpuError):
        bp.Engine()
    with pytest.raises(bp.BamGpuError):
        bp.Reader.new(bc.bgzf(bc.bam_header()), 4)
    import io
    with pytest.raises(bp.BamGpuError):       # the sort entry point has no CPU path either
        bp.sort_bam(0, bc.bgzf(bc.bam_header()), io.BytesIO(), sort_by=bp.SortBy.Name)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bam_parallel_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("no cpu fallback", ""), f


@pytest.mark.parametrize("shape,n,payload", [("pe150", 5000, 0xFF00), ("long10k", 50, 4096), ("pe150", 200, 300)])
def test_scan_blocks_matches_oracle(oracle, shape, n, payload):
    data = synth.generate(shape, n, payload=payload, threads=2).tobytes()
    blocks, nb, consumed, total, rc = bp.scan_blocks(data)
    ob = oracle.scan_blocks(data)
    u, sizes = oracle.inflate_all(data)
    assert rc == 1 and nb == len(ob) and consumed == len(data) and total == len(u)
    assert [blocks[i].in_off - 18 for i in range(nb)] == ob["coff"].tolist()
    assert [blocks[i].in_len + 26 for i in range(nb)] == ob["bsize"].tolist()
    assert [blocks[i].isize for i in range(nb)] == sizes.tolist()
    assert [blocks[i].out_off for i in range(nb)] == np.concatenate([[0], np.cumsum(sizes)[:-1]]).tolist()
    for i in (0, nb // 2, nb - 1):
        o = int(blocks[i].out_off)
        assert blocks[i].crc32 == oracle.crc32(u[o:o + blocks[i].isize])


def test_scan_blocks_quirks(oracle):
    stream = bc.bam_header() + bc.bam_record()
    good = bc.bgzf_block(stream)
    evil = bytearray(bc.bgzf_block(b"x" * 10)); evil[16:18] = b"\xff\xff"
    _, nb, consumed, _, rc = bp.scan_blocks(good + bytes(evil) + good)
    assert (nb, consumed, rc) == (1, len(good), 1)                  # BSIZE wrap => silent EOF (util.rs:65)
    _, nb, consumed, _, rc = bp.scan_blocks(good + b"\x1f\x8b\x08")
    assert (nb, rc) == (1, 1)                                        # short trailing header => EOF (util.rs:58-60)
    _, nb, consumed, _, rc = bp.scan_blocks(good + good[:40], final=False)
    assert (nb, consumed, rc) == (1, len(good), 0)                   # streaming: partial block left for the next chunk
    _, nb, consumed, _, rc = bp.scan_blocks(good + good[:40], final=True)
    assert rc == -11                                                 # util.rs:42 read_exact failure
    tiny = bytearray(bc.BGZF_EOF); tiny[16:18] = struct.pack("<H", 19)
    assert bp.scan_blocks(bytes(tiny))[4] == -10                     # util.rs:29-38
    noisy = bytearray(good); noisy[0:4] = b"ABCD"
    assert bp.scan_blocks(bytes(noisy))[1] == 1                      # no magic check (util.rs:55-66)
    assert bp.scan_blocks(b"")[1:] == (0, 0, 0, 1)


def test_parse_reference_sequences(oracle):
    hdr = bc.bam_header(refs=(("chr1", 1000), ("chrUn_gl000220", 161802), ("", 5)))
    u = np.frombuffer(hdr, np.uint8)
    hb, off, consumed = oracle.read_header(u)
    assert bp.parse_reference_sequences(hb[off:]) == orc.py_parse_reference_sequences(hb[off:]) == \
        [("chr1", 1000), ("chrUn_gl000220", 161802), ("", 5)]
    bad = struct.pack("<I", 1) + struct.pack("<I", 4) + b"abcd" + struct.pack("<I", 7)   # no NUL terminator
    with pytest.raises(bp.BamGpuError):
        bp.parse_reference_sequences(bad)
    with pytest.raises(bp.BamGpuError):
        bp.parse_reference_sequences(hb[off:-2])


def test_opts_layout_is_the_same_in_every_binding():
    """bamgpu_opts crosses the ABI by pointer: the ctypes mirror and the Rust source must list the header's fields in order."""
    import re
    from bam_parallel_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "bamgpu.h")).read()
    body = re.search(r"typedef struct bamgpu_opts \{(.*?)\} bamgpu_opts;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = [re.sub(r"\[.*\]", "", d.split()[-1]) for d in body.split(";") if d.strip()]
    assert c_fields == [n for n, _ in _lib.Opts._fields_]
    rs = open(os.path.join(root, "rust", "bamgpu-sys", "src", "lib.rs")).read()
    rbody = re.search(r"pub struct bamgpu_opts \{(.*?)\n\}", rs, re.S).group(1)
    assert re.findall(r"pub (\w+):", rbody) == c_fields
    import ctypes as C
    assert C.sizeof(_lib.Opts) == 4 + 4 + 8 + 4 + 4 + 4 + 4 * 16 + 4 + 8 + 4 + 4   # no implicit padding
