"""CPU-only: the oracle's full-size helpers (several threads) are the SAME restatement as its serial functions, and the
DIGEST SPEC has one meaning in its two CPU statements (C: orc_digest, pure Python: py_digest).  The CUDA statement
(csrc/digest.cu) is compared with these on the GPU box (tests/test_gpu_fullsize.py, tests/test_gpu_multi.py)."""
import numpy as np
import pytest

from bam_parallel_b200 import synth
from oracle import oracle as orc


@pytest.fixture(scope="module")
def decoded(oracle):
    out = {}
    for shape, n, payload in (("pe150", 6000, 0xFF00), ("longname", 3000, 0xFF00), ("long10k", 60, 4096)):
        data = synth.generate(shape, n, payload=payload, threads=2).tobytes()
        u, _ = oracle.inflate_all(data)
        _, _, start = oracle.read_header(u)
        off, ln = oracle.walk_records(u, start)
        cols, bf, bt = oracle.extract(u, off, ln, True)
        out[shape] = (data, u, off, ln, cols)
    return out


def test_threaded_helpers_equal_the_serial_restatement(oracle, decoded):
    for shape, (data, u, off, ln, cols) in decoded.items():
        um = oracle.inflate_all_mt(np.frombuffer(data, np.uint8), 3)
        assert np.array_equal(um, u)
        c3, bf, bt = oracle.extract_mt(u, off, ln, True, 3)
        for k in cols:
            assert np.array_equal(c3[k], cols[k]), (shape, k)
        for mode in (0, 1, 2):
            assert np.array_equal(oracle.sort_perm_mt(u, off, cols, mode, 4), oracle.sort_perm(u, off, cols, mode)), (shape, mode)
        oracle.release(um)


def test_sort_perm_mt_large_enough_to_split_into_runs(oracle):
    data = synth.generate("pe150", 140000, threads=4).tobytes()      # > 65536 records: four sorted runs + two merge levels
    u, _ = oracle.inflate_all(data)
    _, _, start = oracle.read_header(u)
    off, ln = oracle.walk_records(u, start)
    cols, _, _ = oracle.extract(u, off, ln, True)
    for mode in (0, 1, 2):
        assert np.array_equal(oracle.sort_perm_mt(u, off, cols, mode, 4), oracle.sort_perm(u, off, cols, mode)), mode
    a, b = oracle.sort_perm(u, off, cols, 1), oracle.sort_perm(u, off, cols, 2)
    assert not np.array_equal(a, b), "mates are emitted in shuffled order: -n and -n -M must differ"


def test_digest_spec_c_equals_python(oracle, decoded):
    for shape, (data, u, off, ln, cols) in decoded.items():
        n = min(off.size, 400)
        sub = {k: v[:n] for k, v in cols.items()}
        d_c = oracle.digest(u, off[:n].copy(), ln[:n].copy(), {k: np.ascontiguousarray(v) for k, v in sub.items()}, 2, index_base=7, stream_base=1000)
        d_p = orc.py_digest(u.tobytes(), off[:n], ln[:n], sub, index_base=7, stream_base=1000)
        assert d_c == d_p, shape
        perm = oracle.sort_perm(u, off, cols, 2)
        s_c = oracle.sorted_digest(u, off, ln, perm, 2)
        if off.size <= 3000:
            assert s_c == orc.py_sorted_digest(u.tobytes(), off, ln, perm), shape


def test_digest_is_additive_over_shards_and_sees_every_byte(oracle, decoded):
    data, u, off, ln, cols = decoded["pe150"]
    whole = oracle.digest(u, off, ln, cols, 3)
    k = off.size // 3
    parts = [(0, k), (k, 2 * k + 5), (2 * k + 5, off.size)]
    tot = {"n_records": 0, "body_bytes": 0, "fields": 0, "bodies": 0}
    for a, b in parts:       # a shard sees offsets relative to its own start: index_base / stream_base make them global
        base = int(off[a]) - 4
        sub = {kk: np.ascontiguousarray(v[a:b]) for kk, v in cols.items()}
        d = oracle.digest(u[base:], np.ascontiguousarray(off[a:b] - np.uint64(base)), np.ascontiguousarray(ln[a:b]), sub, 2,
                          index_base=a, stream_base=base)
        for kk in tot:
            tot[kk] = (tot[kk] + d[kk]) & 0xFFFFFFFFFFFFFFFF
    assert tot == whole
    u2 = u.copy()
    u2[int(off[17]) + 40] ^= 1                                   # one bit of one body
    assert oracle.digest(u2, off, ln, cols, 3)["bodies"] != whole["bodies"]
    c2 = {kk: v.copy() for kk, v in cols.items()}
    c2["tlen"][99] += 1                                          # one field of one record
    assert oracle.digest(u, off, ln, c2, 3)["fields"] != whole["fields"]
    perm = oracle.sort_perm(u, off, cols, 0)
    p2 = perm.copy(); p2[[3, 4]] = p2[[4, 3]]                    # two neighbours swapped in the output order
    assert oracle.sorted_digest(u, off, ln, p2, 2) != oracle.sorted_digest(u, off, ln, perm, 2)


@pytest.mark.parametrize("threads", [3, 5, 8, 20])
def test_cpu_reader_port_counts_every_record(oracle, decoded, threads):
    data, u, off, ln, cols = decoded["pe150"]
    r = oracle.cpu_reader_run(np.frombuffer(data, np.uint8), threads)
    assert r["records"] == off.size and r["ubytes"] == u.size
    r = oracle.cpu_reader_run(np.frombuffer(data, np.uint8), threads, 1000)     # bounded sample: stops early, threads shut down
    assert r["records"] == 1000
