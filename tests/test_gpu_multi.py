"""The Reader's job pipeline and the multi-GPU paths behind the C ABI, against the oracle.

* bamgpu_opts.devices: batches go to the listed GPUs round robin; the partial record at a batch edge is copied in front
  of the next batch on whichever GPU decodes it (cudaMemcpyPeerAsync between different GPUs).  Listing the same GPU
  several times exercises the identical code on a one-GPU box (the copy is then device-to-device).
* bamgpu_engine_edge_*: block-range shards whose heads travel to the left-hand neighbour's inbox in peer memory.
Every result is compared with the oracle's walk of the same file: bodies (md5, main.rs:83-99), record boundaries,
columns and the DIGEST SPEC sums.
"""
import hashlib
import io

import numpy as np
import pytest

import bam_parallel_b200 as bp
from bam_parallel_b200 import _lib, synth
from bam_parallel_b200.sharding import check_edges, shard_ranges
from oracle import oracle as orc
from tests import bamcraft as bc

pytestmark = pytest.mark.gpu

COLS = [n for n in orc.COLUMN_DTYPES]


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _oracle(oracle, data):
    u, _ = oracle.inflate_all(data)
    hdr, off_refs, start = oracle.read_header(u)
    off, ln = oracle.walk_records(u, start)
    cols, bf, bt = oracle.extract(u, off, ln, True)
    return u, start, off, ln, cols


def _md5_bodies(u, off, ln):
    ub = u.tobytes()
    return hashlib.md5(b"".join(ub[o:o + l] for o, l in zip(off.tolist(), ln.tolist()))).hexdigest()


def _reader_all(data, **kw):
    """Everything a Reader hands out through next_batch: bodies md5, concatenated columns, summed digest, #batches."""
    h = hashlib.md5()
    got = {k: [] for k in COLS + ["strand", "rec_len"]}
    stream_off = []
    tot = {"n_records": 0, "body_bytes": 0, "fields": 0, "bodies": 0}
    nb = 0
    with bp.Reader.new(data, 4, flags=bp.COPY_COLUMNS | bp.COPY_DATA | bp.WANT_HI, **kw) as r:
        r.read_header()
        while True:
            b = r.next_batch()
            if b is None:
                break
            assert b.first_record_index == tot["n_records"]
            d = r.batch_digest(b)
            for k in tot:
                tot[k] = (tot[k] + d[k]) & 0xFFFFFFFFFFFFFFFF
            db = b.data.tobytes()
            for o, l in zip(b.columns["rec_off"].tolist(), b.columns["rec_len"].tolist()):
                h.update(db[o:o + l])
            for k in got:
                got[k].append(b.columns[k].copy())
            stream_off.append(b.columns["rec_off"].astype(np.int64) - b.data_start + b.stream_offset)
            nb += 1
        st = r.stats()
    cat = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in got.items()}
    return h.hexdigest(), cat, (np.concatenate(stream_off) if stream_off else np.zeros(0, np.int64)), tot, nb, st


def _assert_reader(oracle, data, **kw):
    u, start, off, ln, cols = _oracle(oracle, data)
    md5, cat, soff, tot, nb, st = _reader_all(data, **kw)
    assert md5 == _md5_bodies(u, off, ln)
    assert np.array_equal(soff, off.astype(np.int64)), "stream offsets of the bodies differ"
    assert np.array_equal(cat["rec_len"], ln)
    for k in COLS:
        if k == "hi_present":
            assert np.array_equal(cat[k] == 1, cols[k] == 1), k
        else:
            assert np.array_equal(cat[k], cols[k]), k
    assert tot == oracle.digest(u, off, ln, cols, 4)
    return nb, st


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], None])
@pytest.mark.parametrize("shape,n,payload,batch_bytes", [("pe150", 60000, 0xFF00, 1 << 20), ("long10k", 600, 4096, 200000),
                                                         ("longname", 12000, 0xFF00, 300000), ("pe150", 4000, 300, 70000)])
def test_reader_jobs_round_robin_one_gpu(oracle, devices, shape, n, payload, batch_bytes):
    data = synth.generate(shape, n, payload=payload).tobytes()
    nb, st = _assert_reader(oracle, data, devices=devices, batch_bytes=batch_bytes, jobs_per_device=2 if devices else 0)
    assert nb > 3


def test_reader_multi_gpu_peer_carry(oracle):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    devs = list(range(min(n, 8)))
    for shape, cnt, payload in (("pe150", 400000, 0xFF00), ("long10k", 3000, 4096), ("longname", 60000, 0xFF00)):
        data = synth.generate(shape, cnt, payload=payload).tobytes()
        bb = max(len(data) // (3 * len(devs)), 70000)          # at least three batches per GPU
        nb, st = _assert_reader(oracle, data, devices=devs, batch_bytes=bb)
        assert nb >= 2 * len(devs)
        assert st["p2p_bytes"] > 0, "no bytes moved between GPUs"


def test_record_longer_than_the_gap_between_batches(oracle):
    """A record larger than the 1 MiB kept free in front of a batch's blocks (round 1 refused it at shard edges):
    the blocks move behind a larger gap and the decode goes on."""
    hdr = bc.bam_header()
    big = bc.bam_record(name=b"big", seq_len=1_400_000, cigar=((1_400_000, 0),))     # ~2.1 MB
    stream = hdr + bc.bam_record(name=b"a") * 50 + big + bc.bam_record(name=b"b") * 300 + big + bc.bam_record(name=b"c") * 7
    data = bc.bgzf(stream, level=1)
    for kw in ({"batch_bytes": 70000}, {"batch_bytes": 70000, "devices": [0, 0], "jobs_per_device": 2}):
        _assert_reader(oracle, data, **kw)


def test_switching_from_batches_to_records_keeps_every_record(oracle):
    """next_batch() (no host mirrors asked for) followed by next_rec(): the batches staged ahead get the mirrors they lack."""
    data = synth.generate("pe150", 40000).tobytes()
    u, start, off, ln, cols = _oracle(oracle, data)
    with bp.Reader.new(data, 4, batch_bytes=1 << 20) as r:
        r.read_header()
        b = r.next_batch()
        n0 = b.n_records
        assert 0 < n0 < off.size
        h = hashlib.md5()
        ub = u.tobytes()
        for o, l in zip(off[n0:].tolist(), ln[n0:].tolist()):
            h.update(ub[o:o + l])
        g = hashlib.md5()
        recs = r.records()
        cnt = 0
        while True:
            rec = recs.next_rec()
            if rec is None:
                break
            g.update(rec)
            cnt += 1
        assert cnt == off.size - n0 and g.hexdigest() == h.hexdigest()


def test_read_after_read_header_continues_the_stream(oracle):
    """impl Read for Reader (reader.rs:114-152) after read_header(): the byte stream goes on behind the header."""
    data = synth.generate("pe150", 20000).tobytes()
    u, _ = oracle.inflate_all(data)
    _, _, start = oracle.read_header(u)
    for bb in (0, 300000):
        with bp.Reader.new(data, 4, batch_bytes=bb) as r:
            r.read_header()
            got = r.read(777) + r.read()
            assert got == u.tobytes()[start:]
            with pytest.raises(bp.BamGpuError):
                r.records().next_rec()


def test_good_records_come_out_before_the_error(oracle):
    """reader.rs:63-81 hands out every complete record and only then panics on the short one; a truncated BGZF block kills the
    reference's reading thread after the blocks in front of it were delivered."""
    recs = [bc.bam_record(name=b"r%d" % i, pos=i) for i in range(2000)]
    good = bc.bam_header() + b"".join(recs)
    # 1. short last record
    data = bc.bgzf(good + bc.bam_record()[:-5], payload=4000)
    for bb in (0, 70000):
        with bp.Reader.new(data, 4, batch_bytes=bb) as r:
            r.read_header()
            it = r.records()
            n = 0
            with pytest.raises(bp.BamGpuError) as e:
                while it.next_rec() is not None:
                    n += 1
            assert n == 2000 and e.value.code == -22
    # 2. truncated final BGZF block
    whole = bc.bgzf(good, payload=4000, eof=False)
    blocks, nb, consumed, total, rc = bp.scan_blocks(whole)
    cut = whole[:-9]
    u, _ = oracle.inflate_all(whole[:blocks[nb - 1].in_off - 18] + bc.BGZF_EOF)
    _, _, start = oracle.read_header(u)
    want = 0
    p = start
    while len(u) - p >= 4:     # complete records inside the blocks in front of the truncated one
        bs = int.from_bytes(u[p:p + 4].tobytes(), "little")
        if len(u) - (p + 4) < bs:
            break
        want += 1
        p += 4 + bs
    for bb in (0, 70000):
        with bp.Reader.new(cut, 4, batch_bytes=bb) as r:
            r.read_header()
            it = r.records()
            n = 0
            with pytest.raises(bp.BamGpuError) as e:
                while it.next_rec() is not None:
                    n += 1
            assert n == want and e.value.code == -11


def test_page_locked_source_is_read_in_place(oracle):
    """A source buffer the caller page-locked (bamgpu_pinned_alloc) is copied to the GPUs straight from where it is."""
    import ctypes as C
    bam = synth.generate("pe150", 50000)
    L = bp.load()
    p = L.bamgpu_pinned_alloc(bam.nbytes)
    assert p
    try:
        C.memmove(p, bam.array.ctypes.data, bam.nbytes)
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(bam.nbytes,))
        _assert_reader(oracle, arr, batch_bytes=1 << 20)
        _assert_reader(oracle, arr, batch_bytes=1 << 20, devices=[0, 0])
    finally:
        L.bamgpu_pinned_free(p)


def test_sort_twice_does_not_leak_device_memory():
    """bamgpu_engine_free gives the sort's scratch, permutation, offsets and gathered bodies back (round 1 leaked them)."""
    import torch
    data = synth.generate("pe150", 300000).tobytes()
    free0 = None
    for it in range(4):
        out = io.BytesIO()
        bp.sort_bam(0, data, out, sort_by=bp.SortBy.CoordinatesAndStrand)
        torch.cuda.synchronize()
        free, _ = torch.cuda.mem_get_info()
        if it == 1:
            free0 = free
    assert free0 - free < (64 << 20), f"device memory shrank by {(free0 - free) >> 20} MiB over two more sorts"


@pytest.mark.parametrize("shape,n,payload,n_shards", [("pe150", 30000, 0xFF00, 3), ("long10k", 400, 4096, 4), ("pe150", 3000, 700, 5)])
def test_edge_exchange_in_peer_memory(oracle, shape, n, payload, n_shards):
    """bamgpu_engine_edge_open / _connect / _exchange: shard g+1 writes its head into shard g's inbox and publishes the step;
    shard g waits on the device, appends the halo, acknowledges.  Engines sit on as many GPUs as there are (round robin)."""
    data = synth.generate(shape, n, level=6, payload=payload).tobytes()
    u, start, off, ln, cols = _oracle(oracle, data)
    blocks, nb, consumed, total, rc = bp.scan_blocks(data)
    bv = np.frombuffer(blocks, dtype=np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("isize", "<u4"),
                                               ("crc32", "<u4"), ("reserved", "<u4")]))[:nb]
    ranges = shard_ranges(bv["in_off"], n_shards)
    ng = max(_n_gpus(), 1)
    engines, handles, shard = [], [], []
    try:
        for g, (b0, b1) in enumerate(ranges):
            e = bp.Engine(device=g % ng)
            engines.append(e)
            handles.append(e.edge_open())
        for g, e in enumerate(engines):
            e.edge_connect(handles[g - 1] if g > 0 else None, handles[g + 1] if g < n_shards - 1 else None)
        for g, (b0, b1) in enumerate(ranges):
            e = engines[g]
            mb = (_lib.Block * max(b1 - b0, 1))()
            mv = np.frombuffer(mb, dtype=bv.dtype)[: b1 - b0]
            mv[:] = bv[b0:b1]
            c0 = int(bv["in_off"][b0]) - 18 if b1 > b0 else 0
            c1 = int(bv["in_off"][b1 - 1]) + int(bv["in_len"][b1 - 1]) + 8 if b1 > b0 else 0
            u0 = int(bv["out_off"][b0]) if b1 > b0 else total
            mv["in_off"] -= np.uint64(c0)
            mv["out_off"] -= np.uint64(u0)
            shard.append((e.upload(data[c0:c1]), mb, b1 - b0, u0, int(mv["isize"].sum())))
        for step in (1, 2, 3):          # three passes: the two inbox buffers and the acknowledgements are reused
            for g, e in enumerate(engines):
                # step 2: BAMGPU_DEFER_CRC - K2 runs beside the exchange and the record kernels, records() delivers its verdict
                e.inflate(shard[g][0], shard[g][1], shard[g][2], flags=bp.DEFER_CRC if step == 2 else 0)
            for g in reversed(range(n_shards)):       # one host thread: a shard's right-hand neighbour must have sent first
                engines[g].edge_exchange(step)
            offs, lens, exits, starts, counts = [], [], [], [], []
            tot = {"n_records": 0, "body_bytes": 0, "fields": 0, "bodies": 0}
            for g, e in enumerate(engines):
                batch, tail, rc = e.records(start if g == 0 else 0, g == n_shards - 1,
                                            (bp.SPECULATIVE_START if g else 0) | bp.COPY_COLUMNS | bp.WANT_HI)
                offs.append(batch.columns["rec_off"].astype(np.uint64) + np.uint64(shard[g][3]))
                lens.append(batch.columns["rec_len"].copy())
                exits.append(tail - shard[g][4])
                starts.append(batch.chain_start)
                counts.append(batch.n_records)
                d = e.digest(index_base=sum(counts[:-1]), stream_base=shard[g][3])
                for k in tot:
                    tot[k] = (tot[k] + d[k]) & 0xFFFFFFFFFFFFFFFF
            check_edges(exits, starts, counts, off.size)
            assert np.array_equal(np.concatenate(offs), off) and np.array_equal(np.concatenate(lens), ln)
            assert tot == oracle.digest(u, off, ln, cols, 4)
        assert sum(e.stats()["p2p_bytes"] for e in engines) > 0
    finally:
        for e in engines:
            e.close()


@pytest.mark.parametrize("sort_by", ["CoordinatesAndStrand", "Name", "NameAndMatchMates"])
def test_streamed_sort_over_several_gpus(oracle, sort_by):
    """bamgpu_sort_bam with sort_mem_limit and a device list: the batches are decoded on every listed GPU, the sort keys
    (and packed read names) collect on the first one — peer copies between different GPUs —, one stable key sort there, the
    bodies gathered on the host.  Output = the oracle's (sort.rs:138-178, comparators.rs:61-147)."""
    from bam_parallel_b200.sorting import SortBy, sort_bam
    sb = SortBy[sort_by]
    ng = _n_gpus()
    devices = list(range(ng)) if ng > 1 else [0, 0, 0]
    data = synth.generate("pe150", 80000, level=6).tobytes()
    u, start, off, ln, cols = _oracle(oracle, data)
    perm = oracle.sort_perm(u, off, cols, {"Name": 1, "NameAndMatchMates": 2, "CoordinatesAndStrand": 0}[sort_by])
    ub = u.tobytes()
    want = b"".join(ub[int(off[i]):int(off[i]) + int(ln[i])] for i in perm.tolist())
    sink = io.BytesIO()
    st = sort_bam(0, data, sink, sort_by=sb, device_mem_limit=8 << 20, devices=devices)
    assert st["batches"] > 3
    assert hashlib.md5(sink.getvalue()).hexdigest() == hashlib.md5(want).hexdigest()
    if ng > 1:
        assert st["p2p_bytes"] > 0            # keys really crossed GPUs


def test_deferred_crc_verdict_comes_with_the_records(oracle):
    """BAMGPU_DEFER_CRC: bamgpu_engine_inflate returns at once, a damaged block is reported by bamgpu_engine_records (nothing is
    handed out), and the next inflate on the engine starts clean."""
    data = bytearray(synth.generate("pe150", 20000, level=6).tobytes())
    u, start, off, ln, cols = _oracle(oracle, bytes(data))
    blocks, nb, _, _, _ = bp.scan_blocks(bytes(data))
    e = bp.Engine()
    try:
        d_good = e.upload(bytes(data))
        bad = bytearray(data)
        bad[int(blocks[nb // 2].in_off + blocks[nb // 2].in_len)] ^= 0x21          # first CRC32 byte of a block in the middle
        d_bad = e.upload(bytes(bad))
        bad_blocks, _, _, _, _ = bp.scan_blocks(bytes(bad))                        # the footers travel in the block table
        e.inflate(d_bad, bad_blocks, nb, flags=bp.DEFER_CRC)                       # no error yet
        with pytest.raises(bp.BamGpuError) as err:
            e.records(start, True, 0)
        assert err.value.code == -14
        e.inflate(d_good, blocks, nb, flags=bp.DEFER_CRC)
        batch, tail, rc = e.records(start, True, bp.COPY_COLUMNS)
        assert batch.n_records == off.size and np.array_equal(batch.columns["rec_off"], off)
        e.inflate(d_bad, bad_blocks, nb, flags=bp.DEFER_CRC)                       # verdict never collected by a records call ...
        with pytest.raises(bp.BamGpuError) as err:
            e.inflate(d_good, blocks, nb)                                          # ... so the next inflate reports it
        assert err.value.code == -14
        e.inflate(d_good, blocks, nb)
        batch, tail, rc = e.records(start, True, 0)
        assert batch.n_records == off.size
    finally:
        e.close()


def _n_sort_devices(k):
    ng = _n_gpus()
    return list(range(ng)) if ng > 1 else [0] * k


@pytest.mark.parametrize("sort_by", ["CoordinatesAndStrand", "Name", "NameAndMatchMates"])
@pytest.mark.parametrize("shape,n,payload,k", [("pe150", 90000, 0xFF00, 3), ("long10k", 500, 4096, 4), ("longname", 9000, 0xFF00, 2),
                                               ("pe150", 4000, 700, 5)])
def test_sample_sort_over_several_gpus(oracle, shape, n, payload, k, sort_by):
    """bamgpu_sort_bam with a device list and no memory limit ("multi-GPU => sample sort"): contiguous block ranges are
    decoded side by side, cut into buckets by sampled splitters, exchanged with peer copies and sorted bucket by bucket.  The three
    output formats must be byte-identical to the single-GPU sort's and the bodies to the oracle's (sort.rs:138-178, :643,
    comparators.rs:61-147)."""
    from bam_parallel_b200.sorting import SortBy, sort_bam
    sb = SortBy[sort_by]
    devices = _n_sort_devices(k)
    data = synth.generate(shape, n, level=6, payload=payload).tobytes()
    u, start, off, ln, cols = _oracle(oracle, data)
    perm = oracle.sort_perm(u, off, cols, {"Name": 1, "NameAndMatchMates": 2, "CoordinatesAndStrand": 0}[sort_by])
    ub = u.tobytes()
    want = b"".join(ub[int(off[i]):int(off[i]) + int(ln[i])] for i in perm.tolist())
    sink = io.BytesIO()
    st = sort_bam(0, data, sink, sort_by=sb, devices=devices)
    assert hashlib.md5(sink.getvalue()).hexdigest() == hashlib.md5(want).hexdigest()
    assert st["records"] >= n and st["batches"] >= len(devices)          # every range and every bucket went through K3
    if len(set(devices)) > 1:
        # the buckets really crossed GPUs (the generator's names grow with the input position, so the name orders move little)
        assert st["p2p_bytes"] > (len(want) // 4 if sort_by == "CoordinatesAndStrand" else 0)
    for fmt in ("records", "bam"):
        one, many = io.BytesIO(), io.BytesIO()
        sort_bam(0, data, one, sort_by=sb, output=fmt)
        sort_bam(0, io.BytesIO(data) if fmt == "records" else data, many, sort_by=sb, output=fmt, devices=devices)   # read callback / in place
        assert one.getvalue() == many.getvalue(), fmt


def test_sample_sort_skewed_keys_and_fallbacks(oracle):
    """Every record with the same key (one bucket takes all, the others are empty), records longer than the halo between ranges
    and inputs with fewer blocks than GPUs (the single-GPU path takes over) - same bytes as the single-GPU sort; the reference's
    panics stay errors."""
    import struct
    from bam_parallel_b200.sorting import SortBy, sort_bam
    from tests import bamcraft as bc
    devices = _n_sort_devices(3)

    def rec(name, ref_id=0, pos=0, flag=0, tags=b"", seq_len=4):
        body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(name) + 1, 0, 0, 0, flag, seq_len, -1, -1, 0) + name + b"\0" + \
            b"\x11" * ((seq_len + 1) // 2) + b"\x20" * seq_len + tags
        return struct.pack("<I", len(body)) + body

    def bam(recs, payload=0xFF00):
        return bc.bgzf(b"BAM\x01" + struct.pack("<I", 0) + struct.pack("<I", 1) + struct.pack("<I", 4) + b"ch1\0" + struct.pack("<I", 1000)
                       + b"".join(recs), payload=payload)

    same = [rec(b"r%05d" % (i % 7), 0, 5, 0, seq_len=40 + (i % 50)) for i in range(20000)]
    for data in (bam(same, 3000), bam([rec(b"big%d" % i, 0, 9 - i, 0, seq_len=1500000) for i in range(3)], 0xFF00),
                 bam([rec(b"a", 0, 2), rec(b"b", 0, 1)])):
        for sb in SortBy:
            for fmt in ("bodies", "bam"):
                one, many = io.BytesIO(), io.BytesIO()
                sort_bam(0, data, one, sort_by=sb, output=fmt)
                sort_bam(0, data, many, sort_by=sb, output=fmt, devices=devices)
                assert one.getvalue() == many.getvalue(), (sb, fmt, len(data))
    # a block whose CRC32 footer is wrong / whose DEFLATE bytes are damaged, in the middle of the file: the same error either way
    good = bytearray(bam(same, 3000))
    blocks, nb, _, _, _ = bp.scan_blocks(bytes(good))
    mid = blocks[nb // 2]
    for where, code in ((int(mid.in_off + mid.in_len), -14), (int(mid.in_off + 5), None)):
        bad = bytearray(good)
        bad[where] ^= 0x5A
        codes = []
        for kw in ({}, {"devices": devices}):
            with pytest.raises(bp.BamGpuError) as e:
                sort_bam(0, bytes(bad), io.BytesIO(), sort_by=SortBy.Name, **kw)
            codes.append(e.value.code)
        assert codes[0] == codes[1] and (code is None or codes[0] == code), codes
    data = bam([rec(b"q%d" % (i % 3), 0, i, 0x1000 if i == 7777 else 0) for i in range(20000)], 2000)
    with pytest.raises(bp.BamGpuError) as e:
        sort_bam(0, data, io.BytesIO(), sort_by=SortBy.CoordinatesAndStrand, devices=devices)
    assert e.value.code == -31
    data = bam([rec(b"q%d" % (i % 3), 0, i, 0, tags=(b"HIi" + struct.pack("<i", 1)) if i != 7777 else b"") for i in range(20000)], 2000)
    with pytest.raises(bp.BamGpuError) as e:
        sort_bam(0, data, io.BytesIO(), sort_by=SortBy.NameAndMatchMates, devices=devices)
    assert e.value.code == -30
