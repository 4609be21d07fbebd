"""Host-side logic of block-range sharding (bam_parallel_b200/sharding.py): range cutting, the neighbour-only head
exchange over torch.distributed (gloo, world_size 2 and 3, CPU) and the per-edge confirmation of speculated entries.
The record walk on each shard is done here by the oracle (this is a test of the plumbing, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_everything():
    from bam_parallel_b200.sharding import shard_ranges
    rng = np.random.default_rng(1)
    for nb in (0, 1, 2, 7, 100, 1001):
        in_off = np.cumsum(rng.integers(30, 65000, size=nb)).astype(np.uint64) + np.uint64(18) if nb else np.zeros(0, np.uint64)
        for world in (1, 2, 3, 4, 8):
            r = shard_ranges(in_off, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == nb
            assert all(a <= b for a, b in r) and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            if nb >= 100:   # balanced by compressed bytes
                sizes = [int(in_off[b - 1]) - int(in_off[a]) for a, b in r if b > a]
                assert max(sizes) < 1.5 * (int(in_off[-1]) / world) + 70000


def test_check_edges():
    from bam_parallel_b200.sharding import check_edges
    check_edges([10, 0, 0], [99, 10, 0], [5, 6, 7], 18)
    with pytest.raises(ValueError):
        check_edges([10, 0], [99, 11], [5, 6])
    with pytest.raises(ValueError):
        check_edges([10, 0], [99, 10], [5, 6], 12)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, halo, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        from bam_parallel_b200 import synth
        from bam_parallel_b200.sharding import EdgeExchange, shard_ranges
        from oracle.oracle import COracle

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        orc = COracle()
        data = synth.generate("pe150", 6000, level=1, payload=3000, threads=1).tobytes()
        u, sizes = orc.inflate_all(data)
        _, _, start = orc.read_header(u)
        off_all, len_all = orc.walk_records(u, start)
        blocks = orc.scan_blocks(data)
        in_off = (blocks["coff"] + np.uint64(18)).astype(np.uint64)
        (b0, b1) = shard_ranges(in_off, world)[rank]
        uoff = np.concatenate([[0], np.cumsum(sizes.astype(np.uint64))]).astype(np.int64)
        lo, hi = int(uoff[b0]), int(uoff[b1])
        own = u[lo:hi]
        ex = EdgeExchange(rank, world, torch, dist, halo=halo, device="cpu")
        got = ex.exchange_tensors(torch.from_numpy(own[:halo].copy()) if rank > 0 else None)
        buf = np.concatenate([own, got.numpy()]) if got is not None else own
        if rank < world - 1:
            assert np.array_equal(buf[len(own):len(own) + min(halo, u.size - hi)], u[hi:hi + halo])
        # records STARTING in my bytes, found from my true first record start (the CUDA path guesses it; here: oracle)
        first = off_all[np.searchsorted(off_all, lo + 4)] - 4 - lo if rank > 0 else start
        mine = [(o, l) for o, l in zip(off_all.tolist(), len_all.tolist()) if lo <= o - 4 < hi]
        for o, l in mine:
            assert o - lo + l <= buf.size, "straddling record larger than the halo"
            assert bytes(buf[o - lo:o - lo + l]) == bytes(u[o:o + l])
        last_o, last_l = mine[-1]
        exit_rel = (last_o + last_l) - hi if rank < world - 1 else 0
        ex.verify(exit_rel, int(first), len(mine), off_all.size)
        # the digests of the shards add up to the digest of the file (what bench.py gathers over its host-side group at N > 1)
        cols, _, _ = orc.extract(u, off_all, len_all, True)
        a = int(np.searchsorted(off_all, lo + 4))
        b = a + len(mine)
        sub = {k: np.ascontiguousarray(v[a:b]) for k, v in cols.items()}
        d = orc.digest(buf, np.ascontiguousarray(off_all[a:b] - np.uint64(lo)), np.ascontiguousarray(len_all[a:b]), sub, 2,
                       index_base=a, stream_base=lo)
        alld = [None] * world
        dist.all_gather_object(alld, d)
        tot = {"n_records": 0, "body_bytes": 0, "fields": 0, "bodies": 0}
        for x in alld:
            for k in tot:
                tot[k] = (tot[k] + x[k]) & 0xFFFFFFFFFFFFFFFF
        assert tot == orc.digest(u, off_all, len_all, cols, 2), "shard digests do not add up"
        dist.destroy_process_group()
        q.put((rank, "ok", len(mine)))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0))


@pytest.mark.parametrize("world", [2, 3])
def test_edge_exchange_gloo(world):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, 4096, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
