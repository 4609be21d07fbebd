#!/usr/bin/env python
"""bench.py — BGZF inflate + BAM record decode throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU reader (oracle port)

A "step" is one full pass of the hot path (K1 inflate -> K2 CRC32 -> K3 record chain, fixed fields, keys)
over the whole synthetic BAM of BASELINE.json configs[1]: paired-end 2x150 bp, 50M records, BGZF level 6,
written by the committed seeded generator (tools/bamgen.cpp) because the boxes have no network.

  value  = inflated GB/s, compressed input already resident in HBM when the timed region starts
  e2e    = the same metric through the drop-in Reader API over a HOST buffer: pinned H2D of the compressed
           bytes and D2H of the record bodies + columns inside the timed region
  roofline / cpu_baseline: see DESIGN.md "Measurement"

With N > 1 (torchrun, one rank per GPU) the SAME 50M-record file is cut into N contiguous BGZF block ranges
(BASELINE.json configs[2]: strong scaling); rank g+1 writes the first <= 1 MiB of its inflated bytes into rank g's
inbox over NVLink peer memory (CUDA IPC; bamgpu_engine_edge_exchange) so that the record straddling the shard edge
is whole, and each speculated shard entry is confirmed by its left neighbour.  No collective on the data path.
The e2e line at N > 1 is ONE Reader on rank 0 driving all N GPUs (bamgpu_opts.devices).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "inflated_GB_per_s"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=50_000_000)
    ap.add_argument("--shape", default="pe150")
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--cpu-sample-records", type=int, default=4_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sort", action="store_true", help="skip the sort-path timing (BASELINE.json configs[3], reported beside the headline)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--batch-mib", type=int, default=0, help="compressed MiB per Reader pipeline batch (e2e); 0 = the library's default")
    ap.add_argument("--verify-oracle", action="store_true", help="also fold the CPU oracle's decode of the same file and compare digests (minutes)")
    return ap.parse_args()


def workload_name(a):
    return f"{a.shape} paired-end 2x150bp BAM, {a.records} records, BGZF level {a.level}, payload 0xff00 (BASELINE.json configs[1])"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# input
# ------------------------------------------------------------------------------------------------
def make_input(a, rank: int, world: int, n_records: int, host_group=None):
    """Rank 0 generates; other ranks map the same bytes from /dev/shm. Returns a uint8 numpy array."""
    from bam_parallel_b200 import synth

    if world == 1:
        t0 = time.time()
        cache = os.environ.get("BAMGPU_BENCH_CACHE")   # profiling sessions run bench.py several times on one box
        if cache:
            path = f"{cache}.{a.shape}.{n_records}.L{a.level}.bam"
            if not os.path.exists(path):
                bam = synth.generate(a.shape, n_records, level=a.level, header_block=True)
                with open(path + ".tmp", "wb") as f:
                    f.write(memoryview(bam.array))
                os.replace(path + ".tmp", path)
                bam.close()
            return np.fromfile(path, dtype=np.uint8), None, time.time() - t0
        bam = synth.generate(a.shape, n_records, level=a.level, header_block=True)
        return bam.array, bam, time.time() - t0
    import torch.distributed as dist

    path = f"/dev/shm/bamgpu_bench_{os.environ.get('MASTER_PORT', '0')}_{n_records}.bam"
    t0 = time.time()
    if rank == 0:
        bam = synth.generate(a.shape, n_records, level=a.level, header_block=True)
        with open(path + ".tmp", "wb") as f:
            f.write(memoryview(bam.array))
        os.replace(path + ".tmp", path)
        bam.close()
    dist.barrier(group=host_group)
    arr = np.memmap(path, dtype=np.uint8, mode="r")
    dist.barrier(group=host_group)
    if rank == 0:
        os.unlink(path)  # the mappings keep it alive
    return arr, None, time.time() - t0


BLOCK_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("isize", "<u4"), ("crc32", "<u4"),
                        ("reserved", "<u4")])
M64 = 0xFFFFFFFFFFFFFFFF


def blocks_view(blocks, nb: int) -> np.ndarray:
    """Zero-copy structured numpy view of a ctypes bamgpu_block array."""
    return np.frombuffer(blocks, dtype=BLOCK_DTYPE)[:nb]


def reference_threads():
    """main.rs:84: hash mode reads with min(num_cpus, 20) threads; main.rs:72: the sort modes read with 5."""
    return min(os.cpu_count() or 1, 20), 5


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU reader (oracle port; the Rust crate cannot be built in this image)
# ------------------------------------------------------------------------------------------------
def run_reference(a, rank: int, world: int):
    if rank != 0:
        return
    from bam_parallel_b200 import synth
    from oracle.oracle import COracle

    orc = COracle()
    threads, n5 = reference_threads()
    bam = synth.generate(a.shape, a.records, level=a.level, header_block=True)   # the SAME file the CUDA arm decodes
    for _ in range(min(a.warmup, 1)):
        orc.cpu_reader_run(bam.array, threads)
    secs, ub, recs = [], 0, 0
    for _ in range(a.steps):
        r = orc.cpu_reader_run(bam.array, threads)
        secs.append(r["seconds"])
        ub, recs = r["ubytes"], r["records"]
    t = float(np.mean(secs))
    v = ub / t / 1e9
    # beside it: what bam_binary's sort modes really use (5 reader threads = 3 inflate workers), on a bounded prefix
    r5 = orc.cpu_reader_run(bam.array, n5, min(a.records, a.cpu_sample_records))
    sample = f"the whole workload file per step ({recs} records, {ub} inflated bytes)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": min(a.warmup, 1),
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic (seeded generator tools/bamgen.cpp)", "config": {"workload": workload_name(a), "sample": sample},
        "records_per_s": recs / t,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + f"; {threads} threads = min(num_cpus, 20) as bam_binary's hash mode (main.rs:84); C restatement "
                                            "of the reference reader topology (1 reader + 1 orderer + n-2 zlib inflaters + single-threaded "
                                            "record loop), not the Rust binary",
                         "host_cpus": os.cpu_count(),
                         "n5": {"value": r5["ubytes"] / r5["seconds"] / 1e9, "unit": UNIT, "cores": n5,
                                "sample": f"first {r5['records']} records with 5 reader threads (what bam_binary's sort modes use, main.rs:72)"}},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def add_digest(tot, d):
    for k in ("n_records", "body_bytes", "fields", "bodies"):
        tot[k] = (tot.get(k, 0) + d[k]) & M64
    return tot


def pinned_copy(L, arr):
    """The workload file in page-locked host memory (what the e2e legs read from)."""
    p = L.bamgpu_pinned_alloc(max(arr.size, 1))
    if not p:
        raise SystemExit("bench.py: cannot page-lock the input")
    C.memmove(p, arr.ctypes.data, arr.size)
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(arr.size,))


def run_b200(a, rank: int, world: int, local_rank: int, host_group=None):
    import torch

    import bam_parallel_b200 as bp
    from bam_parallel_b200 import _lib
    from bam_parallel_b200.sharding import _DevView, check_edges, shard_ranges

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

    arr, keep, gen_s = make_input(a, rank, world, a.records, host_group)
    blocks, nb, consumed, total_isize, rc = bp.scan_blocks(arr)
    L = _lib.load()

    # BAM header length: taken from the product's own Reader over the first blocks (host logic + K1)
    hb = min(nb, 8) - 1
    head_bytes = int(blocks[hb].in_off + blocks[hb].in_len + 8) if nb else 0
    with bp.Reader.new(np.ascontiguousarray(arr[:head_bytes]), 4, flags=bp.NO_CRC, batch_bytes=1 << 20) as r0:
        hdr, off_refs = r0.read_header()
    stream_start = 4 + len(hdr)
    n_ref = int.from_bytes(hdr[off_refs:off_refs + 4], "little")

    # ---- sharding: the SAME file at every N, cut into N contiguous BGZF block ranges of ~equal compressed bytes
    # (BASELINE.json configs[2]: strong scaling) ----
    bv = blocks_view(blocks, nb)
    b0, b1 = shard_ranges(bv["in_off"], world)[rank]
    my_nb = b1 - b0
    my_blocks = (_lib.Block * max(my_nb, 1))()
    mv = blocks_view(my_blocks, my_nb)
    mv[:] = bv[b0:b1]
    c0 = int(bv["in_off"][b0]) - 18 if my_nb else 0
    c1 = int(bv["in_off"][b1 - 1]) + int(bv["in_len"][b1 - 1]) + 8 if my_nb else 0
    u0 = int(bv["out_off"][b0]) if my_nb else int(total_isize)
    mv["in_off"] -= np.uint64(c0)
    mv["out_off"] -= np.uint64(u0)
    my_u = int(mv["isize"].sum(dtype=np.uint64))
    my_c_payload = int(mv["in_len"].sum(dtype=np.uint64))
    shard = arr[c0:c1]
    compressed_bytes = int(consumed)

    # ---- e2e: the drop-in Reader over a HOST buffer, H2D + D2H inside the timed region.  Measured FIRST, in a process that
    # holds nothing else on the device (the same loop run after the device-resident phase, with its 45 GB of buffers alive,
    # was measured 15-20 % slower on the D2H side).  N > 1: rank 0 drives all N GPUs through ONE Reader
    # (bamgpu_opts.devices), the other ranks wait on the host. ----
    e2e = e2e_cols = e2e_all = e2e_rec = pcie = sort_multi = None
    reader_digest = None
    if not a.no_e2e and rank == 0:
        if world == 1:
            hbuf = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
            dbuf = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
            pcie = {}
            for name, dst, srcb in (("d2h_GB_per_s", hbuf, dbuf), ("h2d_GB_per_s", dbuf, hbuf)):
                dst.copy_(srcb, non_blocking=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    dst.copy_(srcb, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                pcie[name] = round(3 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
            del hbuf, dbuf
        if world > 1:
            pcie = pcie_aggregate(torch, world)
        pin_ptr, parr = pinned_copy(L, arr[:compressed_bytes])
        devs = list(range(world)) if world > 1 else None
        e2e, reader_digest = run_e2e(a, bp, parr, torch, int(total_isize), devices=devs, want_digest=True)
        if world == 1:
            e2e_all, _ = run_e2e(a, bp, parr, torch, int(total_isize), flags=bp.COPY_DATA | bp.COPY_COLUMNS,
                                 api="same, plus all 21 SoA columns (fixed fields, offsets, sort keys) D2H")
            e2e_cols, _ = run_e2e(a, bp, parr, torch, int(total_isize), flags=bp.COPY_COLUMNS,
                                  api="only the SoA columns + sort keys come back (what a host-side sort of keys needs); bodies stay in HBM")
            e2e_rec = run_e2e_next_rec(a, bp, parr, int(total_isize))
        if world > 1 and not a.no_sort:
            # the sort over all N GPUs behind bamgpu_sort_bam (sample sort, peer-copy exchange): host buffer in,
            # sorted bodies out (written to a sink that drops them); device stage times are the slowest GPU's
            from bam_parallel_b200.sorting import SortBy, sort_bam
            sort_multi = {}
            for sb in (SortBy.CoordinatesAndStrand, SortBy.Name):
                try:
                    for rep in range(2):
                        t0 = time.perf_counter()
                        stt = sort_bam(0, parr, None, sort_by=sb, devices=devs)
                        wall = time.perf_counter() - t0
                except Exception as ex:      # reported, never fatal for the headline
                    sort_multi[sb.name] = {"error": repr(ex)}
                    continue
                sort_multi[sb.name] = {"wall_s": round(wall, 3), "ms_inflate": round(stt["ms_inflate"], 2), "ms_records": round(stt["ms_records"], 2),
                                       "ms_partition_and_key_sort": round(stt["ms_sort"], 2), "ms_gathers": round(stt["ms_gather"], 2),
                                       "ms_d2h_to_sink": round(stt["ms_d2h"], 1), "p2p_bytes": int(stt["p2p_bytes"]), "records": a.records}
            sort_multi["note"] = ("bamgpu_sort_bam(devices = all N GPUs) on rank 0: N block ranges decoded side by side, sampled splitters, stable "
                                  "partition, N x N peer copies of (block_size, body) streams, K3 + stable sort per bucket; output bytes = the "
                                  "single-GPU sort's (tests/test_gpu_multi.py::test_sample_sort_over_several_gpus)")
        L.bamgpu_pinned_free(pin_ptr)
        del parr
    if dist is not None:
        dist.barrier(group=host_group)     # host-side wait: no NCCL kernel spins on the idle GPUs meanwhile

    eng = bp.Engine(device=local_rank)
    eng.set_n_ref(n_ref)
    d_comp = eng.upload(np.ascontiguousarray(shard))
    my_start = stream_start if rank == 0 else 0

    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    step_no = [0]
    if world > 1:
        # shard-edge exchange over NVLink peer memory (CUDA IPC between the per-GPU processes): the 80-byte inbox handles
        # are passed around once; the data path has no collective and no NCCL call
        mine = eng.edge_open()
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=host_group)
        eng.edge_connect(handles[rank - 1] if rank > 0 else None, handles[rank + 1] if rank < world - 1 else None)

    def step():
        """One pass of the hot path over this rank's shard, compressed bytes already resident in HBM."""
        if world == 1:
            batch, tail, rc = eng.decode(d_comp, my_blocks, my_nb, my_start, True, 0, sptr)
            return batch, tail
        step_no[0] += 1
        eng.inflate(d_comp, my_blocks, my_nb, bp.DEFER_CRC, sptr)   # K1 over the shard's BGZF blocks; K2 on a side stream beside what follows
        eng.edge_exchange(step_no[0], sptr)                      # my head -> left neighbour; right neighbour's head behind my bytes
        flags = bp.SPECULATIVE_START if rank > 0 else 0
        batch, tail, rc = eng.records(my_start, rank == world - 1, flags, sptr)   # K3: records that START in my bytes
        return batch, tail

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        batch, tail = step()
    n_rec_local = batch.n_records
    chain_repairs = batch.chain_repairs
    # ---- verification (outside the timed region) ----
    verify = {}
    if world > 1:
        # one u64 per edge: shard g's chain must leave its bytes exactly where shard g+1's chain begins
        info = [None] * world
        dist.all_gather_object(info, (int(tail - my_u), int(batch.chain_start), int(batch.n_records)), group=host_group)
        check_edges([x[0] for x in info], [x[1] for x in info], [x[2] for x in info], a.records)
        index_base = sum(x[2] for x in info[:rank])
    else:
        # size-independent properties at full size: every block passed its CRC32 + ISIZE check (else decode raises), the
        # chain covers the stream exactly, and the record count is the generator's
        assert batch.n_records == a.records and tail == my_u, (batch.n_records, a.records, tail, my_u)
        rl = torch.as_tensor(_DevView(batch.dev["rec_len"], 4 * batch.n_records), device="cuda").view(torch.int32)
        assert int(rl.sum(dtype=torch.int64).item()) + 4 * batch.n_records + stream_start == my_u, "record lengths do not tile the stream"
        index_base = 0
    # DIGEST SPEC sums of what each rank decoded (fields, offsets, keys, every body byte), added up over the ranks
    dg = eng.digest(index_base=index_base, stream_base=u0, cuda_stream=sptr)
    if world > 1:
        alld = [None] * world
        dist.all_gather_object(alld, dg, group=host_group)
        dg = {}
        for d in alld:
            add_digest(dg, d)
    verify["digest"] = {k: int(v) for k, v in dg.items()}
    if rank == 0:
        assert dg["n_records"] == a.records, (dg["n_records"], a.records)
        if reader_digest is not None:
            assert reader_digest == dg, ("the Reader's digest differs from the sharded engines'", reader_digest, dg)
            verify["reader_vs_engine_digest"] = "equal"
        if a.verify_oracle:
            verify["oracle"] = verify_with_oracle(a, arr[:compressed_bytes], dg)
    sync_all()
    eng.reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    sync_all()
    ms_total = ev0.elapsed_time(ev1)
    st = eng.stats()
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        cnt = torch.tensor([float(n_rec_local), float(my_u), float(st["kernel_launches"]), float(st["p2p_bytes"])], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        n_rec_all, u_all, launches, p2p = int(cnt[0].item()), int(cnt[1].item()), int(cnt[2].item()), int(cnt[3].item())
        stage = torch.tensor([st["ms_inflate"], st["ms_crc"], st["ms_records"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(stage, op=dist.ReduceOp.MAX)
        stage_max = [float(x) / a.steps for x in stage.tolist()]
    else:
        n_rec_all, u_all, launches, p2p = n_rec_local, my_u, st["kernel_launches"], st["p2p_bytes"]
        stage_max = [st["ms_inflate"] / a.steps, st["ms_crc"] / a.steps, st["ms_records"] / a.steps]
    ms_step = ms_total / a.steps
    value = u_all / (ms_step * 1e-3) / 1e9
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (K1 inflate), rank 0's shard ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    k1_ms = st["ms_inflate"] / a.steps
    k1_bytes = my_c_payload + my_u   # algorithmic: compressed payload read once + inflated bytes written once
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    traffic = traffic_src = None
    try:   # DRAM bytes of one K1 launch: NOT measured in this run (that needs ncu) - taken from the committed capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        if tj.get("records") == a.records and world == 1:
            traffic = tj["dram_bytes_per_launch"]
            traffic_src = f"ncu --set full capture {tj.get('source')} at commit {tj.get('commit')}; not re-measured in this run"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k1_inflate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes,
                "ms_per_launch": k1_ms,
                "stage_ms_per_step": {"k1_inflate": k1_ms, "k2_crc32": st["ms_crc"] / a.steps, "k3_records": st["ms_records"] / a.steps},
                "stage_ms_max_over_ranks": {"k1_inflate": stage_max[0], "k2_crc32": stage_max[1], "k3_records": stage_max[2]}}

    # ---- sort path (BASELINE.json configs[3]): stable sort of the decoded records + body gather, device-resident ----
    sort_info = None
    if world == 1 and not a.no_sort:
        from bam_parallel_b200.sorting import SortBy
        sort_info = {}
        for sb in (SortBy.CoordinatesAndStrand, SortBy.Name):
            for rep in range(2):
                eng.reset_stats()
                _, _, srt = eng.sort(int(sb), flags=0, cuda_stream=sptr)
                stt = eng.stats()
            nb_out = int(srt.bodies_len)
            sort_info[sb.name] = {"ms_key_sort": round(stt["ms_sort"], 2), "ms_body_gather": round(stt["ms_gather"], 2),
                                  "records_per_s": a.records / ((stt["ms_sort"] + stt["ms_gather"]) * 1e-3),
                                  "gather_GB_per_s": 2 * nb_out / (stt["ms_gather"] * 1e-3) / 1e9 if stt["ms_gather"] > 0 else None,
                                  "output_bytes": nb_out}
        # K6: the sorted bodies (still in HBM) through the BGZF writers: stored blocks, fixed-Huffman blocks, per-block Huffman codes
        try:
            bound = int(L.bamgpu_bgzf_bound(nb_out))
            d_bg = L.bamgpu_device_alloc(eng._h, bound)
            writer = {}
            if d_bg:
                for level, name in ((0, "stored"), (1, "fixed_huffman"), (2, "per_block_huffman")):
                    for rep in range(2):
                        eng.reset_stats()
                        wrote = C.c_size_t()
                        rcw = L.bamgpu_engine_bgzf_write(eng._h, srt.bodies, nb_out, d_bg, bound, C.byref(wrote), 1, level, sptr)
                        assert rcw == 0, rcw
                        stt = eng.stats()
                    writer[name] = {"ms": round(stt["ms_bgzf_write"], 2), "input_GB_per_s": nb_out / (stt["ms_bgzf_write"] * 1e-3) / 1e9,
                                    "output_bytes": int(wrote.value), "ratio": round(nb_out / max(wrote.value, 1), 3)}
                L.bamgpu_device_free(eng._h, d_bg)
                writer["note"] = ("bamgpu_engine_bgzf_write over the name-sorted record bodies in HBM (0xff00-byte payloads, CRC32 by K2's warp CRC): "
                                  "level 0 stored blocks, 1 fixed Huffman code, 2 a Huffman code built per block (K6, csrc/deflate.cu); "
                                  "read back by zlib, the oracle and K1 in tests/test_gpu_deflate.py")
                sort_info["bgzf_writer"] = writer
        except Exception as ex:   # reported, never fatal for the headline
            sort_info["bgzf_writer"] = {"error": repr(ex)}
        sort_info["note"] = ("bamgpu_engine_sort on the batch just decoded: output bytes = bam_binary's (record bodies in stable sorted "
                             "order, sort.rs:643); parity vs the oracle in tests/test_gpu_sort.py and tests/test_gpu_fullsize.py; gather GB/s "
                             "counts bodies read + written")

    # ---- K5: variable-length field decode of the same batch, reported beside the headline ----
    fields_info = None
    if world == 1 and not a.no_sort:
        for rep in range(2):
            fd = eng.decode_fields(7, 0, cuda_stream=sptr)
        out_bytes = sum(fd[k]["total"] for k in ("cigar", "seq", "qual"))
        ms = fd["ms_measure"] + fd["ms_decode"]
        fields_info = {"ms_measure": round(fd["ms_measure"], 2), "ms_decode": round(fd["ms_decode"], 2), "records_per_s": a.records / (ms * 1e-3),
                       "decoded_bytes": out_bytes, "decoded_GB_per_s": out_bytes / (ms * 1e-3) / 1e9,
                       "bad": {k: fd[k]["bad"] for k in ("cigar", "seq", "qual")},
                       "note": "bamgpu_engine_decode_fields on the batch just decoded: CIGAR strings (decode_cigar over get_cigar), sequence "
                               "strings (decode_seq), raw qualities; parity vs the oracle in tests/test_gpu_fields.py and tests/test_gpu_fullsize.py"}

    # ---- CPU baseline (rank 0, N=1 only): the whole file, min(num_cpus, 20) threads as bam_binary's hash mode ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle.oracle import COracle
        orc = COracle()
        threads, n5 = reference_threads()
        r = orc.cpu_reader_run(arr, threads)
        r5 = orc.cpu_reader_run(arr, n5, min(a.records, a.cpu_sample_records))
        cpu = {"value": r["ubytes"] / r["seconds"] / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "records_per_s": r["records"] / r["seconds"], "host_cpus": os.cpu_count(),
               "sample": f"the whole file once ({r['records']} records, {r['ubytes']} inflated bytes, {r['seconds']:.2f} s) with min(num_cpus, 20) "
                         "threads (main.rs:84); C restatement of the reference reader (zlib inflate), not the Rust binary",
               "n5": {"value": r5["ubytes"] / r5["seconds"] / 1e9, "unit": UNIT, "cores": n5,
                      "sample": f"first {r5['records']} records with 5 reader threads (bam_binary's sort modes, main.rs:72)"}}

    if rank == 0:
        limiter = None
        if world > 1:
            limiter = ("no collective on the data path; per-GPU step = K1 over 1/N of the blocks, which at N >= 4 is less than one wave "
                       "of K1's 56.8 k streams, so K1's per-block latency (~9 ms for a 64 KiB block) bounds the step, plus the serial "
                       "edge hand-over (host wait for the neighbour's halo)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic (seeded generator tools/bamgen.cpp)",
            "config": {"workload": workload_name(a), "records": n_rec_all,
                       "inflated_bytes": u_all, "compressed_bytes": compressed_bytes, "bgzf_blocks": int(nb),
                       "sharding": (f"{world} contiguous BGZF block ranges of the same file, one per GPU; the head of shard g+1 (<= {bp._lib.MAX_HALO} B) is "
                                    "written into shard g's inbox over NVLink peer memory (CUDA IPC, bamgpu_engine_edge_exchange), no collective"
                                    if world > 1 else "single GPU"),
                       "l2": "inputs (compressed + inflated) far larger than the 126 MB L2; no flush needed",
                       "generator_seconds": round(gen_s, 1), "k3_chain_repairs": chain_repairs, "p2p_bytes_per_step": p2p // max(a.steps, 1)},
            "records_per_s": n_rec_all / (ms_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_with_columns": e2e_all, "e2e_columns_only": e2e_cols,
            "e2e_next_rec": e2e_rec, "pcie_pinned_copy": pcie, "sort": sort_info, "sort_multi_gpu": sort_multi, "fields": fields_info, "gpu_launches": launches,
            "verify": verify, "scaling_limiter": limiter, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if keep is not None:
        keep.close()


def pcie_aggregate(torch, n_dev, mib=512, reps=4):
    """What the box gives N plain pinned copies running at once (one per GPU): the bound of the N-GPU e2e line.
    On the round-2 boxes it is far below N x the single-GPU figure (the GPUs share host links)."""
    out = {}
    hb = [torch.empty(mib << 20, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
    db = [torch.empty(mib << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
    st = [torch.cuda.Stream(device=d) for d in range(n_dev)]
    for name in ("d2h", "h2d"):
        for rep in range(2):       # first pass warms up
            for d in range(n_dev):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for _ in range(reps):
                for d in range(n_dev):
                    with torch.cuda.stream(st[d]):
                        (hb[d] if name == "d2h" else db[d]).copy_(db[d] if name == "d2h" else hb[d], non_blocking=True)
            for d in range(n_dev):
                torch.cuda.synchronize(d)
            dt = time.perf_counter() - t0
        out[f"{name}_GB_per_s_all_{n_dev}_gpus_at_once"] = round(n_dev * reps * (mib << 20) / dt / 1e9, 1)
    return out


def verify_with_oracle(a, arr, dg):
    """--verify-oracle: the CPU restatement of the reference over the same file, folded with the same DIGEST SPEC."""
    from oracle.oracle import COracle
    orc = COracle()
    th = min(os.cpu_count() or 4, 32)
    t0 = time.time()
    u = orc.inflate_all_mt(arr, th)
    _, _, start = orc.read_header(u)
    off, ln = orc.walk_records(u, start)
    cols, bf, bt = orc.extract_mt(u, off, ln, False, th)
    ref = orc.digest(u, off, ln, cols, th)
    orc.release(u)
    assert ref == dg, ("GPU digest differs from the oracle", ref, dg)
    return {"equal": True, "seconds": round(time.time() - t0, 1), "threads": th}


def run_e2e(a, bp, parr, torch, u_all, flags=None, api=None, devices=None, want_digest=False):
    """Reader::new(host bytes) -> read_header -> next_batch until EOF, with the batch's host mirrors filled as `flags` say;
    default: bodies + (rec_off, rec_len), i.e. exactly what records().next_rec() serves from.  The source is page-locked
    host memory; every step copies it to the GPU(s) and the results back."""
    if flags is None:
        flags = bp.COPY_DATA | bp.COPY_OFFSETS
    per_rec = 74 if flags & bp.COPY_COLUMNS else (12 if flags & bp.COPY_OFFSETS else 0)   # bytes of columns D2H per record
    full = bool(flags & bp.COPY_DATA)
    times, d2h = [], 0
    warm = 2   # the first Readers of a process page-lock their buffers (parked in the library's cache afterwards)
    threads = reference_threads()[0]
    digest = None
    kw = {}
    if devices:
        kw["devices"] = devices
    if a.batch_mib:
        kw["batch_bytes"] = a.batch_mib << 20
    for it in range(warm + a.e2e_steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2h = 0
        dg = {} if (want_digest and it == 0) else None
        with bp.Reader.new(parr, threads, flags=flags, **kw) as r:   # main.rs:84 thread count
            r.read_header()
            n = 0
            while True:
                b = r.next_batch()
                if b is None:
                    break
                n += b.n_records
                d2h += (b.data_len - b.data_start if full else 0) + per_rec * b.n_records
                if dg is not None:
                    add_digest(dg, r.batch_digest(b))
            st = r.stats()
        for d in (devices or [torch.cuda.current_device()]):
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        if dg is not None:
            digest = dg
        if it >= warm:
            times.append(dt)
    t = float(np.mean(times))
    return {"value": u_all / t / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(parr.size), "d2h_bytes_per_step": int(d2h),
            "records_per_s": n / t, "ms_per_step": t * 1e3, "steps": a.e2e_steps, "warmup": warm, "n_gpus": len(devices) if devices else 1,
            "device_ms": {k: round(st[k], 1) for k in ("ms_inflate", "ms_crc", "ms_records", "ms_d2h")}, "batches": st["batches"],
            "p2p_bytes": st["p2p_bytes"], "step_ms": [round(x * 1e3, 1) for x in times],
            "api": api or "Reader.new(page-locked host bytes, min(cpus, 20)" + (f", devices={devices}" if devices else "") +
                          ") / read_header / next_batch: every record body + (offset, length) in pinned host memory = what "
                          "records().next_rec() lends (the reference hands out bodies, nothing else)"}, digest


def run_e2e_next_rec(a, bp, parr, u_all):
    """The reference's own consumer loop: records().next_rec() once per record (main.rs:83-99), from C
    (tools/next_rec_harness.c = what the Rust shim does per call), H2D + D2H inside the timed region."""
    hp = os.path.join(ROOT, "bam_parallel_b200", "lib", "libbamharness.so")
    if not os.path.exists(hp):
        return None
    H = C.CDLL(hp)
    H.harness_next_rec_all.restype = C.c_int
    H.harness_next_rec_all.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3 + [C.POINTER(C.c_double)]
    times = []
    n = C.c_uint64()
    for it in range(1 + max(a.e2e_steps - 1, 1)):
        t0 = time.perf_counter()
        with bp.Reader.new(parr, reference_threads()[0], **({"batch_bytes": a.batch_mib << 20} if a.batch_mib else {})) as r:
            r.read_header()
            nb, ck, secs = C.c_uint64(), C.c_uint64(), C.c_double()
            rc = H.harness_next_rec_all(r._h, C.byref(n), C.byref(nb), C.byref(ck), C.byref(secs))
            if rc != 0:
                return {"error": rc}
        dt = time.perf_counter() - t0
        if it >= 1:
            times.append(dt)
    t = float(np.mean(times))
    return {"value": u_all / t / 1e9, "unit": UNIT, "records_per_s": n.value / t, "ms_per_step": t * 1e3, "calls_per_step": n.value,
            "api": "Reader.new / read_header / bamgpu_next_rec once per record from a C loop (tools/next_rec_harness.c), each body touched"}


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    host_group = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")   # handles, verification and "wait while rank 0 measures e2e": host side
    try:
        run_b200(a, rank, world, local_rank, host_group)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
