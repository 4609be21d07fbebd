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

With N > 1 (torchrun, one rank per GPU) the workload is ONE BAM of N x 50M records (the 50M-record body repeated
N times behind one header: weak scaling, per-GPU work fixed), sharded by contiguous BGZF block ranges; rank g
receives the first 1 MiB of rank g+1's inflated bytes (point-to-point) so that the record straddling the shard
edge is whole, and each speculated shard entry is confirmed by its left neighbour.  No collective on the data path.
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
    ap.add_argument("--batch-mib", type=int, default=256, help="compressed MiB per Reader pipeline batch (e2e)")
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
def make_input(a, rank: int, world: int, n_records: int):
    """Rank 0 generates; other ranks map the same bytes from /dev/shm. Returns a uint8 numpy array."""
    from bam_parallel_b200 import synth

    if world == 1:
        t0 = time.time()
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
    dist.barrier()
    arr = np.memmap(path, dtype=np.uint8, mode="r")
    dist.barrier()
    if rank == 0:
        os.unlink(path)  # the mappings keep it alive
    return arr, None, time.time() - t0


BLOCK_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("isize", "<u4"), ("crc32", "<u4"),
                        ("reserved", "<u4")])


def blocks_view(blocks, nb: int) -> np.ndarray:
    """Zero-copy structured numpy view of a ctypes bamgpu_block array."""
    return np.frombuffer(blocks, dtype=BLOCK_DTYPE)[:nb]


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU reader (oracle port; the Rust crate cannot be built in this image)
# ------------------------------------------------------------------------------------------------
def run_reference(a, rank: int, world: int):
    if rank != 0:
        return
    from bam_parallel_b200 import synth
    from oracle.oracle import COracle

    orc = COracle()
    cores = os.cpu_count() or 1
    n = min(a.records, a.cpu_sample_records)
    bam = synth.generate(a.shape, n, level=a.level)
    for _ in range(a.warmup):
        orc.cpu_reader_run(bam.array, cores)
    secs, ub, recs = [], 0, 0
    for _ in range(a.steps):
        r = orc.cpu_reader_run(bam.array, cores)
        secs.append(r["seconds"])
        ub, recs = r["ubytes"], r["records"]
    t = float(np.mean(secs))
    v = ub / t / 1e9
    sample = f"first {n} records of the workload ({ub} inflated bytes) per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
        "records_per_s": recs / t,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + "; C restatement of the reference reader topology (1 reader + 1 orderer + n-2 zlib "
                                            "inflaters + single-threaded record loop), not the Rust binary"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def run_b200(a, rank: int, world: int, local_rank: int):
    import torch

    import bam_parallel_b200 as bp
    from bam_parallel_b200 import _lib
    from bam_parallel_b200.sharding import _DevView

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

    arr, keep, gen_s = make_input(a, rank, world, a.records)
    blocks, nb, consumed, total_isize, rc = bp.scan_blocks(arr)
    L = _lib.load()

    # BAM header length: taken from the product's own Reader over the first blocks (host logic + K1)
    hb = min(nb, 8) - 1
    head_bytes = int(blocks[hb].in_off + blocks[hb].in_len + 8) if nb else 0
    with bp.Reader.new(np.ascontiguousarray(arr[:head_bytes]), 4, flags=bp.NO_CRC) as r0:
        hdr, off_refs = r0.read_header()
    stream_start = 4 + len(hdr)
    n_ref = int.from_bytes(hdr[off_refs:off_refs + 4], "little")

    from bam_parallel_b200.sharding import shard_ranges
    bv = blocks_view(blocks, nb)
    # The workload at N GPUs is ONE BAM of N x a.records records: header block + the record blocks repeated N times
    # (+ EOF marker), sharded by contiguous block ranges of ~equal compressed bytes.  Per-GPU work is what N=1 does.
    assert nb >= 3 and int(bv["isize"][0]) == stream_start and int(bv["isize"][nb - 1]) == 0, "unexpected generator layout"
    nbody = nb - 2
    src = np.concatenate([[0], np.tile(np.arange(1, nb - 1), world), [nb - 1]])      # virtual block -> block of the file
    csz = bv["in_len"][src].astype(np.uint64) + np.uint64(26)
    v_in_off = np.concatenate([[0], np.cumsum(csz)[:-1]]).astype(np.uint64) + np.uint64(18)
    b0, b1 = shard_ranges(v_in_off, world)[rank]
    my_nb = b1 - b0
    my_src = src[b0:b1]
    my_blocks = (_lib.Block * max(my_nb, 1))()
    mv = blocks_view(my_blocks, my_nb)
    mv[:] = bv[my_src]
    my_csz = csz[b0:b1]
    mv["in_off"] = np.concatenate([[0], np.cumsum(my_csz)[:-1]]).astype(np.uint64) + np.uint64(18)
    mv["out_off"] = np.concatenate([[0], np.cumsum(mv["isize"].astype(np.uint64))[:-1]]).astype(np.uint64)
    my_u = int(mv["isize"].sum(dtype=np.uint64))
    my_c_payload = int(mv["in_len"].sum(dtype=np.uint64))
    # materialise the shard's compressed bytes: runs of consecutive file blocks are copied as one slice
    f_lo = bv["in_off"].astype(np.int64) - 18
    f_hi = f_lo + bv["in_len"].astype(np.int64) + 26
    pieces, i = [], 0
    brk = np.flatnonzero(np.diff(my_src) != 1) + 1
    for lo_i, hi_i in zip(np.concatenate([[0], brk]), np.concatenate([brk, [my_nb]])):
        if hi_i > lo_i:
            pieces.append(arr[f_lo[my_src[lo_i]]:f_hi[my_src[hi_i - 1]]])
    shard = np.concatenate(pieces) if len(pieces) != 1 else np.ascontiguousarray(pieces[0])
    assert shard.size == int(my_csz.sum())
    c0, c1 = 0, shard.size
    total_records = a.records * world
    # ---- e2e (N=1): drop-in Reader over the HOST buffer, H2D + D2H inside the timed region.  Measured FIRST, in a process
    # that holds nothing else on the device or in pinned memory (the same loop run after the device-resident phase, with its
    # 45 GB of buffers alive, was measured 15-20 % slower on the D2H side; tools/debug/e2e_timeline.py is the standalone form) ----
    e2e = e2e_cols = e2e_all = pcie = None
    if not a.no_e2e and world == 1:
        # what the link gives a plain pinned copy: the bound of every e2e number below
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
    if not a.no_e2e:
        if world == 1:
            e2e = run_e2e(a, bp, arr, c0, c1, rank, world, dist, torch, int(total_isize) * 1)
            e2e_all = run_e2e(a, bp, arr, c0, c1, rank, world, dist, torch, int(total_isize), flags=bp.COPY_DATA | bp.COPY_COLUMNS,
                              api="same, plus all 21 SoA columns (fixed fields, offsets, sort keys) D2H")
            e2e_cols = run_e2e(a, bp, arr, c0, c1, rank, world, dist, torch, int(total_isize), flags=bp.COPY_COLUMNS,
                               api="only the SoA columns + sort keys come back (what a host-side sort of keys needs); bodies stay in HBM")


    eng = bp.Engine(device=local_rank)
    eng.set_n_ref(n_ref)
    d_comp = eng.upload(shard)
    my_start = stream_start if rank == 0 else 0

    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    halo = edge = None
    if world > 1:
        from bam_parallel_b200.sharding import EdgeExchange
        edge = EdgeExchange(rank, world, torch, dist)

    def step():
        """One pass of the hot path over this rank's shard, compressed bytes already resident in HBM."""
        if world == 1:
            batch, tail, rc = eng.decode(d_comp, my_blocks, my_nb, my_start, True, 0, sptr)
            return batch, tail
        # K1 + K2 over the shard's BGZF blocks
        eng.inflate(d_comp, my_blocks, my_nb, 0, sptr)
        # shard-edge exchange (point-to-point, neighbour only): the head of my inflated bytes goes to the left-hand
        # shard, the right-hand shard's head lands behind my bytes, so the record straddling the edge is whole
        edge.exchange(eng, sptr)
        # K3 over own bytes (+ halo): records that START in my bytes
        flags = bp.SPECULATIVE_START if rank > 0 else 0
        batch, tail, rc = eng.records(my_start, rank == world - 1, flags, sptr)
        return batch, tail

    def step_host(h_shard, h_stage, stage_bytes):
        """The sharded pass with HOST buffers: pinned H2D of the shard's compressed bytes, K1-K3 with the edge exchange,
        then record bodies (through a pinned staging buffer) and SoA columns back to the host."""
        L.bamgpu_memcpy_h2d(eng._h, d_comp, h_shard, shard.size)
        eng.inflate(d_comp, my_blocks, my_nb, 0, sptr)
        edge.exchange(eng, sptr)
        flags = (bp.SPECULATIVE_START if rank > 0 else 0) | bp.COPY_COLUMNS
        batch, tail, rc = eng.records(my_start, rank == world - 1, flags, sptr)
        ptr, n = eng.data()
        for o in range(0, n, stage_bytes):
            L.bamgpu_memcpy_d2h(eng._h, h_stage, ptr + o, min(stage_bytes, n - o))
        return batch

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        batch, tail = step()
    n_rec_local = batch.n_records
    chain_repairs = batch.chain_repairs
    if world > 1:
        # one u64 per edge: shard g's chain must leave its bytes exactly where shard g+1's chain begins
        edge.verify(tail - my_u, batch.chain_start, batch.n_records, total_records)
    else:
        # size-independent properties at full size: every block passed its CRC32 + ISIZE check (else decode raises), the
        # chain covers the stream exactly, and the record count is the generator's
        assert batch.n_records == a.records and tail == my_u, (batch.n_records, a.records, tail, my_u)
        rl = torch.as_tensor(_DevView(batch.dev["rec_len"], 4 * batch.n_records), device="cuda").view(torch.int32)
        assert int(rl.sum(dtype=torch.int64).item()) + 4 * batch.n_records + stream_start == my_u, "record lengths do not tile the stream"
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
        cnt = torch.tensor([float(n_rec_local), float(my_u), float(st["kernel_launches"])], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        n_rec_all, u_all, launches = int(cnt[0].item()), int(cnt[1].item()), int(cnt[2].item())
    else:
        n_rec_all, u_all, launches = n_rec_local, my_u, st["kernel_launches"]
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
    traffic = None
    try:   # DRAM bytes of one K1 launch from the committed ncu --set full capture of this same workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        if tj.get("records") == a.records:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k1_inflate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
                "stage_ms_per_step": {"k1_inflate": k1_ms, "k2_crc32": st["ms_crc"] / a.steps, "k3_records": st["ms_records"] / a.steps}}

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
        sort_info["note"] = ("bamgpu_engine_sort on the batch just decoded: output bytes = bam_binary's (record bodies in stable sorted "
                             "order, sort.rs:643); parity vs the oracle in tests/test_gpu_sort.py; gather GB/s counts bodies read + written")

    # ---- e2e: drop-in Reader over the HOST buffer (H2D + D2H inside the timed region) ----
    # ---- e2e (N>1): per-rank pinned H2D / D2H around the sharded decode ----
    if not a.no_e2e and world > 1:
        e2e = run_e2e_sharded(a, bp, eng, shard, d_comp, step_host, sync_all, rank, world, dist, torch, u_all, my_u, n_rec_local)

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle.oracle import COracle
        orc = COracle()
        cores = os.cpu_count() or 1
        n = min(a.records, a.cpu_sample_records)
        # the first n records' worth of BGZF blocks of the same file
        r = orc.cpu_reader_run(arr, cores, n)
        cpu = {"value": r["ubytes"] / r["seconds"] / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "records_per_s": r["records"] / r["seconds"],
               "sample": f"first {r['records']} records ({r['ubytes']} inflated bytes) of the same file, {r['seconds']:.2f} s; "
                         "C restatement of the reference reader (zlib inflate), not the Rust binary"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic (seeded generator tools/bamgen.cpp)" + (f"; one BAM of {world} x {a.records} records: the record blocks "
                    f"of the {a.records}-record file repeated {world} times behind one header" if world > 1 else ""),
            "config": {"workload": workload_name(a) + (f" per GPU, x{world}" if world > 1 else ""), "records": n_rec_all,
                       "inflated_bytes": u_all, "compressed_bytes": int(csz.sum()),
                       "bgzf_blocks": int(src.size), "sharding": f"{world} contiguous BGZF block ranges, neighbour P2P halo "
                       f"of {bp._lib.MAX_HALO} B per shard edge, no collective" if world > 1 else "single GPU",
                       "l2": "inputs (compressed + inflated) far larger than the 126 MB L2; no flush needed",
                       "generator_seconds": round(gen_s, 1), "k3_chain_repairs": chain_repairs},
            "records_per_s": n_rec_all / (ms_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_with_columns": e2e_all, "e2e_columns_only": e2e_cols, "pcie_pinned_copy": pcie, "sort": sort_info, "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if keep is not None:
        keep.close()


def run_e2e(a, bp, arr, c0, c1, rank, world, dist, torch, u_all, flags=None, api=None):
    """Reader::new(host bytes) -> read_header -> next_batch until EOF, with the batch's host mirrors filled as `flags` say;
    default: bodies + (rec_off, rec_len), i.e. exactly what records().next_rec() serves from."""
    if flags is None:
        flags = bp.COPY_DATA | bp.COPY_OFFSETS
    per_rec = 74 if flags & bp.COPY_COLUMNS else (12 if flags & bp.COPY_OFFSETS else 0)   # bytes of columns D2H per record
    full = bool(flags & bp.COPY_DATA)
    times, d2h = [], 0
    warm = 2   # the first Readers of a process page-lock their buffers (parked in the library's cache afterwards)
    for it in range(warm + a.e2e_steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2h = 0
        with bp.Reader.new(arr, min(os.cpu_count() or 4, 20), flags=flags, batch_bytes=a.batch_mib << 20) as r:   # main.rs:84 thread count
            r.read_header()
            n = 0
            while True:
                b = r.next_batch()
                if b is None:
                    break
                n += b.n_records
                d2h += (b.data_len if full else 0) + per_rec * b.n_records
            st = r.stats()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    t = float(np.mean(times))
    return {"value": u_all / t / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(c1 - c0), "d2h_bytes_per_step": int(d2h),
            "records_per_s": n / t, "ms_per_step": t * 1e3, "steps": a.e2e_steps, "warmup": warm,
            "device_ms": {k: round(st[k], 1) for k in ("ms_inflate", "ms_crc", "ms_records", "ms_d2h")}, "batches": st["batches"],
            "step_ms": [round(x * 1e3, 1) for x in times],
            "api": api or "Reader.new(host bytes, min(cpus, 20)) / read_header / next_batch: every record body + (offset, length) in pinned "
                          "host memory = what records().next_rec() lends (the reference hands out bodies, nothing else)"}


def run_e2e_sharded(a, bp, eng, shard, d_comp, step_host, sync_all, rank, world, dist, torch, u_all, my_u, n_rec_local):
    """N > 1: every rank moves its shard through pinned host memory (H2D compressed bytes, D2H bodies + columns) around
    the sharded decode; time = max over ranks."""
    L = bp.load()
    stage_bytes = 256 << 20
    h_shard = L.bamgpu_pinned_alloc(max(shard.size, 1))
    h_stage = L.bamgpu_pinned_alloc(stage_bytes)
    if not h_shard or not h_stage:
        return None
    C.memmove(h_shard, shard.ctypes.data, shard.size)
    times = []
    for it in range(1 + a.e2e_steps):
        sync_all()
        t0 = time.perf_counter()
        step_host(h_shard, h_stage, stage_bytes)
        sync_all()
        times.append(time.perf_counter() - t0)
    L.bamgpu_pinned_free(h_shard)
    L.bamgpu_pinned_free(h_stage)
    t = torch.tensor([float(np.mean(times[1:]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    io = torch.tensor([float(shard.size), float(my_u + 74 * n_rec_local), float(n_rec_local)], device="cuda", dtype=torch.float64)
    dist.all_reduce(io, op=dist.ReduceOp.SUM)
    t = float(t.item())
    return {"value": u_all / t / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
            "records_per_s": io[2].item() / t, "ms_per_step": t * 1e3, "steps": a.e2e_steps,
            "api": "per rank: pinned H2D of its block range, bamgpu_engine_inflate / append_halo / records, bodies + SoA columns D2H"}


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
