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

With N > 1 (torchrun, one rank per GPU) the file is sharded by BGZF block ranges (same file => strong
scaling); rank g sends the bytes in front of its first record start to rank g-1 (point-to-point), which
completes the record straddling the shard edge.  No collective sits on the data path.
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
    ap.add_argument("--e2e-steps", type=int, default=2)
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
        bam = synth.generate(a.shape, n_records, level=a.level)
        return bam.array, bam, time.time() - t0
    import torch.distributed as dist

    path = f"/dev/shm/bamgpu_bench_{os.environ.get('MASTER_PORT', '0')}_{n_records}.bam"
    t0 = time.time()
    if rank == 0:
        bam = synth.generate(a.shape, n_records, level=a.level)
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


def shard_blocks(bv: np.ndarray, world: int):
    """Cuts the block list into `world` contiguous ranges of ~equal compressed bytes. Returns [(b0, b1)]."""
    nb = bv.size
    in_off = bv["in_off"]
    total = int(in_off[-1]) if nb else 0
    cuts = [0]
    for g in range(1, world):
        cuts.append(int(np.searchsorted(in_off, np.uint64(total * g // world))))
    cuts.append(nb)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


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
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
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

    bv = blocks_view(blocks, nb)
    ranges = shard_blocks(bv, world)
    b0, b1 = ranges[rank]
    my_nb = b1 - b0
    # compressed byte range of my shard (whole BGZF blocks), rebased block table
    c0 = int(bv["in_off"][b0]) - 18 if my_nb else 0
    c1 = int(bv["in_off"][b1 - 1]) + int(bv["in_len"][b1 - 1]) + 8 if my_nb else 0
    u0 = int(bv["out_off"][b0]) if my_nb else 0
    my_blocks = (_lib.Block * max(my_nb, 1))()
    mv = blocks_view(my_blocks, my_nb)
    mv[:] = bv[b0:b1]
    mv["in_off"] -= np.uint64(c0)
    mv["out_off"] -= np.uint64(u0)
    my_u = int(mv["isize"].sum(dtype=np.uint64))
    my_c_payload = int(mv["in_len"].sum(dtype=np.uint64))
    eng = bp.Engine(device=local_rank)
    eng.set_n_ref(n_ref)
    shard = np.ascontiguousarray(arr[c0:c1])
    d_comp = eng.upload(shard)
    my_start = stream_start if rank == 0 else 0

    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    def step():
        # Device-resident pass: K1+K2 over the shard's blocks, K3 over its inflated bytes.
        # TODO(multi-GPU): shard-edge exchange; until then ranks > 0 decode from their first plausible record.
        batch, tail, rc = eng.decode(d_comp, my_blocks, my_nb, my_start, rank == world - 1, 0, sptr)
        return batch

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        batch = step()
    n_rec_local = batch.n_records
    chain_repairs = batch.chain_repairs
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
    roofline = {"bound": "hbm", "kernel": "k1_inflate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
                "stage_ms_per_step": {"k1_inflate": k1_ms, "k2_crc32": st["ms_crc"] / a.steps, "k3_records": st["ms_records"] / a.steps}}

    # ---- e2e: drop-in Reader over the HOST buffer (H2D + D2H inside the timed region) ----
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, bp, arr, c0, c1, rank, world, dist, torch, u_all)

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
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "records": n_rec_all, "inflated_bytes": u_all, "compressed_bytes": int(consumed),
                       "bgzf_blocks": nb, "sharding": f"{world} contiguous BGZF block ranges" if world > 1 else "single GPU",
                       "l2": "inputs (compressed + inflated) far larger than the 126 MB L2; no flush needed",
                       "generator_seconds": round(gen_s, 1), "k3_chain_repairs": chain_repairs},
            "records_per_s": n_rec_all / (ms_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if keep is not None:
        keep.close()


def run_e2e(a, bp, arr, c0, c1, rank, world, dist, torch, u_all):
    """Reader::new(host bytes) -> read_header -> next_batch until EOF, bodies + columns copied back to pinned
    host memory (what records().next_rec() serves from)."""
    if world > 1:
        return None  # TODO(multi-GPU e2e)
    flags = bp.COPY_DATA | bp.COPY_COLUMNS
    times, d2h = [], 0
    for it in range(1 + a.e2e_steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2h = 0
        with bp.Reader.new(arr, 4, flags=flags, batch_bytes=a.batch_mib << 20) as r:
            r.read_header()
            n = 0
            while True:
                b = r.next_batch()
                if b is None:
                    break
                n += b.n_records
                d2h += b.data_len + 74 * b.n_records
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it:
            times.append(dt)
    t = float(np.mean(times))
    return {"value": u_all / t / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(c1 - c0), "d2h_bytes_per_step": int(d2h),
            "records_per_s": n / t, "ms_per_step": t * 1e3, "steps": a.e2e_steps,
            "api": "Reader.new(host bytes) / read_header / next_batch, bodies + SoA columns D2H to pinned memory"}


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
