"""`bam_binary` over the CUDA path: the reference's command line (bam_binary/src/main.rs:16-99), same flags, same outputs.

    python -m bam_parallel_b200.cli -h -i in.bam              # md5 of all record bodies       (main.rs:83-99)
    python -m bam_parallel_b200.cli -i in.bam -o out.bin      # coordinate + strand sort       (main.rs:50-79)
    python -m bam_parallel_b200.cli -n -i in.bam -o out.bin   # name sort
    python -m bam_parallel_b200.cli -n -M -i in.bam -o out.bin  # name sort, mates matched (HI, flag)
    python -m bam_parallel_b200.cli -i in.bam -o sorted.bam --output-format bam --compress-level 2 --devices 0,1   # (not in the reference)

`-h` is calc-hash like in the reference (structopt's help is shadowed there too, main.rs:36-37); `-q` is accepted and
ignored (`_quite_mode`, main.rs:19-20).  This is the parity harness for the reference's own end-to-end suites
(test.py, test_suites.py, tests_list.py): `--reference-suites DIR` runs them against the two BAM files they download
(`match_mates.bam`, `test_parallel_bam_reader.bam`) if those are present in DIR — there is no network on the build and
GPU boxes, so the files cannot be fetched here and the suites report "missing" instead.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys

import numpy as np

from . import COPY_DATA, COPY_OFFSETS, Reader
from .sorting import SortBy, sort_bam

MEGA_BYTE_SIZE = 1 << 20

# tests_list.py:6-48 ("generated … from files produced by sambamba-0.8.0"): (suite flags, file) -> md5
REFERENCE_MD5 = {
    (("-n", "-M"), "match_mates.bam"): "dfc84d497d016ee477771ba35f56e1b6",
    (("-n", "-M"), "test_parallel_bam_reader.bam"): "f314004468f708c3cf989f5ccb55e827",
    (("-n",), "match_mates.bam"): "2a770947df4addd3be9281b0e922cf88",
    ((), "match_mates.bam"): "76bfa67e2092b500e122876bf7346e8c",
    ((), "test_parallel_bam_reader.bam"): "563f30a5ecb85f5a3f77983072ec83bf",
    (("-h",), "test_parallel_bam_reader.bam"): "17c8bc5770e48dde225fd804c4bad013",
    (("-h",), "match_mates.bam"): "390cdd7a23488217351161daa9f59736",
}


def generate_file_hash(reader) -> str:
    """main.rs:83-99: Reader::new(reader, min(num_cpus, 20), None); read_header; md5 over every record body."""
    h = hashlib.md5()
    with Reader.new(reader, min(os.cpu_count() or 1, 20), None, flags=COPY_DATA | COPY_OFFSETS) as r:
        r.read_header()
        while True:
            b = r.next_batch()            # records().next_rec() a batch at a time: same bodies, same order
            if b is None:
                break
            off = b.columns["rec_off"].astype(np.int64)
            ln = b.columns["rec_len"].astype(np.int64)
            lo, hi = int(off[0]), int(off[-1] + ln[-1])
            keep = np.ones(hi - lo, dtype=bool)                  # drop the 4-byte block_size in front of every later body
            gaps = (off[1:] - lo)[:, None] - 4 + np.arange(4)[None, :]
            keep[gaps.ravel()] = False
            assert np.array_equal(off[1:], off[:-1] + ln[:-1] + 4), "records are not back to back"
            h.update(b.data[lo:hi][keep].tobytes())
    return h.hexdigest()


def parse(argv):
    ap = argparse.ArgumentParser(prog="BAM parallel CLI", add_help=False)
    ap.add_argument("-q", dest="quiet", action="store_true")
    ap.add_argument("-i", dest="input")
    ap.add_argument("-o", dest="output")
    ap.add_argument("-n", "--sort-by-name", dest="sort_by_name", action="store_true")
    ap.add_argument("-M", "--match-mates", dest="match_mates", action="store_true")
    ap.add_argument("-h", "--calc-hash", dest="calc_hash", action="store_true")
    ap.add_argument("--reference-suites", dest="suites", metavar="DIR")
    # beyond the reference (it dumps bodies, sort.rs:643): what the sort writes, how hard it compresses, on which GPUs
    ap.add_argument("--output-format", dest="fmt", choices=("bodies", "records", "bam"), default="bodies")
    ap.add_argument("--compress-level", dest="level", type=int, default=0,
                    help="with --output-format bam: 0 stored BGZF blocks, 1 fixed Huffman, 2 per-block Huffman codes (sort.rs:143 _out_compr_level)")
    ap.add_argument("--devices", dest="devices", default="", help="comma-separated GPU ordinals: the sort runs as a sample sort over them")
    return ap.parse_args(argv)


def sort_by_of(opt) -> SortBy:
    if opt.match_mates and not opt.sort_by_name:
        raise SystemExit("Cannot match mates when sorting by coordinates.")      # main.rs:51-53 panic!
    if opt.sort_by_name:
        return SortBy.NameAndMatchMates if opt.match_mates else SortBy.Name
    return SortBy.CoordinatesAndStrand


def run(opt, out=sys.stdout) -> int:
    if opt.calc_hash:
        with open(opt.input, "rb") as f:
            print(generate_file_hash(f), file=out)
        return 0
    sort_by = sort_by_of(opt)
    with open(opt.input, "rb") as f, open(opt.output, "wb") as o:               # main.rs:55 output.unwrap()
        devices = [int(x) for x in opt.devices.split(",") if x != ""] or None
        sort_bam(2000 * MEGA_BYTE_SIZE, f, o, None, opt.level, 5, None, None, sort_by, None, output=opt.fmt, devices=devices)   # main.rs:66-77
    return 0


def run_reference_suites(directory: str, out=sys.stdout) -> int:
    import io
    import tempfile
    bad = 0
    for (flags, name), want in REFERENCE_MD5.items():
        path = os.path.join(directory, name)
        label = " ".join(flags) or "(coordinate sort)"
        if not os.path.exists(path):
            print(f"missing   {label:18s} {name}", file=out)
            continue
        if flags == ("-h",):
            buf = io.StringIO()
            run(parse(["-h", "-i", path]), buf)
            got = buf.getvalue().strip()
        else:
            with tempfile.NamedTemporaryFile() as t:
                run(parse(list(flags) + ["-i", path, "-o", t.name]))
                got = hashlib.md5(open(t.name, "rb").read()).hexdigest()
        ok = got == want
        bad += not ok
        print(f"{'ok' if ok else 'MISMATCH':9s} {label:18s} {name} {got}", file=out)
    return bad


def main(argv=None) -> int:
    opt = parse(sys.argv[1:] if argv is None else argv)
    if opt.suites:
        return run_reference_suites(opt.suites)
    if not opt.input:
        raise SystemExit("-i <input> is required")
    return run(opt)


if __name__ == "__main__":
    sys.exit(main())
