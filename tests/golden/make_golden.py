"""Regenerates tests/golden/selfgolden.json.

These are SELF-goldens: md5s computed by the oracle over seeded generator output, kept so
that generator / oracle drift between rounds is caught.  They are NOT reference-published
values (the reference's own md5s, tests_list.py:6-48, need two network files).
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bam_parallel_b200 import synth  # noqa: E402
from oracle.oracle import COracle  # noqa: E402


def bodies_md5(u, off, ln, order=None):
    h = hashlib.md5()
    ub = bytes(u)
    idx = range(len(off)) if order is None else order
    off, ln = off.tolist(), ln.tolist()
    for i in idx:
        h.update(ub[off[i]:off[i] + ln[i]])
    return h.hexdigest()


def main():
    o = COracle()
    out = {}
    for shape, n, level, payload in [("pe150", 5000, 6, 0xFF00), ("se150", 5000, 6, 0xFF00), ("longname", 3000, 1, 0xFF00),
                                     ("long10k", 100, 9, 4096)]:
        bam = synth.generate(shape, n, level=level, payload=payload, threads=1)
        u, _ = o.inflate_all(bam.tobytes())
        _, _, consumed = o.read_header(u)
        off, ln = o.walk_records(u, consumed)
        cols, _, _ = o.extract(u, off, ln, True)
        tag = f"{shape}_{n}_L{level}" + ("" if payload == 0xFF00 else f"_p{payload}")
        out[tag + "_md5_bodies"] = bodies_md5(u, off, ln)                      # main.rs:83-99 (-h)
        for mode, nm in ((0, "coord"), (1, "name"), (2, "name_mates")):       # sort.rs:643 body-only dump
            out[tag + f"_md5_sorted_{nm}"] = bodies_md5(u, off, ln, o.sort_perm(u, off, cols, mode).tolist())
        out[tag + "_md5_file"] = hashlib.md5(bam.tobytes()).hexdigest()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "selfgolden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
