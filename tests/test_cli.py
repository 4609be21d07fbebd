"""bam_binary's command line over the CUDA path: flag handling on the CPU, outputs on the GPU."""
import hashlib
import io
import os

import pytest

from bam_parallel_b200 import cli, synth
from bam_parallel_b200.sorting import SortBy


def test_flags_follow_main_rs():
    assert cli.sort_by_of(cli.parse(["-i", "x"])) == SortBy.CoordinatesAndStrand          # main.rs:50
    assert cli.sort_by_of(cli.parse(["-n", "-i", "x"])) == SortBy.Name
    assert cli.sort_by_of(cli.parse(["-n", "-M", "-i", "x", "-q"])) == SortBy.NameAndMatchMates
    with pytest.raises(SystemExit, match="Cannot match mates when sorting by coordinates"):   # main.rs:51-53
        cli.sort_by_of(cli.parse(["-M", "-i", "x"]))
    assert cli.parse(["-h", "-i", "x"]).calc_hash                                             # -h is calc-hash, not help
    assert len(cli.REFERENCE_MD5) == 7                                                        # tests_list.py:6-48
    # beyond the reference: output format, compression level (sort.rs:143 _out_compr_level), device list
    o = cli.parse(["-i", "x", "-o", "y"])
    assert (o.fmt, o.level, o.devices) == ("bodies", 0, "")                                   # defaults = what bam_binary writes
    o = cli.parse(["-n", "-i", "x", "-o", "y", "--output-format", "bam", "--compress-level", "2", "--devices", "0,1,2"])
    assert (o.fmt, o.level, o.devices) == ("bam", 2, "0,1,2")


def test_reference_suites_report_missing_files(tmp_path):
    out = io.StringIO()
    assert cli.run_reference_suites(str(tmp_path), out) == 0
    assert out.getvalue().count("missing") == 7


@pytest.mark.gpu
def test_hash_and_sort_modes_match_oracle(oracle, tmp_path):
    data = synth.generate("pe150", 60000, level=6).tobytes()
    path = tmp_path / "in.bam"
    path.write_bytes(data)
    u, _ = oracle.inflate_all(data)
    _, _, start = oracle.read_header(u)
    off, ln = oracle.walk_records(u, start)
    ub = u.tobytes()
    bodies = [ub[int(o):int(o) + int(l)] for o, l in zip(off, ln)]
    out = io.StringIO()
    assert cli.run(cli.parse(["-h", "-i", str(path)]), out) == 0
    assert out.getvalue().strip() == hashlib.md5(b"".join(bodies)).hexdigest()                 # main.rs:83-99
    cols, _, _ = oracle.extract(u, off, ln, True)
    for flags, mode in (([], 0), (["-n"], 1), (["-n", "-M"], 2)):
        o = tmp_path / "out.bin"
        assert cli.main(flags + ["-i", str(path), "-o", str(o)]) == 0
        perm = oracle.sort_perm(u, off, cols, mode)
        assert o.read_bytes() == b"".join(bodies[i] for i in perm.tolist()), flags
    # beyond the reference: a complete, compressed BAM, sorted over a device list; the CLI itself (hash mode) reads it back
    srt = tmp_path / "sorted.bam"
    assert cli.main(["-n", "-i", str(path), "-o", str(srt), "--output-format", "bam", "--compress-level", "2", "--devices", "0,0"]) == 0
    perm = oracle.sort_perm(u, off, cols, 1)
    assert srt.stat().st_size < len(ub) // 2
    out = io.StringIO()
    assert cli.run(cli.parse(["-h", "-i", str(srt)]), out) == 0
    assert out.getvalue().strip() == hashlib.md5(b"".join(bodies[i] for i in perm.tolist())).hexdigest()
