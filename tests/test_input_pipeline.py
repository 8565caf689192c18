"""Input pipeline: the multi-threaded parser of hammlet_b200/host/FastParse.hpp must give, for
every input, exactly what the reference's extraction loop `while (input >> v)` (wavelet.hpp:131) gives: the
same float bits, and the same stopping point at the first token std::num_get rejects."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "hammlet_b200", "bin", "parse_tool")


def parse(mode, path, out, threads=0):
    p = subprocess.run([TOOL, mode, str(threads), str(path), str(out)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    n, secs = p.stdout.split()
    return np.fromfile(out, dtype=np.float32), int(n), float(secs)


def both(tmp_path, text, threads=0):
    if not os.path.exists(TOOL):
        pytest.skip("parse_tool not built")
    src = tmp_path / "in.txt"
    src.write_text(text)
    fast, nf, _ = parse("fast", src, tmp_path / "fast.bin", threads)
    slow, ns, _ = parse("slow", src, tmp_path / "slow.bin")
    assert nf == fast.size and ns == slow.size
    assert fast.size == slow.size, (fast.size, slow.size, text[:80])
    assert np.array_equal(fast.view(np.uint32), slow.view(np.uint32))
    return fast


@pytest.mark.parametrize("text,count", [
    ("", 0), ("   \n\t ", 0), ("1", 1), ("1 2\n3\t4\r\n5", 5), ("+1.5 -2.25 .5 5. 1e3 1E-3 1.e2 -.5e+1", 8),
    ("1 2 abc 3", 2), ("1 2 3abc 4", 3), ("1 nan 2", 1), ("1 inf 2", 1), ("1 0x10 2", 2), ("1 1e 2", 1),
    ("1 1e+ 2", 1), ("1 - 2", 1), ("1 . 2", 1), ("1 1e50 2", 1), ("1 -1e50 2", 1), ("1e-50 2", 2),
    ("1e-45 1.4e-45 7e-46 3.4028235e38 3.4028236e38", 4), ("3.4028235677973366e38 1", 2), ("3.4028235677973367e38 1", 0),
    ("0.1 0.2 0.30000001192092896 16777217 16777216.5 16777217.5 1.00000005960464477539", 7),
    ("1,2", 1), ("1;2", 1), ("--1", 0), ("1..2", 2), ("1.2.3", 2), ("1e2e3", 1), ("1 2 +", 2),
])
def test_edge_cases_match_the_reference_loop(tmp_path, text, count):
    assert both(tmp_path, text).size == count


def test_random_numbers_and_piece_boundaries(tmp_path):
    rng = np.random.default_rng(3)
    n = 400_000   # > 1 MiB of text: several pieces
    vals = np.concatenate([rng.normal(0, 1, n // 2), rng.normal(0, 1e-30, n // 8), rng.normal(0, 1e30, n // 8),
                           rng.integers(-10**9, 10**9, n // 8).astype(np.float64), rng.random(n // 8) * 1e-42])
    rng.shuffle(vals)
    fmts = ["%.5f", "%.9g", "%.17g", "%e", "%+.3E", "%g"]
    seps = [" ", "\n", "\t", "  ", " \n"]
    text = "".join((fmts[i % len(fmts)] % v) + seps[(i * 7) % len(seps)] for i, v in enumerate(vals))
    for threads in (1, 3, 16):
        out = both(tmp_path, text, threads)
        assert out.size == n
    # a rejected token in the middle cuts the result at the same place whichever piece it falls into
    k = len(text) // 2
    k = text.index(" ", k)
    out = both(tmp_path, text[:k] + " oops " + text[k:], 8)
    assert 0 < out.size < n


def test_speed_report(tmp_path):
    """Not a pass/fail speed test: prints both timings so the log shows the gain."""
    if not os.path.exists(TOOL):
        pytest.skip("parse_tool not built")
    rng = np.random.default_rng(0)
    x = rng.normal(0, 1, 1_000_000)
    src = tmp_path / "big.txt"
    src.write_text("\n".join("%.5f" % v for v in x))
    f, _, tf = parse("fast", src, tmp_path / "f.bin")
    s, _, ts = parse("slow", src, tmp_path / "s.bin")
    assert np.array_equal(f.view(np.uint32), s.view(np.uint32))
    print(f"parse 1e6 values: fast {tf * 1e3:.1f} ms, reference loop {ts * 1e3:.1f} ms")


# ---- further input formats of the command line (-F auto|text|gz|f32): gzip'd text, which is what the reference's
# preprocessing writes (bin/samToCounts: *-count.csv.gz, one count per line; the reference reads it through `zcat |`),
# and raw little-endian float32.  Each must yield exactly the values the reference's loop extracts from the plain text.

def _reference_values(tmp_path, text):
    src = tmp_path / "plain.txt"
    src.write_text(text)
    slow, ns, _ = parse("slow", src, tmp_path / "slow.bin")
    assert ns == slow.size
    return slow


@pytest.mark.parametrize("members", [1, 3])
def test_gzip_text_equals_the_reference_loop_on_the_plain_text(tmp_path, members):
    import gzip
    if not os.path.exists(TOOL):
        pytest.skip("parse_tool not built")
    rng = np.random.default_rng(5)
    n = 2_500_000    # ~25 MB of text: several 8 MB pieces handed to the parser threads while inflation continues
    counts = rng.poisson(30, n)
    text = "\n".join(map(str, counts)) + "\n"      # samToCounts: one count per line
    ref = _reference_values(tmp_path, text)
    assert ref.size == n
    data = text.encode()
    step = (len(data) + members - 1) // members
    with open(tmp_path / "counts.csv.gz", "wb") as f:  # several gzip members in a row (`cat a.gz b.gz`) are one stream
        for i in range(members):
            f.write(gzip.compress(data[i * step:(i + 1) * step], compresslevel=1))
    for mode, threads in (("auto", 0), ("gz", 1), ("gz", 5)):
        got, ng, _ = parse(mode, tmp_path / "counts.csv.gz", tmp_path / "gz.bin", threads)
        assert ng == n and np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (mode, threads)


def test_gzip_text_stops_at_the_first_rejected_token(tmp_path):
    import gzip
    if not os.path.exists(TOOL):
        pytest.skip("parse_tool not built")
    text = " ".join(f"{v:.4f}" for v in np.random.default_rng(1).normal(size=1_500_000))
    k = text.index(" ", len(text) // 2)
    broken = text[:k] + " x12 " + text[k:]
    ref = _reference_values(tmp_path, broken)
    (tmp_path / "b.gz").write_bytes(gzip.compress(broken.encode(), compresslevel=1))
    got, _, _ = parse("auto", tmp_path / "b.gz", tmp_path / "b.bin")
    assert 0 < ref.size < 1_500_000 and np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_corrupt_gzip_is_an_error(tmp_path):
    import gzip
    if not os.path.exists(TOOL):
        pytest.skip("parse_tool not built")
    blob = gzip.compress(b"1 2 3 4 5 6 7 8 9 10\n" * 1000)
    (tmp_path / "t.gz").write_bytes(blob[:len(blob) // 2])
    p = subprocess.run([TOOL, "auto", "0", str(tmp_path / "t.gz"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "decompress" in p.stderr
    (tmp_path / "plain.txt").write_text("1 2 3")
    p = subprocess.run([TOOL, "gz", "0", str(tmp_path / "plain.txt"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "not in gzip format" in p.stderr


def test_raw_float32_round_trip(tmp_path):
    if not os.path.exists(TOOL):
        pytest.skip("parse_tool not built")
    x = np.random.default_rng(2).normal(size=1_000_003).astype(np.float32)
    x[:4] = [0.0, -0.0, np.float32(1e-45), np.float32(3.4028235e38)]
    x.tofile(tmp_path / "x.f32")
    got, n, _ = parse("f32", tmp_path / "x.f32", tmp_path / "o.bin")
    assert n == x.size and np.array_equal(got.view(np.uint32), x.view(np.uint32))
    # ... and equals what the text route gives for the same numbers written with 9 significant digits
    text = "\n".join("%.9g" % v for v in x[:200000])
    ref = _reference_values(tmp_path, text)
    assert np.array_equal(ref.view(np.uint32), x[:200000].view(np.uint32))
    (tmp_path / "odd.f32").write_bytes(b"\x00" * 7)
    p = subprocess.run([TOOL, "f32", "0", str(tmp_path / "odd.f32"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "whole number of 4-byte values" in p.stderr
