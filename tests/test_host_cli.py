"""Host-side (C++) surface: Parser / Records / StateMarginals logic and the `hammlet` command line.

CPU part: record files produced by our Records/StateMarginals from the reference's own sampled
iterations must equal the reference's files byte for byte (marginals = common refinement, sequences,
blocks, compression, segments incl. the reference's internal code length); command-line errors must
read like the reference's (main.cpp:467-474).  GPU part: full runs of bin/hammlet against the
reference binaries built from /root/reference (oracle/_ref/, they travel to the GPU box).
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "hammlet_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")


def need(path):
    if not os.path.exists(path):
        pytest.skip(f"{path} not built")
    return path


def run(cmd, stdin=None, cwd=None):
    return subprocess.run(cmd, input=stdin, capture_output=True, text=True, cwd=cwd)


@pytest.mark.parametrize("tag", ["32", "64"])
def test_records_files_match_reference(tmp_path, tag):
    tool = need(os.path.join(BIN, "records_tool"))
    g = np.load(os.path.join(GOLD, "fb_T20000_K3_dyn5.npz"), allow_pickle=False)
    T, K, nsw = int(g["T"]), int(g["K"]), int(g["nsweeps"])
    sizes = [[int(v) for v in line.split("\t")] for line in str(g["file_blocks" + tag]).strip().split("\n")]
    states = g["all_states" + tag]
    txt, off = [f"{T} {K} {nsw}"], 0
    for it in range(nsw):
        B = len(sizes[it])
        txt.append(str(B))
        txt.append(" ".join(map(str, sizes[it])))
        txt.append(" ".join(str(int(s)) for s in states[off:off + B]))
        off += B
    assert off == states.size
    p = run([tool, str(tmp_path / "out-"), ".csv"], stdin="\n".join(txt) + "\n")
    assert p.returncode == 0, p.stderr
    for kind in ("marginals", "sequences", "blocks", "compression", "segments"):
        assert (tmp_path / f"out-{kind}.csv").read_text() == str(g[f"file_{kind}{tag}"]), kind


@pytest.mark.parametrize("tag", ["32", "64"])
def test_records_from_runs_match_reference(tmp_path, tag):
    """Records::recordRun — recorded iterations delivered as equal-state runs, as the device hands them over —
    must produce the reference's marginals, sequences, compression and segments files byte for byte."""
    tool = need(os.path.join(BIN, "records_tool"))
    g = np.load(os.path.join(GOLD, "fb_T20000_K3_dyn5.npz"), allow_pickle=False)
    T, K, nsw = int(g["T"]), int(g["K"]), int(g["nsweeps"])
    sizes = [[int(v) for v in line.split("\t")] for line in str(g["file_blocks" + tag]).strip().split("\n")]
    states = g["all_states" + tag]
    txt, off = [f"{T} {K} {nsw}"], 0
    for it in range(nsw):
        B = len(sizes[it])
        txt += [str(B), " ".join(map(str, sizes[it])), " ".join(str(int(s)) for s in states[off:off + B])]
        off += B
    p = run([tool, str(tmp_path / "out-"), ".csv", "runs"], stdin="\n".join(txt) + "\n")
    assert p.returncode == 0, p.stderr
    for kind in ("marginals", "sequences", "compression", "segments"):
        assert (tmp_path / f"out-{kind}.csv").read_text() == str(g[f"file_{kind}{tag}"]), kind


def test_records_refuse_overwrite_and_overrun(tmp_path):
    tool = need(os.path.join(BIN, "records_tool"))
    p = run([tool, str(tmp_path / "o-"), ".csv"], stdin="10 2 1\n2\n6 5\n0 1\n")
    assert p.returncode == 1 and "Cannot record block, exceeding data size!" in p.stderr


BAD_COMMAND_LINES = [
    ["-x"],                                  # first token is not a flag
    ["-a", "-a"],                            # duplicate flag
    ["-s", "1", "-a"],                       # fewer than two states
    ["-s", "3"],                             # manual priors are not implemented
    ["-a", "-t", "abc"],                     # conversion failure
    ["-a", "-s", "Q", "3"],                  # unknown mapping type
]


@pytest.mark.parametrize("argv", BAD_COMMAND_LINES, ids=lambda a: " ".join(a))
def test_cli_errors_read_like_the_reference(argv, tmp_path):
    ours = need(os.path.join(BIN, "hammlet"))
    p = run([ours] + argv, stdin="", cwd=tmp_path)
    assert p.returncode == 1
    assert p.stderr.startswith("\n[ERROR] ") and p.stderr.endswith("Terminating HaMMLET. The rest is silence.\n")
    ref = os.path.join(REF, "hammlet")
    if os.path.exists(ref):
        r = run([ref] + argv, stdin="", cwd=tmp_path)
        assert (r.returncode, r.stderr) == (p.returncode, p.stderr)


def test_cli_help_and_arguments(tmp_path):
    ours = need(os.path.join(BIN, "hammlet"))
    p = run([ours, "-h"], cwd=tmp_path)
    assert p.returncode == 0 and "-iterations" in p.stdout
    ref = os.path.join(REF, "hammlet")
    if os.path.exists(ref):   # -g prints the parsed token groups; ours has three extra flags at the end
        a = run([ours, "-g", "-h", "-s", "C", "2", "1"], cwd=tmp_path).stdout.split("\n")
        b = run([ref, "-g", "-h", "-s", "C", "2", "1"], cwd=tmp_path).stdout.split("\n")
        assert a[:16] == b[:16]


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ours = need(os.path.join(BIN, "hammlet"))
    p = run([ours, "-a", "-R", "1"], stdin="1 2 3 4\n", cwd=tmp_path)
    assert p.returncode == 1 and "no CUDA device available" in p.stderr and "no CPU fallback" in p.stderr


# ------------------------------------------------------------------------------------------ GPU: full runs

def write_input(path, T, K, L, seed):
    from hammlet_b200.synth import piecewise_gaussian
    x = piecewise_gaussian(T, K, L, seed, quantum_bits=10)
    with open(path, "w") as f:
        f.write("\n".join(f"{v:.10f}" for v in x.astype(np.float64)) + "\n")
    return x


def read_marginals(path):
    rows = [list(map(int, line.split("\t"))) for line in open(path).read().strip().split("\n")]
    sizes = np.array([r[0] for r in rows])
    width = max(len(r) for r in rows) - 1
    counts = np.array([r[1:] + [0] * (width - len(r) + 1) for r in rows], dtype=np.float64)
    return sizes, counts


@pytest.mark.gpu
@pytest.mark.parametrize("outputs", [["M"], ["M", "C", "P"]], ids=lambda o: "".join(o))
@pytest.mark.parametrize("seed,scheme", [(5, "F 60 1"), (11, "M 20 0 S P F 30 2 D F 30 1"), (3, "M 30 1")])
def test_replay_run_with_marginals_on_the_device(tmp_path, seed, scheme, outputs):
    """With no per-iteration host output requested (-O M, the default, optionally with compression / parameters) the
    recorded iterations are merged into the state marginals on the device (hml_marginals_add) and the marginals file is
    written from one copy at the end: it must equal the reference's file byte for byte."""
    ours, ref = need(os.path.join(BIN, "hammlet64")), need(os.path.join(REF, "hammlet64"))
    write_input(tmp_path / "in.txt", 60000, 3, 300, seed)
    common = ["-f", "in.txt", "-a", "-R", str(seed), "-s", "3", "-i"] + scheme.split() + ["-O"] + outputs + ["-w"]
    r = run([ref] + common + ["-o", "ref-", ".csv"], cwd=tmp_path)
    p = run([ours, "-replay"] + common + ["-o", "our-", ".csv"], cwd=tmp_path)
    assert r.returncode == 0 and p.returncode == 0, p.stderr + r.stderr
    kinds = {"M": "marginals", "C": "compression", "P": "parameters"}
    for o in outputs:
        assert (tmp_path / f"our-{kinds[o]}.csv").read_text() == (tmp_path / f"ref-{kinds[o]}.csv").read_text(), kinds[o]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,scheme", [(5, "F 60 1"), (11, "M 20 0 S P F 30 2 D F 30 1")])
def test_replay_run_without_block_output_uses_device_runs(tmp_path, seed, scheme):
    """Without -O B the recorded iterations reach Records as equal-state runs formed on the device
    (Records::recordRun); marginals, sequences, compression, segments and parameters must still equal the
    reference's files byte for byte."""
    ours, ref = need(os.path.join(BIN, "hammlet64")), need(os.path.join(REF, "hammlet64"))
    write_input(tmp_path / "in.txt", 60000, 3, 300, seed)
    common = ["-f", "in.txt", "-a", "-R", str(seed), "-s", "3", "-i"] + scheme.split() + ["-O", "M", "S", "P", "C", "G", "-w"]
    r = run([ref] + common + ["-o", "ref-", ".csv"], cwd=tmp_path)
    p = run([ours, "-replay"] + common + ["-o", "our-", ".csv"], cwd=tmp_path)
    assert r.returncode == 0 and p.returncode == 0, p.stderr + r.stderr
    for kind in ("compression", "sequences", "parameters", "marginals", "segments"):
        assert (tmp_path / f"our-{kind}.csv").read_text() == (tmp_path / f"ref-{kind}.csv").read_text(), kind


@pytest.mark.gpu
@pytest.mark.parametrize("seed,scheme", [(5, "F 60 1"), (11, "M 20 0 S P F 30 2 D F 30 1")])
def test_replay_run_is_identical_to_the_double_reference(tmp_path, seed, scheme):
    """bin/hammlet64 -replay draws every uniform and every parameter from the shared mt19937 in the
    reference's order; with fp64 host parameters the whole run — all six output files — must equal the
    real_t = double build of the reference byte for byte."""
    ours, ref = need(os.path.join(BIN, "hammlet64")), need(os.path.join(REF, "hammlet64"))
    write_input(tmp_path / "in.txt", 60000, 3, 300, seed)
    common = ["-f", "in.txt", "-a", "-R", str(seed), "-s", "3", "-i"] + scheme.split() + ["-O", "M", "S", "P", "B", "C", "G", "-w"]
    r = run([ref] + common + ["-o", "ref-", ".csv"], cwd=tmp_path)
    p = run([ours, "-replay"] + common + ["-o", "our-", ".csv"], cwd=tmp_path)
    assert r.returncode == 0 and p.returncode == 0, p.stderr + r.stderr
    for kind in ("blocks", "compression", "sequences", "parameters", "marginals", "segments"):
        assert (tmp_path / f"our-{kind}.csv").read_text() == (tmp_path / f"ref-{kind}.csv").read_text(), kind
    # the reference's own post-processing tool runs unchanged on our marginals
    tool = os.path.join(REF, "maxSegmentation")
    if os.path.exists(tool):
        seg = run([tool, "-i", "our-marginals.csv"], cwd=tmp_path)
        assert seg.returncode == 0, seg.stderr
        rows = [line.split("\t") for line in seg.stdout.strip().split("\n")]
        assert sum(int(r[0]) for r in rows) == 60000 and all(0 <= int(r[1]) < 3 for r in rows)
        assert seg.stdout == run([tool, "-i", "ref-marginals.csv"], cwd=tmp_path).stdout


@pytest.mark.gpu
@pytest.mark.parametrize("P,D,seed,scheme", [(2, 2, 7, "F 40 1"), (3, 2, 9, "M 10 0 F 30 2"), (2, 3, 4, "F 20 1 S F 20 1")])
def test_multivariate_replay_run_is_identical_to_the_double_reference(tmp_path, P, D, seed, scheme):
    """`-s C P D`: D values per position, P shared emission parameters, P**D states.  As in the
    univariate case the -replay run of bin/hammlet64 must reproduce every output file of the real_t = double
    reference byte for byte (maxlet weights, mapped emission terms, per-parameter statistics, auto priors)."""
    from hammlet_b200.synth import piecewise_gaussian_md
    ours, ref = need(os.path.join(BIN, "hammlet64")), need(os.path.join(REF, "hammlet64"))
    T = 40000
    x = piecewise_gaussian_md(T, P, D, 250, seed, quantum_bits=10)
    with open(tmp_path / "in.txt", "w") as f:
        f.write("\n".join(" ".join(f"{v:.10f}" for v in row) for row in x.astype(np.float64)) + "\n")
    common = ["-f", "in.txt", "-a", "-R", str(seed), "-s", "C", str(P), str(D), "-i"] + scheme.split() + \
             ["-O", "M", "S", "P", "B", "C", "G", "-w"]
    r = run([ref] + common + ["-o", "ref-", ".csv"], cwd=tmp_path)
    p = run([ours, "-replay"] + common + ["-o", "our-", ".csv"], cwd=tmp_path)
    assert r.returncode == 0 and p.returncode == 0, p.stderr + r.stderr
    for kind in ("blocks", "compression", "sequences", "parameters", "marginals", "segments"):
        assert (tmp_path / f"our-{kind}.csv").read_text() == (tmp_path / f"ref-{kind}.csv").read_text(), kind
    sizes, counts = read_marginals(tmp_path / "our-marginals.csv")
    assert sizes.sum() == T and counts.shape[1] <= P ** D


@pytest.mark.gpu
def test_gzip_and_float32_inputs_give_the_text_run(tmp_path):
    """-f takes gzip'd text (recognised by its magic number: the *-count.csv.gz files of the reference's samToCounts)
    and, with -F f32, raw little-endian float32; both runs must write the files of the plain-text run."""
    import gzip
    ours = need(os.path.join(BIN, "hammlet"))
    x = write_input(tmp_path / "in.txt", 80000, 3, 400, 6)
    (tmp_path / "in.txt.gz").write_bytes(gzip.compress((tmp_path / "in.txt").read_bytes(), compresslevel=1))
    np.loadtxt(tmp_path / "in.txt", dtype=np.float32).tofile(tmp_path / "in.f32")
    assert x.size == 80000
    common = ["-a", "-R", "4", "-s", "3", "-i", "F", "40", "2", "-O", "M", "P", "C", "-w"]
    runs = {"txt": ["-f", "in.txt"], "gz": ["-f", "in.txt.gz"], "f32": ["-f", "in.f32", "-F", "f32"]}
    for tag, src in runs.items():
        p = run([ours] + src + common + ["-o", tag + "-", ".csv"], cwd=tmp_path)
        assert p.returncode == 0, p.stderr
    for kind in ("marginals", "parameters", "compression"):
        ref = (tmp_path / f"txt-{kind}.csv").read_text()
        assert (tmp_path / f"gz-{kind}.csv").read_text() == ref and (tmp_path / f"f32-{kind}.csv").read_text() == ref, kind
    p = run([ours, "-f", "in.txt", "-F", "bogus"] + common, cwd=tmp_path)
    assert p.returncode == 1 and "Unknown input format" in p.stderr


@pytest.mark.gpu
def test_multivariate_input_must_fill_all_dimensions(tmp_path):
    ours = need(os.path.join(BIN, "hammlet"))
    p = run([ours, "-a", "-s", "C", "2", "2", "-w"], stdin="1 2 3 4 5\n", cwd=tmp_path)
    assert p.returncode == 1
    assert "Input stream did not contain enough values to fill all dimensions at last position!" in p.stderr


@pytest.mark.gpu
def test_default_scheme_marginals_agree_in_distribution(tmp_path):
    """Philox uniforms, float host parameters, the reference's default sampling scheme: posterior state
    marginals agree with the float reference within a total-variation tolerance (different RNG streams);
    the tolerance is set against the reference's own seed-to-seed spread."""
    ours, ref = need(os.path.join(BIN, "hammlet")), need(os.path.join(REF, "hammlet"))
    T, K = 200000, 3
    write_input(tmp_path / "in.txt", T, K, 2000, 3)

    def marg(exe, seed, tag, extra=()):
        p = run([exe, "-f", "in.txt", "-a", "-R", str(seed), "-s", str(K), "-w", "-o", tag + "-", ".csv"] + list(extra), cwd=tmp_path)
        assert p.returncode == 0, p.stderr
        sizes, counts = read_marginals(tmp_path / f"{tag}-marginals.csv")
        assert sizes.sum() == T and np.all(counts.sum(1) == 100)     # 300 F sweeps, thinning 3
        # label switching: order columns by their weighted mean data level
        dense = np.repeat(counts / 100.0, sizes, axis=0)
        return dense

    x = np.loadtxt(tmp_path / "in.txt")

    def canon(d):
        lvl = [(d[:, k] * x).sum() / max(d[:, k].sum(), 1e-9) for k in range(d.shape[1])]
        d = d[:, np.argsort(lvl)]
        return np.pad(d, ((0, 0), (0, K - d.shape[1])))

    a, b = canon(marg(ref, 1, "r1")), canon(marg(ref, 2, "r2"))
    o = canon(marg(ours, 1, "o1"))
    tv_ref = 0.5 * np.abs(a - b).sum(1).mean()
    tv_ours = 0.5 * np.abs(a - o).sum(1).mean()
    assert tv_ours <= max(3 * tv_ref, 0.01), (tv_ours, tv_ref)
