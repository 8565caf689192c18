"""Segment-split mode on real GPUs: one process per GPU under torchrun, checked against a
single handle holding the whole sequence (tests/mgpu_worker.py).  Needs >= 2 GPUs; the 1-GPU box skips it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,transport", [(2, "peer"), (2, "nccl"), (4, "peer"), (8, "peer")])
def test_segment_split_matches_single_handle(world, transport):
    """transport: the per-sweep carries travel through peer mailboxes over NVLink (default) or NCCL all-gathers."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    env = dict(os.environ, HML_EXCHANGE=transport, HML_EXPECT_TRANSPORT=transport)
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "MGPU WORKER OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{path} not built")
    return path


@pytest.mark.gpu
@pytest.mark.parametrize("outputs", [["M"], ["M", "S", "P", "C", "G"], ["M", "S", "P", "B", "C", "G"]], ids=lambda o: "".join(o))
@pytest.mark.parametrize("world", [2, 4, 8])
def test_cli_devices_replay_run_equals_the_reference(tmp_path, world, outputs):
    """`hammlet64 -devices 0 1 ... -replay`: the sequence is split over `world` GPUs (one forked process each), process 0
    writes the files.  Every file must equal the real_t = double reference's byte for byte — the state marginals
    (accumulated per rank on the devices and merged at save time, or on the host from the gathered runs), the sequences
    (runs that cross a rank border are one entry), blocks, compression, parameters, segments."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import numpy as np
    from hammlet_b200.synth import piecewise_gaussian
    ours = _need(os.path.join(ROOT, "hammlet_b200", "bin", "hammlet64"))
    ref = _need(os.path.join(ROOT, "oracle", "_ref", "hammlet64"))
    x = piecewise_gaussian(90000, 3, 400, 5, quantum_bits=10)
    with open(tmp_path / "in.txt", "w") as f:
        f.write("\n".join(f"{v:.10f}" for v in x.astype(np.float64)) + "\n")
    common = ["-f", "in.txt", "-a", "-R", "5", "-s", "3", "-i", "M", "10", "0", "S", "P", "F", "20", "2", "D", "F", "20", "1",
              "-O"] + outputs + ["-w"]
    r = subprocess.run([ref] + common + ["-o", "ref-", ".csv"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    p = subprocess.run([ours, "-replay", "-devices"] + [str(d) for d in range(world)] + common + ["-o", "our-", ".csv"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0 and p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:] + r.stderr[-2000:]
    kinds = {"M": "marginals", "S": "sequences", "P": "parameters", "B": "blocks", "C": "compression", "G": "segments"}
    for o in outputs:
        assert (tmp_path / f"our-{kinds[o]}.csv").read_text() == (tmp_path / f"ref-{kinds[o]}.csv").read_text(), kinds[o]


@pytest.mark.gpu
def test_cli_devices_philox_run_equals_single_device(tmp_path):
    """Without -replay the uniforms are Philox counters indexed by the global block number, so a run split over two
    GPUs must write the same marginals file as the same run on one GPU.  Both runs draw the parameters on the host
    (`HAMMLET_HOST_PARAMS=1`): left to itself the single-device run would use the device-resident chain, whose parameter
    draws are Philox streams and not the host's mt19937."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    from hammlet_b200.synth import piecewise_gaussian
    ours = _need(os.path.join(ROOT, "hammlet_b200", "bin", "hammlet"))
    x = piecewise_gaussian(400000, 4, 800, 9, quantum_bits=10)
    with open(tmp_path / "in.txt", "w") as f:
        f.write("\n".join(f"{v:.10f}" for v in x.astype(np.float64)) + "\n")
    common = ["-f", "in.txt", "-a", "-R", "9", "-s", "4", "-i", "F", "60", "3", "-O", "M", "P", "C", "-w"]
    env = dict(os.environ, HAMMLET_HOST_PARAMS="1")
    a = subprocess.run([ours] + common + ["-o", "one-", ".csv"], cwd=tmp_path, capture_output=True, text=True, timeout=600,
                       env=env)
    b = subprocess.run([ours, "-devices", "0,1"] + common + ["-o", "two-", ".csv"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=150, env=env)
    assert a.returncode == 0 and b.returncode == 0, a.stderr[-2000:] + b.stderr[-2000:]
    for kind in ("marginals", "compression", "parameters"):
        assert (tmp_path / f"one-{kind}.csv").read_text() == (tmp_path / f"two-{kind}.csv").read_text(), kind


@pytest.mark.gpu
def test_cli_devices_multivariate_replay_run_equals_the_reference(tmp_path):
    """`-s C 2 2` (two values per position, four states) on a sequence split over two GPUs: all six files equal the
    real_t = double reference's byte for byte."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    from hammlet_b200.synth import piecewise_gaussian_md
    ours = _need(os.path.join(ROOT, "hammlet_b200", "bin", "hammlet64"))
    ref = _need(os.path.join(ROOT, "oracle", "_ref", "hammlet64"))
    x = piecewise_gaussian_md(50000, 2, 2, 250, 7, quantum_bits=10)
    with open(tmp_path / "in.txt", "w") as f:
        f.write("\n".join(" ".join(f"{v:.10f}" for v in row) for row in x.astype(np.float64)) + "\n")
    common = ["-f", "in.txt", "-a", "-R", "7", "-s", "C", "2", "2", "-i", "M", "5", "0", "F", "30", "2",
              "-O", "M", "S", "P", "B", "C", "G", "-w"]
    r = subprocess.run([ref] + common + ["-o", "ref-", ".csv"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    p = subprocess.run([ours, "-replay", "-devices", "0", "1"] + common + ["-o", "our-", ".csv"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=150)
    assert r.returncode == 0 and p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:] + r.stderr[-2000:]
    for kind in ("blocks", "compression", "sequences", "parameters", "marginals", "segments"):
        assert (tmp_path / f"our-{kind}.csv").read_text() == (tmp_path / f"ref-{kind}.csv").read_text(), kind
