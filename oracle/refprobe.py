"""TEST INFRASTRUCTURE — thin driver for oracle/_ref/ref_probe* (see oracle/ref_probe.cpp).

Only tests/, oracle/make_golden.py and bench.py's reference arm may import this module.  It runs
the *reference's own classes* (compiled from /root/reference/src by oracle/Makefile) on inputs we
choose and returns their intermediates as numpy arrays.
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def probe_path(fp64=False, trellis=False):
    name = "ref_probe" + ("64" if fp64 else "") + ("_t" if trellis else "")
    return os.path.join(REF_DIR, name)


def available(fp64=False, trellis=False):
    return os.path.exists(probe_path(fp64, trellis))


def _fmt(v):
    return " ".join(repr(float(x)) for x in np.asarray(v, dtype=np.float64).ravel())


def run(mode, data, fp64=False, trellis=False, keep_dir=None, **kw):
    """Run one probe job. `data` is a 1-D array of float32-representable values.

    Returns a dict name -> numpy array (float64 or int64) plus 'files' -> {name: text} for the
    Records outputs of sweep mode.
    """
    exe = probe_path(fp64, trellis)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe + " not built (make -C oracle ref)")
    tmp = keep_dir or tempfile.mkdtemp(prefix="refprobe_")
    os.makedirs(tmp, exist_ok=True)
    if kw.pop("raw32", False):
        # large inputs (the bench arm): raw float32 on disk, the probe renders it as text piece by piece
        np.ascontiguousarray(np.asarray(data, dtype=np.float32)).tofile(os.path.join(tmp, "data.f32"))
        lines = ["mode " + mode, "data32 " + os.path.join(tmp, "data.f32"), "out " + tmp, "quiet 1"]
    else:
        data = np.ascontiguousarray(np.asarray(data, dtype=np.float32).astype(np.float64))
        data.tofile(os.path.join(tmp, "data.f64"))
        lines = ["mode " + mode, "data " + os.path.join(tmp, "data.f64"), "out " + tmp]
    for k, v in kw.items():
        if v is None:
            continue
        if isinstance(v, str):
            lines.append(f"{k} {v}")
        elif isinstance(v, (bool, int, np.integer)):
            lines.append(f"{k} {int(v)}")
        else:
            lines.append(f"{k} {_fmt(v)}")
    job = os.path.join(tmp, "job.txt")
    with open(job, "w") as f:
        f.write("\n".join(lines) + "\n")
    p = subprocess.run([exe, job], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"ref_probe failed ({p.returncode}): {p.stderr.strip()}")
    out = {"stdout": p.stdout, "files": {}}
    for fn in sorted(os.listdir(tmp)):
        path = os.path.join(tmp, fn)
        if fn.endswith(".f64"):
            if fn == "data.f64":
                continue
            out[fn[:-4]] = np.fromfile(path, dtype=np.float64)
        elif fn.endswith(".i64"):
            out[fn[:-4]] = np.fromfile(path, dtype=np.int64)
        elif fn.startswith("rec-") and fn.endswith(".csv"):
            with open(path) as f:
                out["files"][fn[4:-4]] = f.read()
    if keep_dir is None:
        for fn in os.listdir(tmp):
            os.unlink(os.path.join(tmp, fn))
        os.rmdir(tmp)
    return out
