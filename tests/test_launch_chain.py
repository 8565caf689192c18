"""Static check of the launch-chain invariant ("Launch chain"): every kernel that is launched with
programmatic stream serialization (launch_k) must execute pdl_enter() — griddepcontrol.wait — as its FIRST statement.
A kernel that touched memory before the wait, or let a CTA return without it, could run ahead of its predecessor in
the stream (and let the stream's next operation, e.g. the result copy, run ahead too)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_pdl_launched_kernel_waits_first():
    src = ""
    for f in sorted(glob.glob(os.path.join(ROOT, "hammlet_b200", "csrc", "*.cu*"))):
        src += open(f).read() + "\n"
    launched = sorted(set(re.findall(r"launch_k\(\s*(k_[a-z_0-9]+)", src)))
    assert len(launched) >= 20
    for name in launched:
        m = re.search(r"__global__[^;{}]*?\b" + name + r"\s*\([^;{}]*?\)\s*\{\s*\n\s*(.*)", src)
        assert m, f"definition of {name} not found"
        assert m.group(1).strip().startswith("pdl_enter();"), f"{name} is launched with launch_k but does not start with pdl_enter()"


def test_pdl_enter_waits_before_it_triggers():
    text = open(os.path.join(ROOT, "hammlet_b200", "csrc", "hml_common.cuh")).read()
    body = text[text.index("void pdl_enter()"):]
    body = body[:body.index("}")]
    assert body.index("griddepcontrol.wait") < body.index("griddepcontrol.launch_dependents")
