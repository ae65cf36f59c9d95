"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference (AgamChopra/TorchRegister) so that (a) the
restatements in this directory can be checked against it, (b) golden vectors can
be generated (tests/golden/make_golden.py) and (c) bench.py can time the
reference's own CPU path (`--impl reference`, `cpu_baseline.kind = "reference"`).
Where it is found, in this order: $TRB_REFERENCE_ROOT, /root/reference (build
container only) and baseline/_ref/ — the `pip install --target baseline/_ref`
copy that __graft_entry__.build() makes (git-ignored, travels to the GPU box).
The GPU parity tests and smoke() never call `load()`.

Shim: the reference imports matplotlib (absent here) and uses
absolute intra-package imports, so we stub matplotlib.pyplot and put
src/TorchRegister itself on sys.path.  No reference file is modified or copied.
With debug=True the reference calls plt.plot(losses_train) at the last epoch
(TR/warpings.py:96-97,162-163,223-224); the stub records that argument, which
hands the per-epoch loss list over without touching reference code.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.path.join(os.environ["TRB_REFERENCE_ROOT"], "src", "TorchRegister")] if os.environ.get("TRB_REFERENCE_ROOT") else []
_CANDIDATES += [os.path.join("/root/reference", "src", "TorchRegister"), os.path.join(_REPO, "baseline", "_ref", "TorchRegister")]


def _find():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "torchregister.py")):
            return c
    return None


REF_SRC = _find() or _CANDIDATES[0]

captured_losses: list = []


def available() -> bool:
    return _find() is not None


def source() -> str:
    """Directory the reference is imported from (for reports)."""
    return _find() or ""


def _install_stubs():
    if "matplotlib.pyplot" in sys.modules and getattr(sys.modules["matplotlib.pyplot"], "_trb_stub", False):
        return
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt._trb_stub = True

    def plot(*args, **kwargs):
        if args:
            captured_losses.append(list(args[0]))

    plt.plot = plot
    for name in ("title", "xlabel", "ylabel", "legend", "show", "figure", "close"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def load():
    """Return (torchregister, warpings, utils) modules of the reference."""
    src = _find()
    if src is None:
        raise RuntimeError("reference not present (looked in %s)" % ", ".join(_CANDIDATES))
    _install_stubs()
    if src not in sys.path:
        sys.path.insert(0, src)
    import torchregister as ref_api   # noqa: E402  (reference module names)
    import warpings as ref_warp       # noqa: E402
    import utils as ref_utils         # noqa: E402
    # silence tqdm bars
    ref_warp.trange = lambda n, *a, **k: range(n)
    return ref_api, ref_warp, ref_utils


def last_losses():
    """Per-epoch losses captured from the most recent debug=True run."""
    return captured_losses[-1] if captured_losses else None
