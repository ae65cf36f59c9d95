"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference (AgamChopra/TorchRegister) from
/root/reference so that (a) the restatements in this directory can be checked
against it and (b) golden vectors can be generated (tests/golden/make_golden.py).
/root/reference only exists in the build container, never on the GPU box:
nothing under tests -m gpu, smoke() or bench.py may call `load()`.

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

REF_ROOT = os.environ.get("TRB_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src", "TorchRegister")

captured_losses: list = []


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "torchregister.py"))


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
    if not available():
        raise RuntimeError("reference not present at %s" % REF_SRC)
    _install_stubs()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import torchregister as ref_api   # noqa: E402  (reference module names)
    import warpings as ref_warp       # noqa: E402
    import utils as ref_utils         # noqa: E402
    # silence tqdm bars
    ref_warp.trange = lambda n, *a, **k: range(n)
    return ref_api, ref_warp, ref_utils


def last_losses():
    """Per-epoch losses captured from the most recent debug=True run."""
    return captured_losses[-1] if captured_losses else None
