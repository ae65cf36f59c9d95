"""CPU: the C-ABI library loads and exports every symbol include/trb.h declares (no compute calls),
and the host-side logic mirrors the reference's conventions."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "trb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_header_symbols():
    from torchregister_b200 import _lib, build
    build.build_library()
    lib = ctypes.CDLL(build.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), "libtrb_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared, "ctypes signature table and header disagree"
    handle = _lib.load()
    assert handle.trb_abi_version() == 1
    assert handle.trb_affine_workspace_bytes(1) > 0 and handle.trb_flow_workspace_bytes() > 0


def test_kernels_are_sm100a():
    import shutil
    import subprocess
    from torchregister_b200 import build
    build.build_library()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    import torchregister_b200 as tr
    import torchregister_b200.functional as TF
    x = torch.zeros(1, 1, 8, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA only"):
        tr.Register(mode="rigid").optim(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TF.warp_affine(torch.eye(3, 4), x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TF.AffineProblem(x, x, "rigid", torch.zeros(6), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.SpatialTransformer((8, 8, 8))(x, torch.zeros(1, 3, 8, 8, 8))


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (or any CPU restatement)."""
    pkg = os.path.join(ROOT, "torchregister_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in text.replace("no oracle", ""), fn


def test_criterion_conventions():
    import torch.nn as nn
    from torchregister_b200.warpings import similarity_weights, _split_criteria
    from torchregister_b200.utils import NCCLoss
    # reference warpings.py:38-40,125-127: any user criterion -> MSE only
    assert similarity_weights([nn.L1Loss()], [0.2], "t") == (1.0, 0.0, 0.0)
    assert similarity_weights(None, [0.0, 1.0, 0.0], "t") == (0.0, 1.0, 0.0)
    assert similarity_weights(None, [0.33, 0.33, 0.33], "t") == (0.33, 0.33, 0.33)
    w_mse, w_ncc, other = _split_criteria([nn.MSELoss(), NCCLoss(alpha=50), nn.L1Loss()], [0.5, 0.4, 0.1])
    assert w_mse == 0.5 and abs(w_ncc - 0.2) < 1e-12 and len(other) == 1


def test_register_signature_matches_reference():
    import inspect
    import torchregister_b200 as tr
    sig = inspect.signature(tr.Register.__init__)
    assert list(sig.parameters)[1:7] == ["mode", "device", "criterion", "weight", "grad_edges", "debug"]
    extras = list(sig.parameters.values())[7:]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in extras)      # extensions never shift positions
    assert sig.parameters["flow_param"].default == "unet" and sig.parameters["optm"].default == "SGD"
    assert sig.parameters["mode"].default == "rigid" and sig.parameters["device"].default == "cpu"
    osig = inspect.signature(tr.Register.optim)
    names = list(osig.parameters)[1:7]
    assert names == ["moving", "target", "lr", "max_epochs", "n", "per"]
    assert osig.parameters["lr"].default == 1e-5 and osig.parameters["max_epochs"].default == 1000
    assert osig.parameters["n"].default == 32 and osig.parameters["per"].default == 0.1
    r = tr.Register()
    for attr in ("mode", "device", "criterion", "weight", "debug", "grad_edges", "theta", "warp"):
        assert hasattr(r, attr)
    assert r.warp is tr.get_affine_warp and tr.Register(mode="flow").warp is None


def test_unet_topology_and_state_dict_names():
    import torchregister_b200 as tr
    m2 = tr.Attention_UNet((160, 160), "bilinear", in_c=1, n=32)
    m3 = tr.Attention_UNet((160, 160, 160), "bilinear", in_c=1, n=32)
    assert sum(p.numel() for p in m2.parameters()) == 31278      # SURVEY.md §2.1 row 10 [probed]
    assert sum(p.numel() for p in m3.parameters()) == 89189
    keys = set(m2.state_dict())
    for k in ("layer1.0.weight", "layer1.3.bias", "layer5.6.weight", "skip4.input_filter.weight",
              "skip1.gate_filter.bias", "skip2.psi.weight", "out.bias"):
        assert k in keys
    with torch.no_grad():
        flow = m2.flow_field(torch.rand(1, 1, 160, 164), "cpu")
    assert tuple(flow.shape) == (1, 2, 160, 164)


def test_unet_matches_reference_when_available():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present (GPU box)")
    import warnings
    import torchregister_b200 as tr
    _, _, ru = ref_shim.load()
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ru.Attention_UNet((160, 172), "bilinear", in_c=1, n=32)
    mine = tr.Attention_UNet((160, 172), "bilinear", in_c=1, n=32)
    sd = {k: v for k, v in ref.state_dict().items() if k != "warp.grid"}
    mine.load_state_dict(sd, strict=True)
    x = torch.rand(1, 1, 160, 172)
    with torch.no_grad():
        _, ref_flow = ref(x, "cpu")
        flow = mine.flow_field(x, "cpu")
    assert torch.equal(flow, ref_flow)


def test_shard_planner():
    from torchregister_b200.parallel import shard_pairs, slab_range
    for n, w in ((64, 8), (64, 3), (5, 8), (512, 4), (192, 7)):
        spans = [shard_pairs(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert slab_range(512, 8, 3) == (192, 256)
    with pytest.raises(ValueError):
        shard_pairs(4, 2, 2)


def test_nmi_loss_matches_port():
    """NMILoss (blocked KDE, hand-written backward) against the oracle restatement of utils.py:18-79,224-259."""
    import torchregister_b200 as tr
    from oracle import torch_port as tp
    from torchregister_b200.synth import make_pair
    for shape, patch in (((48, 40), 100), ((14, 12, 10), 6)):
        mov, tgt = make_pair(shape, "rigid")
        w1 = (3.0 * mov + 0.1).clone().requires_grad_(True)          # range 3: the KDE is not degenerate
        w2 = w1.detach().clone().requires_grad_(True)
        a = tr.NMILoss(patch_size=patch)(3.0 * tgt, w1)
        b = tp.nmi_loss(3.0 * tgt, w2, patch=patch)
        a.backward(); b.backward()
        assert abs(a.item() - b.item()) <= 1e-3 * abs(b.item()) + 1e-4
        assert (w1.grad - w2.grad).abs().max() <= 2e-3 * w2.grad.abs().max()


def test_compose_theta_matches_chained_matrices():
    """EXTENSION f-2: compose_theta(first, second) is the homogeneous product T_first @ T_second (2-D and 3-D, batched)."""
    import torch
    from torchregister_b200.warpings import compose_theta
    g = torch.Generator().manual_seed(3)
    for nd in (2, 3):
        a = torch.eye(nd, nd + 1).repeat(4, 1, 1) + 0.1 * torch.randn(4, nd, nd + 1, generator=g)
        b = torch.eye(nd, nd + 1).repeat(4, 1, 1) + 0.1 * torch.randn(4, nd, nd + 1, generator=g)
        c = compose_theta(a, b)
        assert tuple(c.shape) == (4, nd, nd + 1)
        bottom = torch.zeros(4, 1, nd + 1); bottom[:, 0, nd] = 1
        ha, hb = torch.cat([a, bottom], 1), torch.cat([b, bottom], 1)
        assert torch.allclose(c, (ha @ hb)[:, :nd], atol=1e-6)
        x = torch.randn(4, nd, 1, generator=g)
        chained = a[:, :, :nd] @ (b[:, :, :nd] @ x + b[:, :, nd:]) + a[:, :, nd:]
        assert torch.allclose(c[:, :, :nd] @ x + c[:, :, nd:], chained, atol=1e-5)


def test_tile_fits_helper_runs_on_the_host():
    """trb_affine_tile_fits (kernel-variant hint) is a pure host function: identity and small rotations fit the staged box,
    the reference's own torch.rand(6) start (angles up to 1 rad, utils.py:317) does not."""
    import ctypes as C
    import math
    from torchregister_b200 import _lib
    lib = _lib.load()

    def rot_z(a):
        c, s = math.cos(a), math.sin(a)
        return (C.c_float * 12)(c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0)
    assert lib.trb_affine_tile_fits(192, 192, 160, rot_z(0.0)) == 1
    assert lib.trb_affine_tile_fits(192, 192, 160, rot_z(0.05)) == 1
    assert lib.trb_affine_tile_fits(192, 192, 160, rot_z(0.5)) == 0
    assert lib.trb_affine_tile_fits(256, 256, 256, rot_z(1.0)) == 0
    assert lib.trb_affine_tile_fits(0, 0, 0, rot_z(0.0)) == 0


def test_committed_bench_line_follows_the_contract():
    """The bench line committed at the end of the round (profiles/r02_bench_n1_final.json) carries every key of the driver's
    contract, the two roofline objects with consistent arithmetic, a reference-kind CPU baseline and a real end-to-end figure."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_bench_n1_final.json")
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "roofline_256", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    for r in (d["roofline"], d["roofline_256"]):
        assert r["bound"] == "hbm" and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert abs(r["achieved"] - r["algorithmic_bytes_per_epoch"] / (r["kernel_us_per_epoch"] * 1e-6) / 1e9) < 1e-6 * r["achieved"]
    assert d["roofline"]["traffic"] is None or d["roofline"]["traffic_per_epoch"] >= 0.99 * d["roofline"]["algorithmic_bytes_per_epoch"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert "rejected" not in d["clocks"] or d["clocks"]["rejected"] is False
