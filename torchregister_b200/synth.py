"""Synthetic registration inputs.

Band-limited, brain-MRI-like volumes built analytically from coordinates, so a
"target" can be produced by evaluating the same field at transformed
coordinates (no interpolation, no dependency on any warp implementation).
White-noise volumes are deliberately avoided: the fp32 noise floor of the
reference's own warp on them (1.2e-5) is above the 1e-5 parity budget.

Everything here is plain torch and device agnostic; tests generate on the CPU
(bit-reproducible) and copy to the GPU, bench.py generates on the device.
"""
from __future__ import annotations

import math
import torch

__all__ = ["blob_field", "make_pair", "rigid_theta_star", "affine_theta_star",
           "smooth_flow", "axis_coords"]


def axis_coords(size: int, device="cpu", dtype=torch.float64) -> torch.Tensor:
    """Voxel-centre coordinates in [-1, 1] (align_corners=False convention)."""
    i = torch.arange(size, device=device, dtype=dtype)
    return (2.0 * i + 1.0) / size - 1.0


def _blob_params(ndim: int, seed: int, n_blobs: int):
    g = torch.Generator(device="cpu").manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(n_blobs, generator=g, dtype=torch.float64)
    cen = torch.rand(n_blobs, ndim, generator=g, dtype=torch.float64) - 0.5
    wid = 0.1 + 0.4 * torch.rand(n_blobs, ndim, generator=g, dtype=torch.float64)
    return amp, cen, wid


def blob_field(coords, seed: int = 1234, n_blobs: int = 6) -> torch.Tensor:
    """Evaluate the synthetic intensity field at `coords`.

    coords: tuple of ndim broadcastable tensors ordered (x=W axis, y=H, [z=D]),
    normalised to [-1, 1].  Returns values in [0, ~1].
    """
    ndim = len(coords)
    amp, cen, wid = _blob_params(ndim, seed, n_blobs)
    dt, dev = coords[0].dtype, coords[0].device
    out = None
    for k in range(n_blobs):
        e = None
        for a in range(ndim):
            q = ((coords[a] - float(cen[k, a])) / float(wid[k, a])) ** 2
            e = q if e is None else e + q
        term = float(amp[k]) * torch.exp(-0.5 * e)
        out = term if out is None else out + term
    # ellipsoidal "skull" with a smooth edge
    radii = (0.85, 0.75, 0.8)[:ndim]
    r2 = None
    for a in range(ndim):
        q = (coords[a] / radii[a]) ** 2
        r2 = q if r2 is None else r2 + q
    mask = torch.sigmoid((1.0 - torch.sqrt(r2 + 1e-12)) * 12.0)
    out = out * mask
    return (out / (float(amp.sum()) * 0.6)).clamp(0.0, 1.0).to(dt).to(dev)


def _rot3(ax, ay, az):
    cx, sx = math.cos(ax), math.sin(ax)
    cy, sy = math.cos(ay), math.sin(ay)
    cz, sz = math.cos(az), math.sin(az)
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float64)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=torch.float64)
    return rz @ ry @ rx


def rigid_theta_star(ndim: int) -> torch.Tensor:
    """Ground-truth rigid transform: rotation (3,-2,4) deg + small translation."""
    d = math.pi / 180.0
    if ndim == 3:
        th = torch.zeros(3, 4, dtype=torch.float64)
        th[:, :3] = _rot3(3 * d, -2 * d, 4 * d)
        th[:, 3] = torch.tensor([0.03, -0.02, 0.01], dtype=torch.float64)
    else:
        a = 3 * d
        th = torch.tensor([[math.cos(a), -math.sin(a), 0.03],
                           [math.sin(a), math.cos(a), -0.02]], dtype=torch.float64)
    return th


def affine_theta_star(ndim: int) -> torch.Tensor:
    th = rigid_theta_star(ndim)
    sc = torch.tensor([1.03, 0.98, 1.01][:ndim], dtype=torch.float64)
    th[:, :ndim] = th[:, :ndim] * sc[None, :]
    return th


def smooth_flow(shape, amplitude: float = 3.0, device="cpu", dtype=torch.float32):
    """Low-frequency sinusoidal displacement field in voxel units.

    Returns [1, ndim, *shape]; channel i displaces spatial axis i (the
    SpatialTransformer convention, reference TR/utils.py:350-356).
    """
    ndim = len(shape)
    axes = [axis_coords(s, device, torch.float64) for s in shape]
    grids = torch.meshgrid(*axes, indexing="ij")
    chans = []
    for c in range(ndim):
        f = None
        for a in range(ndim):
            ph = 0.7 * c + 1.3 * a
            term = torch.sin(math.pi * (0.8 + 0.3 * ((a + c) % ndim)) * grids[a] + ph)
            f = term if f is None else f * term
        chans.append(amplitude * f)
    return torch.stack(chans, 0)[None].to(dtype)


def make_pair(shape, kind: str = "rigid", seed: int = 1234, noise: float = 1e-3,
              device="cpu", dtype=torch.float32):
    """Return (moving, target) of shape [1, 1, *shape].

    kind: 'rigid' | 'affine' — target is the field evaluated at theta*-mapped
          coordinates (the analytic equivalent of get_affine_warp(theta*, moving));
          'flow' — target is the field evaluated at coordinates displaced by
          `smooth_flow`;  'identity' — target = moving + noise.
    """
    ndim = len(shape)
    # spatial order of `shape` is ([D,] H, W); normalised coord order is (x, y[, z])
    axes = [axis_coords(s, device, torch.float64) for s in shape]
    grids = torch.meshgrid(*axes, indexing="ij")      # ordered ([D], H, W)
    base = tuple(reversed(grids))                      # (x, y[, z])
    moving = blob_field(base, seed)
    if kind in ("rigid", "affine"):
        th = (rigid_theta_star(ndim) if kind == "rigid" else affine_theta_star(ndim)).to(device)
        warped = []
        for r in range(ndim):
            acc = th[r, ndim].item() + 0.0 * base[0]
            for c in range(ndim):
                acc = acc + th[r, c].item() * base[c]
            warped.append(acc)
        target = blob_field(tuple(warped), seed)
    elif kind == "flow":
        fl = smooth_flow(shape, 3.0, device, torch.float64)[0]      # [ndim, *shape]
        disp = []
        for a in range(ndim):                                         # spatial axis a
            disp.append(grids[a] + fl[a] * (2.0 / shape[a]))
        target = blob_field(tuple(reversed(disp)), seed)
    elif kind == "identity":
        target = moving.clone()
    else:
        raise ValueError(kind)
    if noise:
        g = torch.Generator(device="cpu").manual_seed(4321 + seed)
        n = torch.randn(tuple(shape), generator=g, dtype=torch.float64).to(device)
        target = target + noise * n
    return (moving[None, None].to(dtype).contiguous(),
            target[None, None].to(dtype).contiguous())
