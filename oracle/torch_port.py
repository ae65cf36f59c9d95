"""TEST INFRASTRUCTURE — CPU oracle ("port"), not product code.

A restatement of the reference's registration hot path built from the SAME
third-party ops the reference composes (torch: F.affine_grid, F.grid_sample,
autograd, SGD), so that it runs anywhere torch runs (the reference itself only
exists in the build container).  The arithmetic of this path lives in PyTorch
(reference pins `torch>=2.0.0`, setup.py:11; this image: 2.11.0), not under
/root/reference.

Pinning: tests/golden/make_golden.py runs the UNMODIFIED reference (through
oracle/ref_shim.py) and this port on identical inputs and stores the
reference's outputs; tests/test_oracle_golden.py re-checks this port (and the
plain-C restatement oracle/c_oracle) against those stored reference outputs.
The reference owns no tests or golden vectors of its own.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.

Each function cites the reference lines it follows (paths relative to
/root/reference/src/TorchRegister/).
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

EPS_NCC = 1e-10        # utils.py:15


# --------------------------------------------------------------------------- #
# transform parametrisation
# --------------------------------------------------------------------------- #
def rigid_theta(p: torch.Tensor, max_translate: float = 0.25) -> torch.Tensor:
    """utils.py:287-310 (Theta.forward) + :324-330 (Regressor.forward).

    3-D: p = (psi, theta, phi, a, b, c) -> [1,3,4]; 2-D: p = (theta, tx, ty) -> [1,2,3].
    """
    if p.numel() > 3:
        psi, th, phi = p[0], p[1], p[2]
        cps, sps = torch.cos(psi), torch.sin(psi)
        cth, sth = torch.cos(th), torch.sin(th)
        cph, sph = torch.cos(phi), torch.sin(phi)
        rows = [cps * cth, sph * sps * cth - cph * sth, cph * sps * cth + sph * sth,
                max_translate * torch.tanh(p[3]),
                cps * sth, sph * sps * sth + cph * cth, cph * sps * sth - sph * cth,
                max_translate * torch.tanh(p[4]),
                -sps, sph * cps, cph * cps,
                max_translate * torch.tanh(p[5])]
        return torch.stack(rows).view(1, 3, 4)
    th = p[0]
    return torch.stack([torch.cos(th), -torch.sin(th), p[1],
                        torch.sin(th), torch.cos(th), p[2]]).view(1, 2, 3)


def identity_params(ndim: int, dtype=torch.float32) -> torch.Tensor:
    """warpings.py:47-48,54-55 — bias of the (otherwise inert) affine MLP."""
    if ndim == 3:
        return torch.tensor([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=dtype)
    return torch.tensor([1, 0, 0, 0, 1, 0], dtype=dtype)


# --------------------------------------------------------------------------- #
# warps
# --------------------------------------------------------------------------- #
def affine_warp(theta: torch.Tensor, moving: torch.Tensor) -> torch.Tensor:
    """warpings.py:18-26 (get_affine_warp)."""
    nd = moving.dim() - 2
    theta = theta.reshape(1, nd, nd + 1)
    grid = F.affine_grid(theta, list(moving.shape), align_corners=False)
    return F.grid_sample(moving, grid, mode="bilinear", padding_mode="zeros",
                         align_corners=False)


def flow_warp(src: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """utils.py:343-365 (SpatialTransformer).  flow in voxel units, channel i
    displaces spatial axis i; normalise, reverse channel order, sample with
    align_corners=True."""
    shape = flow.shape[2:]
    vecs = [torch.arange(0, s, dtype=torch.float32, device=flow.device) for s in shape]
    ident = torch.stack(torch.meshgrid(*vecs, indexing="ij"))[None].to(flow.dtype)
    locs = ident + flow
    comps = []
    for i, s in enumerate(shape):
        comps.append(2 * (locs[:, i] / (s - 1) - 0.5))
    comps.reverse()
    grid = torch.stack(comps, dim=-1)
    return F.grid_sample(src, grid, mode="bilinear", padding_mode="zeros",
                         align_corners=True)


# --------------------------------------------------------------------------- #
# similarity terms   (all called as crit(target, warped): warpings.py:78,144,213)
# --------------------------------------------------------------------------- #
def mse_loss(t, w):
    return torch.mean((t - w) ** 2)


def ncc_loss(t, w, alpha: float = 100.0):
    """utils.py:197-205 (global NCC, EPSILON inside the sqrt)."""
    a = t - torch.mean(t)
    b = w - torch.mean(w)
    ncc = torch.sum(a * b) / ((torch.sum(a ** 2) * torch.sum(b ** 2) + EPS_NCC) ** 0.5)
    return (1 - ncc) * alpha


def _kde_pdf(data, steps, bandwidth):
    """utils.py:18-53 (K_gauss, PDF_xis, get_pdf).  Note the swapped min/max
    (:45-46): the line samples run from max down to min, detached via .item()."""
    sig = torch.flatten(data, start_dim=1)
    hi, lo = torch.max(sig).item(), torch.min(sig).item()
    xs = torch.linspace(hi, lo, steps, dtype=torch.float, device=sig.device).to(sig.dtype)
    cols = []
    for i in range(steps):                      # chunked: avoids the [n, P, 256] temporary
        z = (sig - xs[i]) / bandwidth
        k = (1 / (2 * math.pi)) * torch.exp(-(z ** 2) / 2)
        cols.append((1 / bandwidth) * torch.mean(k, dim=1))
    return torch.stack(cols, dim=1)


def nmi_loss(t, w, alpha: float = 1000.0, bins: int = 256, patch: int = 100,
             bandwidth: float = 3.0):
    """utils.py:56-79 (NMI) + :236-259 (NMILoss.forward)."""
    r = patch * 2
    nd = t.dim() - 2
    size = (r,) * nd
    n_chunks = 2 ** nd
    tt = F.interpolate(t, size=size, mode="nearest")
    ww = F.interpolate(w, size=size, mode="nearest")
    tt = tt.view(n_chunks * tt.shape[0] * tt.shape[1], *([patch] * nd))
    ww = ww.view(n_chunks * ww.shape[0] * ww.shape[1], *([patch] * nd))
    h1 = _kde_pdf(tt, bins, bandwidth)
    h2 = _kde_pdf(ww, bins, bandwidth)
    hj = _kde_pdf(torch.stack((tt, ww), dim=1), bins, bandwidth)
    p1 = h1 / h1.sum(dim=1, keepdim=True)
    p2 = h2 / h2.sum(dim=1, keepdim=True)
    pj = hj / hj.sum(dim=1, keepdim=True)
    e1 = -torch.sum(p1 * -torch.log2(p1 + 1e-10), dim=1)
    e2 = -torch.sum(p2 * -torch.log2(p2 + 1e-10), dim=1)
    ej = -torch.sum(pj * -torch.log2(pj + 1e-10), dim=1)
    mi = e1 + e2 - ej
    nmi = 2 * mi / (e1 + e2)
    return torch.mean(torch.abs(nmi - 1.0) * alpha)


def weighted_loss(t, w, weights):
    """warpings.py:78-79,144-145,213-214: sum(w_i * crit_i(target, warped)) over
    (MSE, NCC, NMI); zero-weighted terms are skipped (verified equivalent,
    SURVEY.md §4)."""
    terms = (mse_loss, ncc_loss, nmi_loss)
    total = 0
    for wt, fn in zip(weights, terms):
        if wt != 0:
            total = total + wt * fn(t, w)
    return total


# --------------------------------------------------------------------------- #
# rigid / affine epoch loop
# --------------------------------------------------------------------------- #
def affine_like_loop(moving, target, mode, p0, lr, epochs, weights=(1.0, 0.0, 0.0),
                     keep_warped=False, record_grads=False, optimiser="sgd", betas=(0.9, 0.999), eps=1e-8):
    """warpings.py:117-174 (rigid_register) and :30-113 (affine_register).

    mode 'rigid': p = Regressor.reg (utils.py:313-322), theta = rigid_theta(p).
    mode 'affine': p IS theta flattened, initialised to identity — the reference's
    zero-initialised MLP is inert under momentum-free SGD (probed
    bit-exact), so only its output bias (= p) moves.
    Plain SGD, lr, no momentum (warpings.py:58,131).  Best = strictly lower loss,
    evaluated pre-step (:85-93,151-159).
    optimiser='adam' (EXTENSION, north_star item 3, no reference counterpart): torch.optim.Adam on the same
    parameters with the same autograd gradient.
    """
    nd = moving.dim() - 2
    p = p0.clone().detach().to(moving.dtype).requires_grad_(True)
    opt = torch.optim.Adam([p], lr, betas=betas, eps=eps) if optimiser == "adam" else None
    losses, grads = [], []
    best_loss, best_theta, best_warped = None, None, None
    for _ in range(epochs):
        if p.grad is not None:
            p.grad = None
        theta = rigid_theta(p) if mode == "rigid" else p.view(1, nd, nd + 1)
        warped = affine_warp(theta, moving)
        err = weighted_loss(target, warped, weights)
        err.backward()
        if record_grads:
            grads.append(p.grad.detach().clone())
        lv = err.item()
        losses.append(lv)
        if best_loss is None or lv < best_loss:      # pre-step theta (in affine mode theta views p)
            best_loss, best_theta = lv, theta.detach().clone()
            if keep_warped:
                best_warped = warped.detach().clone()
        if opt is not None:
            opt.step()
        else:
            with torch.no_grad():
                p -= lr * p.grad
    with torch.no_grad():
        final_theta = (rigid_theta(p) if mode == "rigid" else p.view(1, nd, nd + 1)).clone()
        final_warped = affine_warp(final_theta, moving) if keep_warped else None
    return {"losses": losses, "final_theta": final_theta, "best_theta": best_theta,
            "final_params": p.detach().clone(), "best_loss": best_loss,
            "final_warped": final_warped, "best_warped": best_warped, "grads": grads}


def affine_step_terms(moving, target, theta, weights):
    """One forward+backward at a fixed theta: returns (loss, dL/dtheta, warped)."""
    nd = moving.dim() - 2
    th = theta.clone().detach().reshape(1, nd, nd + 1).to(moving.dtype).requires_grad_(True)
    warped = affine_warp(th, moving)
    err = weighted_loss(target, warped, weights)
    err.backward()
    return err.item(), th.grad.detach().clone(), warped.detach()


# --------------------------------------------------------------------------- #
# flow node (what sits between the U-Net output and the scalar loss)
# --------------------------------------------------------------------------- #
def flow_node(moving, target, flow, weights=(0.5, 0.5, 0.0), grad_out_warped=None):
    """utils.py:350-365 + warpings.py:213-215: loss(target, warp(moving, flow))
    and d(loss)/d(flow).  With `grad_out_warped` given, returns the plain
    vector-Jacobian product of the warp instead (for user-supplied criteria)."""
    fl = flow.clone().detach().requires_grad_(True)
    warped = flow_warp(moving, fl)
    if grad_out_warped is not None:
        warped.backward(grad_out_warped)
        return None, fl.grad.detach().clone(), warped.detach()
    err = weighted_loss(target, warped, weights)
    err.backward()
    return err.item(), fl.grad.detach().clone(), warped.detach()


def smoothness_l2(flow):
    """VoxelMorph-style Grad('l2') penalty: mean over axes of the mean squared
    forward difference.  EXTENSION — no counterpart in the reference."""
    nd = flow.dim() - 2
    tot = 0
    for a in range(nd):
        d = torch.diff(flow, dim=2 + a)
        tot = tot + torch.mean(d * d)
    return tot / nd


def direct_flow_loop(moving, target, lr, epochs, weights=(0.5, 0.5, 0.0), smooth=0.0,
                     optimiser="sgd", betas=(0.9, 0.999), eps=1e-8, flow0=None):
    """EXTENSION (north_star item 2b/3), parity unpinned by the reference: per-voxel
    flow optimised directly with SGD or Adam, assembled from the reference's own
    SpatialTransformer + MSE/NCC terms + a smoothness penalty."""
    nd = moving.dim() - 2
    flow = (torch.zeros(1, nd, *moving.shape[2:], dtype=moving.dtype) if flow0 is None
            else flow0.clone().detach()).requires_grad_(True)
    opt = (torch.optim.SGD([flow], lr) if optimiser == "sgd"
           else torch.optim.Adam([flow], lr, betas=betas, eps=eps))
    losses = []
    for _ in range(epochs):
        opt.zero_grad()
        warped = flow_warp(moving, flow)
        err = weighted_loss(target, warped, weights)
        if smooth:
            err = err + smooth * smoothness_l2(flow)
        err.backward()
        opt.step()
        losses.append(err.item())
    return {"losses": losses, "flow": flow.detach().clone()}


# --------------------------------------------------------------------------- #
# Edge3D pre-filter
# --------------------------------------------------------------------------- #
def sobel_kernels3d(n1=1, n2=2, n3=2):
    """utils.py:82-127 (get_sobel_kernel3D), restated with numpy exactly as the reference writes it."""
    import numpy as np
    Sx = np.asarray([[[-n1, 0, n1], [-n2, 0, n2], [-n1, 0, n1]], [[-n2, 0, n2], [-n3 * n2, 0, n3 * n2], [-n2, 0, n2]],
                     [[-n1, 0, n1], [-n2, 0, n2], [-n1, 0, n1]]])
    Sy = np.asarray([[[-n1, -n2, -n1], [0, 0, 0], [n1, n2, n1]], [[-n2, -n3 * n2, -n2], [0, 0, 0], [n2, n3 * n2, n2]],
                     [[-n1, -n2, -n1], [0, 0, 0], [n1, n2, n1]]])
    Sz = np.asarray([[[-n1, -n2, -n1], [-n2, -n3 * n2, -n2], [-n1, -n2, -n1]], [[0, 0, 0], [0, 0, 0], [0, 0, 0]],
                     [[n1, n2, n1], [n2, n3 * n2, n2], [n1, n2, n1]]])
    Sd11 = np.asarray([[[0, n1, n2], [-n1, 0, n1], [-n2, -n1, 0]], [[0, n2, n2 * n3], [-n2, 0, n2], [-n2 * n3, -n2, 0]],
                       [[0, n1, n2], [-n1, 0, n1], [-n2, -n1, 0]]])
    Sd12 = np.asarray([[[-n2, -n1, 0], [-n1, 0, n1], [0, n1, n2]], [[-n2 * n3, -n2, 0], [-n2, 0, n2], [0, n2, n2 * n3]],
                       [[-n2, -n1, 0], [-n1, 0, n1], [0, n1, n2]]])
    Sd21, Sd22 = Sd11.T, Sd12.T
    Sd31 = np.asarray([-S.T for S in Sd11.T])
    Sd32 = np.asarray([S.T for S in Sd12.T])
    return [Sx, Sy, Sz, Sd11, Sd12, Sd21, Sd22, Sd31, Sd32]


def edge3d(img, a=1, thresh=(0.2, 0.9), return_norm=False):
    """utils.py:153-183 (Edge3D.__call__): reflect pad by a, nine 3x3x3 cross-correlations (conv3d, padding 1), magnitude
    (1/C) sqrt(sum_s (sum_c (corr + eps))^2 + eps), crop, min-max norm (:262-267), band threshold."""
    import numpy as np
    eps = 1e-10
    B, C = img.shape[:2]
    x = F.pad(img, (a,) * 6, mode="reflect")
    mags = []
    for k in sobel_kernels3d():
        wgt = torch.from_numpy(k.astype(np.float32)).reshape(1, 1, 3, 3, 3).to(img)
        per_c = torch.cat([F.conv3d(x[:, c:c + 1], wgt, padding=1) for c in range(C)], dim=1)
        mags.append(torch.sum(per_c + eps, dim=1) ** 2)
    g = (1 / C) * torch.sum(torch.stack(mags, dim=1) + eps, dim=1) ** 0.5
    g = g[:, a:-a, a:-a, a:-a].reshape(B, 1, *img.shape[2:])
    e = (g - torch.min(g)) / ((torch.max(g) - torch.min(g)) + 1e-9)
    out = ((e > thresh[0]) & (e < thresh[1])).to(torch.float32)
    return (out, e) if return_norm else out
