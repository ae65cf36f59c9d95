"""TEST INFRASTRUCTURE — ctypes front-end of the plain-C CPU oracle (c_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this.  numpy in, numpy out; `dtype` selects the f32 / f64 instantiation.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("c_oracle.c", "c_oracle_impl.inc", "Makefile")]
    stale = (not os.path.isfile(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True, env=env)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", C.c_float
    if dtype == np.float64:
        return "_f64", C.c_double
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _dims(vol, ndim):
    if ndim == 3:
        D, H, W = vol.shape[-3:]
    else:
        D, (H, W) = 1, vol.shape[-2:]
    return int(D), int(H), int(W)


def affine_terms(moving, target, theta, w_mse, w_ncc, want_warped=False):
    """-> (loss: float, dtheta: float64[ndim, ndim+1], warped or None)."""
    moving = np.ascontiguousarray(moving); target = np.ascontiguousarray(target, dtype=moving.dtype)
    ndim = 3 if moving.squeeze().ndim == 3 else 2
    vol = moving.reshape(moving.shape[-ndim:])
    D, H, W = _dims(vol, ndim)
    sfx, _ = _sfx(moving.dtype)
    th = np.ascontiguousarray(np.asarray(theta, dtype=moving.dtype).reshape(-1))
    loss = C.c_double(0)
    dth = np.zeros(12, np.float64)
    warped = np.empty_like(vol) if want_warped else None
    getattr(lib(), "orc_affine_terms" + sfx)(
        C.c_int(ndim), _p(vol), _p(target), C.c_int(D), C.c_int(H), C.c_int(W), _p(th),
        C.c_double(w_mse), C.c_double(w_ncc), C.byref(loss), _p(dth), _p(warped))
    nt = ndim * (ndim + 1)
    return loss.value, dth[:nt].reshape(ndim, ndim + 1).copy(), warped


def rigid_theta(p):
    p = np.ascontiguousarray(p)
    ndim = 3 if p.size > 3 else 2
    sfx, _ = _sfx(p.dtype)
    th = np.zeros(12, p.dtype)
    getattr(lib(), "orc_rigid_theta" + sfx)(C.c_int(ndim), _p(p), _p(th))
    return th[:ndim * (ndim + 1)].reshape(ndim, ndim + 1).copy()


def rigid_chain(p, dtheta):
    p = np.ascontiguousarray(p)
    ndim = 3 if p.size > 3 else 2
    sfx, _ = _sfx(p.dtype)
    dth = np.zeros(12, np.float64); dth[:ndim * (ndim + 1)] = np.asarray(dtheta, np.float64).reshape(-1)
    dp = np.zeros(12, np.float64)
    getattr(lib(), "orc_rigid_chain" + sfx)(C.c_int(ndim), _p(p), _p(dth), _p(dp))
    return dp[:p.size].copy()


def affine_loop(moving, target, mode, p0, lr, epochs, w_mse, w_ncc):
    """-> dict(losses, best_theta, final_theta, final_params)."""
    moving = np.ascontiguousarray(moving); target = np.ascontiguousarray(target, dtype=moving.dtype)
    ndim = 3 if moving.squeeze().ndim == 3 else 2
    vol = moving.reshape(moving.shape[-ndim:])
    D, H, W = _dims(vol, ndim)
    sfx, _ = _sfx(moving.dtype)
    nt = ndim * (ndim + 1)
    params = np.zeros(12, moving.dtype)
    p0 = np.asarray(p0, moving.dtype).reshape(-1)
    params[:p0.size] = p0
    log = np.zeros(max(epochs, 1), moving.dtype)
    best = np.zeros(12, moving.dtype); final = np.zeros(12, moving.dtype)
    getattr(lib(), "orc_affine_loop" + sfx)(
        C.c_int(ndim), C.c_int(0 if mode == "rigid" else 1), _p(vol), _p(target),
        C.c_int(D), C.c_int(H), C.c_int(W), _p(params), C.c_double(lr), C.c_int(epochs),
        C.c_double(w_mse), C.c_double(w_ncc), _p(log), _p(best), _p(final))
    return {"losses": log[:epochs].copy(), "best_theta": best[:nt].reshape(ndim, ndim + 1).copy(),
            "final_theta": final[:nt].reshape(ndim, ndim + 1).copy(),
            "final_params": params[:p0.size].copy()}


def warp_affine(moving, theta):
    moving = np.ascontiguousarray(moving)
    ndim = 3 if moving.squeeze().ndim == 3 else 2
    vol = moving.reshape(moving.shape[-ndim:])
    D, H, W = _dims(vol, ndim)
    sfx, _ = _sfx(moving.dtype)
    th = np.ascontiguousarray(np.asarray(theta, dtype=moving.dtype).reshape(-1))
    out = np.empty_like(vol)
    getattr(lib(), "orc_warp_affine" + sfx)(C.c_int(ndim), _p(vol), C.c_int(D), C.c_int(H), C.c_int(W),
                                            _p(th), _p(out))
    return out


def flow_terms(moving, target, flow, w_mse, w_ncc, gout=None):
    """-> (loss, dflow[ndim, ...], warped).  gout given: plain VJP of the warp."""
    moving = np.ascontiguousarray(moving)
    ndim = 3 if moving.squeeze().ndim == 3 else 2
    vol = moving.reshape(moving.shape[-ndim:])
    tgt = np.ascontiguousarray(target, dtype=moving.dtype).reshape(vol.shape)
    fl = np.ascontiguousarray(flow, dtype=moving.dtype).reshape((ndim,) + vol.shape)
    D, H, W = _dims(vol, ndim)
    sfx, _ = _sfx(moving.dtype)
    loss = C.c_double(0)
    dflow = np.empty_like(fl); warped = np.empty_like(vol)
    g = None if gout is None else np.ascontiguousarray(gout, dtype=moving.dtype).reshape(vol.shape)
    getattr(lib(), "orc_flow_terms" + sfx)(
        C.c_int(ndim), _p(vol), _p(tgt), _p(fl), C.c_int(D), C.c_int(H), C.c_int(W),
        C.c_double(w_mse), C.c_double(w_ncc), _p(g), C.byref(loss), _p(dflow), _p(warped))
    return loss.value, dflow, warped
