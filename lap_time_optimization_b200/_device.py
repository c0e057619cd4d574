"""Device plumbing shared by the facade classes: torch owns device memory and streams, libltk does
the arithmetic.  Nothing here computes on the host."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native


def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        raise _native.LtkUnavailable("no CUDA device visible: the lap-time kernels are sm_100a-only and "
                                     "there is no CPU fallback")
    return torch


def current_device():
    return torch_cuda().cuda.current_device()


def stream_ptr(torch, device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def to_device(arr, device):
    torch = torch_cuda()
    return torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(device, non_blocking=False)


def path_eval(path, u, want):
    """Run ltk_path_eval for `Path` facade objects; returns a dict of numpy arrays / scalars."""
    torch = torch_cuda()
    lib = _native.load()
    dev = torch.device("cuda", current_device())
    xy = to_device(np.asarray(path.controls, dtype=np.float64), dev)
    knots = to_device(path.dists, dev)
    du = to_device(u, dev)
    n = du.numel()
    m = xy.shape[1]
    names = ("x", "y", "dx", "dy", "ddx", "ddy", "k")
    bufs = {nm: (torch.empty(n, dtype=torch.float64, device=dev) if nm in want else None) for nm in names}
    g2 = torch.zeros(1, dtype=torch.float64, device=dev) if "gamma2" in want else None
    rc = lib.ltk_path_eval(dev.index, ptr(xy), ptr(knots), m, ptr(du), n, *[ptr(bufs[nm]) for nm in names],
                           ptr(g2), stream_ptr(torch, dev))
    _native.check(rc)
    out = {nm: b.cpu().numpy() for nm, b in bufs.items() if b is not None}
    if g2 is not None:
        out["gamma2"] = np.float64(g2.item())
    return out


def path_eval_fitpack(path, u, want):
    """Run ltk_path_eval_fitpack (FITPACK arithmetic, closed or open paths); `want` may include "tck"."""
    torch = torch_cuda()
    lib = _native.load()
    dev = torch.device("cuda", current_device())
    xy = to_device(np.asarray(path.controls, dtype=np.float64), dev)
    knots = to_device(path.dists, dev)
    du = to_device(u, dev)
    n = du.numel()
    m = xy.shape[1]
    closed = 1 if path.closed else 0
    names = ("x", "y", "dx", "dy", "ddx", "ddy", "k")
    bufs = {nm: (torch.empty(n, dtype=torch.float64, device=dev) if nm in want else None) for nm in names}
    g2 = torch.zeros(1, dtype=torch.float64, device=dev) if "gamma2" in want else None
    nt = m + 6 if closed else m + 4
    d_t = torch.empty(nt, dtype=torch.float64, device=dev) if "tck" in want else None
    d_c = torch.empty((2, nt - 4), dtype=torch.float64, device=dev) if "tck" in want else None
    rc = lib.ltk_path_eval_fitpack(dev.index, ptr(xy), ptr(knots), m, closed, ptr(du) if n else None, n,
                                   *[ptr(bufs[nm]) for nm in names], ptr(g2), ptr(d_t), ptr(d_c), stream_ptr(torch, dev))
    _native.check(rc)
    out = {nm: b.cpu().numpy() for nm, b in bufs.items() if b is not None}
    if g2 is not None:
        out["gamma2"] = np.float64(g2.item())
    if d_t is not None:
        c = d_c.cpu().numpy()
        out["t"], out["cx"], out["cy"] = d_t.cpu().numpy(), c[0].copy(), c[1].copy()
    return out


def velocity_profile(vehicle, s, k, s_max):
    """Run ltk_velocity_profile; returns (v_local, v_acclim, v_declim, v) as numpy arrays."""
    torch = torch_cuda()
    lib = _native.load()
    dev = torch.device("cuda", current_device())
    ds, dk = to_device(s, dev), to_device(k, dev)
    n = ds.numel()
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)]
    veh = vehicle.to_ltk()
    rc = lib.ltk_velocity_profile(dev.index, C.byref(veh), ptr(ds), ptr(dk), n,
                                  float(-1.0 if s_max is None else s_max), *[ptr(o) for o in outs],
                                  stream_ptr(torch, dev))
    _native.check(rc)
    return tuple(o.cpu().numpy() for o in outs)
