"""Device plumbing: torch tensors are used ONLY as device buffers / streams; all compute goes through
libepb200.so.  Every helper here fails loudly when CUDA is unavailable (no CPU fallback)."""

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import EpbError, epb_cp


def require_cuda():
    if not torch.cuda.is_available():
        raise EpbError("echopype_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    _lib.load()
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def empty(shape, dtype=torch.float32, device=None):
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype, device=device or require_cuda())


def to_device_f32(x, device=None, non_blocking=True):
    """Host array / DataArray / tensor -> contiguous float32 CUDA tensor (no copy if already there)."""
    device = device or require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(getattr(x, "data", x))
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.to(device, non_blocking=non_blocking)
    return t.contiguous()


class ParamPack:
    """Keeps the small float64 (channel, ping) parameter tensors alive while kernels use them."""

    def __init__(self, C, P, device):
        self.C, self.P, self.device = int(C), int(P), device
        self._keep = []

    def cp(self, x):
        """scalar | (C,) | (C,1) | (1,P) | (C,P) | (P,) array -> epb_cp on the device (stride 0 = broadcast).
        A 1-D array of length C is per-channel (also when C == P); per-ping parameters should be passed as (1,P)
        (calibrate_ek._cp and consolidate.add_depth do, from the dimension names); a bare (P,) vector is accepted only
        when P != C."""
        C, P = self.C, self.P
        a = np.asarray(getattr(x, "values", x), dtype=np.float64)
        if a.ndim == 0:
            a2, sc, sp = a.reshape(1), 0, 0
        elif a.ndim == 1 and a.shape[0] == C:
            a2, sc, sp = a, 1, 0
        elif a.ndim == 1 and a.shape[0] == P:
            a2, sc, sp = a, 0, 1
        elif a.ndim == 2 and a.shape == (1, P) and not (C == 1):
            a2, sc, sp = a.reshape(P), 0, 1
        elif a.ndim == 2 and a.shape == (C, 1):
            a2, sc, sp = a.reshape(C), 1, 0
        elif a.ndim == 2 and a.shape == (C, P):
            first = a[:, :1]
            same = (a == first) | (np.isnan(a) & np.isnan(first))
            if P > 1 and bool(same.all()):  # constant along ping_time: ship (C,) only
                a2, sc, sp = np.ascontiguousarray(first.reshape(C)), 1, 0
            else:
                a2, sc, sp = np.ascontiguousarray(a), P, 1
        else:
            raise ValueError(f"parameter of shape {a.shape} does not broadcast to (channel={C}, ping_time={P})")
        t = torch.from_numpy(np.array(a2, dtype=np.float64, order="C", copy=True)).to(self.device)  # (C,)/(C,P) parameters: tiny; copy makes read-only views writable for torch
        self._keep.append(t)
        return epb_cp(t.data_ptr(), sc, sp)

    def vec(self, x, dtype=torch.float64):
        a = np.array(np.asarray(getattr(x, "values", x)), order="C", copy=True)  # small; read-only views become writable
        t = torch.from_numpy(a).to(self.device).to(dtype)
        self._keep.append(t)
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def null():
        return epb_cp(None, 0, 0)
