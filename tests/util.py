"""Helpers shared by the tests: numpy <-> torch views that keep strides."""
from __future__ import annotations

import numpy as np
import torch

TOL = {"s": 5e-5, "c": 5e-5, "d": 1e-12, "z": 1e-12}     # elementwise, relative to max(1, |ref|max)
NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
        np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}


def estr(a: np.ndarray):
    return a.strides[0] // a.itemsize, a.strides[1] // a.itemsize


def _span(a: np.ndarray):
    rs, cs = estr(a)
    assert rs >= 0 and cs >= 0
    return (a.shape[0] - 1) * rs + (a.shape[1] - 1) * cs + 1 if a.size else 1


def to_torch(a: np.ndarray, device="cuda", pin: bool = False) -> torch.Tensor:
    """Tensor on `device` with the same shape AND the same element strides as a
    (pin=True: page-locked host memory)."""
    rs, cs = estr(a)
    flat = np.lib.stride_tricks.as_strided(a, shape=(_span(a),), strides=(a.itemsize,))
    t = torch.from_numpy(np.array(flat, copy=True))
    t = t.pin_memory() if pin else t.to(device)
    return t.as_strided(a.shape, (rs, cs))


def to_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().resolve_conj().numpy()


def rel_err(x: np.ndarray, ref: np.ndarray) -> float:
    if x.size == 0:
        return 0.0
    return float(np.abs(x - ref).max() / max(1.0, np.abs(ref).max()))
