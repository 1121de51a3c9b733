"""``tuotuo.cutils``-compatible entry points (tuotuo/cutils.pyx:11-68) over the sm_100a kernels.

NumPy in / NumPy out like the Cython originals: float32 or float64 arrays accepted, anything else raises
``TypeError`` (the fused-type dispatch error of the reference).  Arithmetic is float64 on the device; a float32
input is widened, processed and narrowed back (the reference evaluates it in float32).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nv


def _dev():
    if not torch.cuda.is_available():
        raise nv.TTLDAError("tuotuo_b200.cutils needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _dirichlet_expectation_2d(arr: np.ndarray) -> np.ndarray:
    if not isinstance(arr, np.ndarray) or arr.dtype not in (np.float32, np.float64) or arr.ndim != 2:
        raise TypeError("No matching signature found")
    rows, cols = arr.shape
    ld = nv.ld_for(max(cols, 1))
    a = torch.zeros((rows, ld), dtype=torch.float64, device=_dev())
    a[:, :cols] = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(a.device)
    out = torch.empty_like(a)
    nv.dirichlet_expectation(a, cols, out_elog=out)
    return out[:, :cols].cpu().numpy().astype(arr.dtype, copy=False)


def _dirichlet_expectation_1d(doc_topic: np.ndarray, doc_topic_prior: float, out: np.ndarray) -> None:
    if (not isinstance(doc_topic, np.ndarray) or doc_topic.dtype not in (np.float32, np.float64)
            or doc_topic.ndim != 1 or out.dtype != doc_topic.dtype):
        raise TypeError("No matching signature found")
    dev = _dev()
    dt = torch.as_tensor(np.ascontiguousarray(doc_topic, dtype=np.float64)).to(dev)
    o = torch.empty_like(dt)
    nv.dirichlet_expectation_1d(dt, float(doc_topic_prior), o)
    doc_topic[:] = dt.cpu().numpy()  # in-place prior add (cutils.pyx:31-33)
    out[:] = o.cpu().numpy()
