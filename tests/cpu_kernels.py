"""TEST-ONLY stand-ins for the libttlda entry points, computing on CPU tensors with NumPy/SciPy and the oracle.

They exist so that the HOST logic of the multi-rank path (document sharding, the statistics all-reduce, identical
convergence decisions on every rank, gathering gamma) can be exercised with the gloo backend on a machine without
GPUs.  ``install(monkeypatch_or_module)`` replaces the wrappers in ``tuotuo_b200._native``; the product package never
imports this file and has no CPU path of its own.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.special import gammaln, psi

import lda_oracle as orc


def _np(t):
    return t.numpy()


def check_device(device=0):
    return None


def reduce_workspace_bytes():
    return 64


def mstep_workspace_bytes(V, ld):
    return 64


def sstats_workspace_bytes(nnz, V, ld, nblocks=1):
    return 64


def csr_to_csc_workspace_bytes(nnz, D, V, nblocks=1):
    return 64


def csr_to_csc(rowptr, colidx, counts, D, V, nnz, nblocks, block_docs, colptr, docidx, ccounts, ws, csr2csc=None):
    rp, ci, ct = _np(rowptr), _np(colidx)[:nnz], _np(counts)[:nnz]
    docs = np.repeat(np.arange(D), np.diff(rp))
    keys = (docs // block_docs) * V + ci
    order = np.argsort(keys, kind="stable")
    _np(docidx)[:nnz] = docs[order]
    _np(ccounts)[:nnz] = ct[order]
    if csr2csc is not None:
        inv = np.empty(nnz, dtype=np.int64)
        inv[order] = np.arange(nnz)
        _np(csr2csc)[:nnz] = inv.astype(np.int32)
    cp = np.zeros(nblocks * V + 1, dtype=np.int64)
    np.add.at(cp, keys + 1, 1)
    _np(colptr)[:] = np.cumsum(cp)


def transpose_pad(src, dst):
    rows, cols = src.shape
    d = _np(dst)
    d[:] = 0.0
    d[:, :rows] = _np(src).T


def transpose_unpad(src, cols, dst):
    _np(dst)[:] = _np(src)[:, :cols].T


def _theta_term(gamma, E, alpha_vec, alpha_sum):
    return float(np.sum((alpha_vec - gamma) * E) + np.sum(gammaln(gamma) - gammaln(alpha_vec))
                 + np.sum(gammaln(alpha_sum) - gammaln(gamma.sum(1))))


def theta_refresh(gamma, K, alpha_vec, alpha_sum, eT, partial, ws):
    g = _np(gamma)[:, :K]
    E = orc.dirichlet_expectation_2d(g) if g.shape[0] else np.zeros_like(g)
    if eT is not None:
        _np(eT)[:] = 0.0
        _np(eT)[:, :K] = np.exp(E)
    p = _np(partial)
    p[:] = 0.0
    p[0] = _theta_term(g, E, _np(alpha_vec), alpha_sum)


def beta_refresh(lam, K, eta, eB, elogbeta, colsum, partial, ws):
    lkv = np.ascontiguousarray(_np(lam)[:, :K].T)  # (K, V)
    E = orc.dirichlet_expectation_2d(lkv)
    _np(eB)[:] = 0.0
    _np(eB)[:, :K] = np.exp(E).T
    if elogbeta is not None:
        _np(elogbeta)[:] = 0.0
        _np(elogbeta)[:, :K] = E.T
    _np(colsum)[:] = 0.0
    _np(colsum)[:K] = lkv.sum(1)
    p = _np(partial)
    p[:] = 0.0
    p[0] = float(np.sum((eta - lkv) * E) + np.sum(gammaln(lkv) - gammaln(eta))
                 + np.sum(gammaln(eta) - gammaln(lkv.sum(1))))  # scalar eta: np.sum(eta) == eta


def _coo(rowptr, colidx, counts):
    rp = _np(rowptr)
    nnz = int(rp[-1])
    return np.repeat(np.arange(len(rp) - 1), np.diff(rp)), _np(colidx)[:nnz].astype(np.int64), _np(counts)[:nnz].astype(np.float64)


def estep(rowptr, colidx, counts, K, V, alpha_vec, alpha_sum, eT_in, eB, gamma_in, gamma_out, eT_out, max_iter, tol,
          iters, partial, ws, csr2csc=None, ratio_csc=None):
    d, w, c = _coo(rowptr, colidx, counts)
    t, b, a = _np(eT_in)[:, :K], _np(eB)[:, :K], _np(alpha_vec)
    dot = np.einsum("ik,ik->i", t[d], b[w])
    r = c / (dot + 1e-100)
    if ratio_csc is not None:
        _np(ratio_csc)[_np(csr2csc)[:len(r)].astype(np.int64)] = r
    acc = np.zeros_like(t)
    np.add.at(acc, d, r[:, None] * b[w])
    g_new = a + t * acc
    if iters is not None:
        l1 = np.abs(g_new - _np(gamma_in)[:, :K]).sum(1)
        passes = np.where((l1 > tol) & (max_iter >= 1), 2, 1)
        _np(iters)[0] = int(passes.sum())
        _np(iters)[1] = int((passes > max_iter).sum())
    _np(gamma_out)[:] = 0.0
    _np(gamma_out)[:, :K] = g_new
    p = _np(partial)
    p[:] = 0.0
    p[0] = float(np.sum(c * np.log(dot)))
    p[2] = float(c.sum())
    if eT_out is not None:  # fused theta refresh; otherwise the driver runs theta_refresh itself
        E = orc.dirichlet_expectation_2d(g_new) if g_new.shape[0] else g_new
        _np(eT_out)[:] = 0.0
        _np(eT_out)[:, :K] = np.exp(E)
        p[1] = _theta_term(g_new, E, a, alpha_sum)


def elbo_tokens(rowptr, colidx, counts, K, eT, eB, partial, ws):
    d, w, c = _coo(rowptr, colidx, counts)
    dot = np.einsum("ik,ik->i", _np(eT)[:, :K][d], _np(eB)[:, :K][w])
    p = _np(partial)
    p[:] = 0.0
    p[0] = float(np.sum(c * np.log(dot)))
    p[2] = float(c.sum())


def sstats_csc(colptr, docidx, ccounts, V, K, nnz, nblocks, eT, eB, sstats, ws, ratio_csc=None):
    cp = _np(colptr)
    vw = np.repeat(np.arange(len(cp) - 1), np.diff(cp))
    w = vw % V
    d = _np(docidx)[:nnz].astype(np.int64)
    c = _np(ccounts)[:nnz].astype(np.float64)
    t, b = _np(eT)[:, :K], _np(eB)[:, :K]
    if ratio_csc is not None:
        r = _np(ratio_csc)[:nnz]
    else:
        r = c / (np.einsum("ik,ik->i", t[d], b[w]) + 1e-100)
    S = _np(sstats)
    S[:] = 0.0
    np.add.at(S[:, :K], w, r[:, None] * t[d])


def sstats_atomic(rowptr, colidx, counts, K, V, eT, eB, sstats):
    d, w, c = _coo(rowptr, colidx, counts)
    t, b = _np(eT)[:, :K], _np(eB)[:, :K]
    r = c / (np.einsum("ik,ik->i", t[d], b[w]) + 1e-100)
    S = _np(sstats)
    S[:] = 0.0
    np.add.at(S[:, :K], w, r[:, None] * t[d])


def mstep(sstats, K, eta, eB, lam, elogbeta, colsum, partial, ws):
    L = _np(lam)
    L[:] = 0.0
    L[:, :K] = eta + _np(sstats)[:, :K] * _np(eB)[:, :K]
    beta_refresh(lam, K, eta, eB, elogbeta, colsum, partial, ws)


def hyper_stats(x, cols, partial, ws):
    a = _np(x)[:, :cols]
    p = _np(partial)
    p[:] = 0.0
    p[0] = float(psi(a).sum())
    p[1] = float(psi(a.sum(1)).sum())


def psi_exact(x, out):
    _np(out)[:] = psi(_np(x))


def dirichlet_expectation(a, cols, out_elog=None, out_exp=None):
    E = orc.dirichlet_expectation_2d(np.ascontiguousarray(_np(a)[:, :cols]))
    if out_elog is not None:
        _np(out_elog)[:] = 0.0
        _np(out_elog)[:, :cols] = E
    if out_exp is not None:
        _np(out_exp)[:] = 0.0
        _np(out_exp)[:, :cols] = np.exp(E)


def estep_hot_rows(ld):
    return 0  # no shared-memory row cache in the stand-ins: the driver then runs the plain `estep`


NAMES = ["check_device", "estep_hot_rows", "reduce_workspace_bytes", "mstep_workspace_bytes", "sstats_workspace_bytes",
         "csr_to_csc_workspace_bytes", "csr_to_csc", "transpose_pad", "transpose_unpad", "theta_refresh", "beta_refresh",
         "estep", "elbo_tokens", "sstats_csc", "sstats_atomic", "mstep", "hyper_stats", "psi_exact",
         "dirichlet_expectation"]


def install():
    """Replace the CUDA wrappers of tuotuo_b200._native by the CPU stand-ins above (test processes only)."""
    from tuotuo_b200 import _native

    for n in NAMES:
        setattr(_native, n, globals()[n])
