"""Synthetic LDA-generative corpora for the benchmark configurations (SURVEY.md §8d / BASELINE.json configs).

Token ids are drawn directly on the device (no strings):
    beta_k  ∝ zipf(rank_v)^s ∘ Gamma(0.05)   row-normalised  (Zipf factor => realistic repeated words per document)
    theta_d ~ Dirichlet(0.1 · 1_K),   z ~ Cat(theta_d),   w ~ Cat(beta_z) by inverse CDF
and handed to ``Document.from_token_ids`` (bag-of-words counting by ``ttlda_bow_to_csr``).  torch is used as an
RNG/plumbing library here; nothing in this file is on the timed path.
"""
from __future__ import annotations

import numpy as np
import torch

from .document import Document

CONFIGS = {
    # name: (D, L, K, V, poisson_length)
    "cfg1": (100, 20, 5, 40, False),
    "cfg2": (10_000, 200, 20, 5_000, False),
    "cfg3": (1_000_000, 300, 100, 50_000, False),
    "cfg4": (300_000, 330, 500, 102_660, True),
    "cfg5": (10_000_000, 200, 1000, 100_000, False),
}


def synthetic_token_stream(D: int, L: int, K: int, V: int, *, device, seed: int = 1234, poisson: bool = False,
                           zipf_s: float = 1.0, chunk_docs: int = 65536, doc_seed: int | None = None):
    """Returns (token_ids int32[T], doc_offsets int64[D+1]) on `device`.

    `seed` fixes the generative MODEL (the topics beta and the vocabulary permutation) and, when `doc_seed` is None, the
    documents drawn from it (one stream, as always).  `doc_seed` draws the documents from their own stream instead:
    document shards of ONE corpus (same model, different documents) for weak-scaling runs — with a different `seed` per
    rank every rank would hold a different model, and shards whose distinct-words-per-document differ by several percent
    (measured: 203.7M ... 218.2M nonzeros for 1M x 300-token documents), i.e. unequal work for equal token counts."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rank = torch.arange(1, V + 1, dtype=torch.float64, device=dev)
    sparse = torch._standard_gamma(torch.full((K, V), 0.05, dtype=torch.float64, device=dev), generator=g)
    beta = (sparse + 1e-12) * rank.pow(-zipf_s)
    perm = torch.randperm(V, generator=g, device=dev)  # decouple vocabulary id from frequency rank
    beta = beta[:, perm]
    beta = beta / beta.sum(1, keepdim=True)
    cdf = torch.cumsum(beta, 1)
    cdf[:, -1] = 1.0
    cdf_flat = (cdf + torch.arange(K, dtype=torch.float64, device=dev)[:, None]).reshape(-1)
    if doc_seed is not None:
        g = torch.Generator(device=dev)
        g.manual_seed(doc_seed)

    if poisson:
        lens = torch.poisson(torch.full((D,), float(L), device=dev), generator=g).to(torch.int64).clamp_(min=1)
    else:
        lens = torch.full((D,), int(L), dtype=torch.int64, device=dev)
    offs = torch.zeros(D + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(lens, 0)
    T = int(offs[-1].item())
    tokens = torch.empty(T, dtype=torch.int32, device=dev)
    Lmax = int(lens.max().item()) if D else 0
    for d0 in range(0, D, chunk_docs):
        d1 = min(D, d0 + chunk_docs)
        n = d1 - d0
        theta = torch._standard_gamma(torch.full((n, K), 0.1, dtype=torch.float64, device=dev), generator=g) + 1e-300
        theta = theta / theta.sum(1, keepdim=True)
        z = torch.multinomial(theta, Lmax, replacement=True, generator=g)  # (n, Lmax)
        u = torch.rand((n, Lmax), dtype=torch.float64, device=dev, generator=g).clamp_(max=1 - 1e-12)
        w = torch.searchsorted(cdf_flat, (u + z.to(torch.float64)).reshape(-1)).reshape(n, Lmax) - z * V
        w = w.clamp_(0, V - 1).to(torch.int32)
        if poisson:
            keep = torch.arange(Lmax, device=dev)[None, :] < lens[d0:d1, None]
            tokens[int(offs[d0].item()):int(offs[d1].item())] = w[keep]
        else:
            tokens[int(offs[d0].item()):int(offs[d1].item())] = w.reshape(-1)
    return tokens, offs


def synthetic_document(name_or_shape, *, device, seed: int = 1234, docs: int | None = None,
                       doc_seed: int | None = None) -> Document:
    """A ``Document`` for one of the BASELINE.json configurations ("cfg2", "cfg3", ...) or a (D, L, K, V) tuple.
    `docs` overrides the number of documents (shards / bounded samples keep the K, V, L of the configuration);
    `doc_seed`: the documents' own random stream (synthetic_token_stream), for shards of one corpus."""
    if isinstance(name_or_shape, str):
        D, L, K, V, pois = CONFIGS[name_or_shape]
    else:
        D, L, K, V = name_or_shape
        pois = False
    if docs is not None:
        D = int(docs)
    tok, offs = synthetic_token_stream(D, L, K, V, device=device, seed=seed, poisson=pois, doc_seed=doc_seed)
    doc = Document.from_token_ids(tok, offs, V)
    doc.config = dict(D=D, L=L, K=K, V=V, poisson=pois, seed=seed, doc_seed=doc_seed)
    return doc


def numpy_corpus(seed: int, D: int, V: int, K: int, L: int, ragged: bool = False):
    """Small seeded host corpus (CSR) for parity tests against the oracle; identical on every machine."""
    rs = np.random.RandomState(seed)
    beta = rs.gamma(0.05, size=(K, V)) + 1e-12
    beta *= (1.0 / np.arange(1, V + 1)) ** 0.8
    beta /= beta.sum(1, keepdims=True)
    cdf = np.cumsum(beta, 1)
    cdf[:, -1] = 1.0
    rowptr = [0]
    ci, ct = [], []
    for d in range(D):
        n = int(rs.poisson(L)) if ragged else L
        theta = rs.dirichlet(np.full(K, 0.2))
        z = rs.choice(K, size=n, p=theta)
        u = rs.rand(n)
        w = np.array([np.searchsorted(cdf[k], x) for k, x in zip(z, u)], dtype=np.int64)
        uq, c = np.unique(np.clip(w, 0, V - 1), return_counts=True)
        ci.append(uq)
        ct.append(c)
        rowptr.append(rowptr[-1] + len(uq))
    return (np.asarray(rowptr, dtype=np.int64), np.concatenate(ci).astype(np.int32) if ci else np.zeros(0, np.int32),
            np.concatenate(ct).astype(np.int32) if ct else np.zeros(0, np.int32))
