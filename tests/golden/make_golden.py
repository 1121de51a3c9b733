"""Generate the committed golden fixtures by running the REAL reference (``/root/reference/tuotuo``).

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz`` / ``*.json``.  Each fixture stores the seeded *inputs* (CSR corpus, kwargs) and
the reference's *outputs* (perplexity trace, gamma, lambda, alpha, eta, ...).  The reference's RNG use is
deterministic here: lambda_0 comes from ``np.random.seed(42); np.random.gamma`` (tuotuo/lda_model.py:356-357)
and the corpora below are drawn from seeded ``np.random.RandomState`` / ``torch.manual_seed``.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from ref_harness import DenseCorpus, import_reference  # noqa: E402

_spec = importlib.util.spec_from_file_location("_text", os.path.join(ROOT, "tuotuo_b200", "text.py"))
_text = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_text)

lda_model, document, generator, cutils = import_reference(stopwords=list(_text.NLTK_ENGLISH_STOPWORDS))


def synth_dense(seed, D, V, K, L):
    """Small LDA-generative corpus (dense counts), seeded."""
    rs = np.random.RandomState(seed)
    beta = rs.dirichlet(np.full(V, 0.05), size=K) * (1.0 / np.arange(1, V + 1)) ** 0.7
    beta /= beta.sum(1, keepdims=True)
    theta = rs.dirichlet(np.full(K, 0.3), size=D)
    X = np.zeros((D, V))
    for d in range(D):
        n = L if isinstance(L, int) else rs.poisson(L[0])
        z = rs.choice(K, size=n, p=theta[d])
        for k in z:
            X[d, rs.choice(V, p=beta[k])] += 1
    return X


def csr_of(X):
    rows, cols = np.nonzero(X)
    rowptr = np.zeros(X.shape[0] + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return dict(rowptr=np.cumsum(rowptr).astype(np.int64), colidx=cols.astype(np.int32),
                counts=X[rows, cols].astype(np.int32), V=np.int64(X.shape[1]))


def run_fit(X, K, exchangeable=True, D_population=None, **kw):
    lda = lda_model.LDASmoothed(num_topics=K, exchangeable_prior=exchangeable, verbose=False)
    alpha0 = np.array(lda._alpha_, dtype=np.float64)
    if D_population is not None:
        lda.D_population = D_population  # never assigned by the reference
    n_copy = [0]
    orig_copy = np.copy

    def counting_copy(*a, **k):
        n_copy[0] += 1
        return orig_copy(*a, **k)

    lda_model.np.copy = counting_copy  # one np.copy per inner iteration (lda_model.py:238)
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            perps = lda.fit(DenseCorpus(X), return_perplexities=True, **kw)
    finally:
        lda_model.np.copy = orig_copy
    return dict(perplexities=np.array(perps, dtype=np.float64), gamma=np.array(lda._gamma_),
                lam=np.array(lda._lambda_), alpha=np.array(lda._alpha_, dtype=np.float64),
                eta=np.array(lda._eta_, dtype=np.float64), alpha0=alpha0, inner_iters=np.int64(n_copy[0]),
                n_warnings=np.int64(len(w)))


def run_nonexch_hyper(X, K, newton_steps):
    """Non-exchangeable update_alpha (lda_model.py:452-456,464-476) — numerically broken upstream (it diverges when run
    to its 501-step cap), so only its first Newton iterations are pinned."""
    lda = lda_model.LDASmoothed(num_topics=K, exchangeable_prior=False, verbose=False)
    lda.fit(DenseCorpus(X), em_num_step=2, update_hyper_parms=False)
    alpha0 = np.array(lda._alpha_, dtype=np.float64)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        lda.update_alpha(step=newton_steps)
    return dict(alpha0=alpha0, alpha=np.array(lda._alpha_, dtype=np.float64), gamma=np.array(lda._gamma_),
                newton_steps=np.int64(newton_steps), n_warnings=np.int64(len(w)))


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def main():
    if sys.argv[1:] == ["hyper_nonexch"]:  # one fixture only (leaves the other files untouched)
        X = synth_dense(12, D=120, V=200, K=6, L=40)
        save("hyper_nonexch.npz", **csr_of(X), K=np.int64(6), **run_nonexch_hyper(X, 6, 3))
        return
    # ---- A. AS-103 psi / Dirichlet expectation known answers from the compiled reference cutils ------------
    rs = np.random.RandomState(7)
    special = np.array([1e-7, 9.99e-7, 1e-6, 1.0000001e-6, 1e-3, 0.5, 1.0, 2.0, 4.99, 5.0, 5.99, 5.999999, 6.0,
                        6.000001, 10.0, 100.0, 1e4, 1e6, 1e12])
    grid = np.concatenate([special, np.exp(rs.uniform(np.log(1e-8), np.log(1e5), size=493))])
    # psi~(x) alone is not exported; DE on rows [x, big] gives psi~(x) - psi~(x+big); store whole rows instead
    rows = np.stack([grid, np.roll(grid, 1), np.roll(grid, 7), np.full_like(grid, 3.25)], axis=1)
    de_rows = cutils._dirichlet_expectation_2d(rows)
    wide = rs.gamma(100.0, 0.01, size=(6, 257))
    de_wide = cutils._dirichlet_expectation_2d(wide)
    wide32 = wide.astype(np.float32)
    de_wide32 = cutils._dirichlet_expectation_2d(wide32)
    d1 = rs.gamma(2.0, 1.0, size=33)
    d1_io = d1.copy()
    out1 = np.empty_like(d1)
    cutils._dirichlet_expectation_1d(d1_io, 0.37, out1)
    save("psi_as103.npz", rows=rows, de_rows=de_rows, wide=wide, de_wide=de_wide, de_wide32=de_wide32,
         d1_in=d1, d1_after=d1_io, d1_prior=np.float64(0.37), d1_out=out1)

    # ---- B. README config (cfg1): doc_generator M=100 L=20 -> Document -> default fit -----------------------
    import torch

    torch.manual_seed(1234)
    gen = generator.doc_generator(M=100, L=20, topic_prior=torch.tensor([1, 1, 1, 1, 1], dtype=torch.double))
    docs = gen.generate_doc()
    doc = document.Document(docs)
    with open(os.path.join(HERE, "cfg1_docs.json"), "w") as f:
        json.dump({"docs": {str(k): v for k, v in docs.items()}, "vocab_to_idx": doc.vocab_to_idx,
                   "sample_docs_list": doc.sample_docs_list}, f)
    Xd = np.ascontiguousarray(doc.doc_vocab_count_array)
    lda = lda_model.LDASmoothed(num_topics=5, verbose=False)
    perps = lda.fit(doc, return_perplexities=True)
    save("fit_cfg1.npz", **csr_of(Xd), X=Xd.astype(np.int32), K=np.int64(5),
         perplexities=np.array(perps), gamma=lda._gamma_, lam=lda._lambda_,
         alpha=np.float64(lda._alpha_), eta=np.float64(lda._eta_))

    # ---- B2. Document on free text (stop words, punctuation, accents, case, empty doc) ----------------------
    texts = {
        "a": "The Quick brown fox, jumps over the lazy dog! FIFA and Olympic are not the same.",
        "b": "Café déjà-vu: the fox is NOT lazy;  the dog is...   quick quick quick",
        "c": "",
        "d": "numbers 123 and symbols #$% stay? fox_dog under_score the THE tHe",
        7: "Olympic FIFA olympic fifa dog dog fox",
    }
    tdoc = document.Document(texts)
    with open(os.path.join(HERE, "document_vocab.json"), "w") as f:
        json.dump({"texts": {str(k): v for k, v in texts.items()}, "vocab_to_idx": tdoc.vocab_to_idx,
                   "sample_docs_list": tdoc.sample_docs_list,
                   "doc_vocab_count_array": np.asarray(tdoc.doc_vocab_count_array).tolist(),
                   "M": tdoc.M, "V": tdoc.V, "num_doc_population": tdoc.num_doc_population}, f)

    # ---- C. small synthetic, hyper update on ------------------------------------------------------------------
    X = synth_dense(11, D=200, V=300, K=8, L=60)
    save("fit_small.npz", **csr_of(X), K=np.int64(8), kw=json.dumps(dict(em_num_step=6)),
         **run_fit(X, 8, em_num_step=6))

    # one e_step / m_step / elbo in isolation, from the initial state of case C
    lda = lda_model.LDASmoothed(num_topics=8, verbose=False)
    lda.fit(DenseCorpus(X), em_num_step=0, update_hyper_parms=False)
    g0, l0 = lda._gamma_.copy(), lda._lambda_.copy()
    elbo0, (Et0, Eb0) = lda.approx_elbo(X)
    g1, (S1, Et1) = lda.e_step(X, Et0, Eb0)
    g1 = g1.copy()
    l1, Eb1 = lda.m_step(S1)
    elbo1, _ = lda.approx_elbo(X, False, Et1, Eb1)
    save("estep_small.npz", **csr_of(X), K=np.int64(8), gamma0=g0, lam0=l0, elbo0=np.float64(elbo0),
         Elogtheta0=Et0, Elogbeta0=Eb0, gamma1=g1, sstats1=S1, Elogtheta1=Et1, lam1=l1, Elogbeta1=Eb1,
         elbo1=np.float64(elbo1))

    # ---- D. non-exchangeable alpha (1,K) array, no hyper update ---------------------------------------------
    X = synth_dense(12, D=120, V=200, K=6, L=40)
    save("fit_nonexch.npz", **csr_of(X), K=np.int64(6),
         kw=json.dumps(dict(em_num_step=4, update_hyper_parms=False)),
         **run_fit(X, 6, exchangeable=False, em_num_step=4, update_hyper_parms=False))

    # ---- D2. non-exchangeable update_alpha, first Newton iterations ------------------------------------------
    save("hyper_nonexch.npz", **csr_of(X), K=np.int64(6), **run_nonexch_hyper(X, 6, 3))

    # ---- E. sampling=True with D_population injected ---------------------------------------------------------
    save("fit_sampling.npz", **csr_of(X), K=np.int64(6), D_population=np.int64(1000),
         kw=json.dumps(dict(em_num_step=3, sampling=True)),
         **run_fit(X, 6, D_population=1000, em_num_step=3, sampling=True))

    # ---- F. early return (large threshold): no hyper update, short trace -------------------------------------
    save("fit_early.npz", **csr_of(X), K=np.int64(6), kw=json.dumps(dict(threshold=5.0, em_num_step=50)),
         **run_fit(X, 6, threshold=5.0, em_num_step=50))

    # ---- G. ragged corpus with empty documents and a never-used vocabulary tail ------------------------------
    X = synth_dense(13, D=90, V=150, K=5, L=(25,))
    X[3, :] = 0
    X[40, :] = 0
    X[89, :] = 0
    X[:, 140:] = 0
    X[17, 5] = 250  # one heavy count
    save("fit_ragged.npz", **csr_of(X), K=np.int64(5), kw=json.dumps(dict(em_num_step=5)),
         **run_fit(X, 5, em_num_step=5))

    # ---- H. cfg2-shaped slice (K=20, V=5k) — 2 EM iterations of the real reference --------------------------
    X = synth_dense(14, D=300, V=5000, K=20, L=200)
    r = run_fit(X, 20, em_num_step=2, update_hyper_parms=False)
    save("fit_cfg2_slice.npz", **csr_of(X), K=np.int64(20),
         kw=json.dumps(dict(em_num_step=2, update_hyper_parms=False)), **r)


if __name__ == "__main__":
    main()
