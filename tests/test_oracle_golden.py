"""The oracle (oracle/lda_oracle.{c,py}) against the golden fixtures produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os
import warnings

import numpy as np
import pytest

import lda_oracle as orc
from lda_oracle import CSRCorpus, OracleLDA


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _corpus(g):
    return CSRCorpus(g["rowptr"], g["colidx"], g["counts"], int(g["V"]))


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("engine", ["c", "numpy"])
def test_dirichlet_expectation_known_answers(golden_dir, engine):
    g = _load(golden_dir, "psi_as103.npz")
    # AS-103 digamma, op-for-op: allow only last-bit differences (libm log vs itself: identical here)
    assert np.max(np.abs(orc.dirichlet_expectation_2d(g["rows"], engine) - g["de_rows"])) <= 1e-9 * 0 + 4e-16 * np.max(np.abs(g["de_rows"]))
    assert relerr(orc.dirichlet_expectation_2d(g["wide"], engine), g["de_wide"]) < 1e-13
    d1 = g["d1_in"].copy()
    out = np.empty_like(d1)
    orc.dirichlet_expectation_1d(d1, float(g["d1_prior"]), out, engine)
    np.testing.assert_allclose(d1, g["d1_after"], rtol=0, atol=0)
    assert relerr(out, g["d1_out"]) < 1e-14


def test_psi_as103_is_not_exact_digamma(golden_dir):
    """The reference's psi is the approximate AS-103 (abs error up to ~2.4e-9): the oracle must reproduce the
    approximation, not scipy's exact digamma."""
    from scipy.special import psi

    x = np.linspace(0.5, 6.5, 2001)
    err = np.abs(orc.psi_as103(x) - psi(x))
    assert 1e-9 < err.max() < 3e-9
    c = np.array([orc._lib().ttor_psi_as103(float(v)) for v in x])
    assert np.max(np.abs(c - orc.psi_as103(x))) < 1e-15


@pytest.mark.parametrize("engine", ["c", "numpy"])
def test_estep_mstep_elbo_single_iteration(golden_dir, engine):
    g = _load(golden_dir, "estep_small.npz")
    X = _corpus(g)
    m = OracleLDA(int(g["K"]), engine=engine)
    m.M, m.V = X.D, X.V
    m.gamma = g["gamma0"].copy()
    m.lam = g["lam0"].copy()
    elbo0, (Et0, Eb0) = m.approx_elbo(X)
    assert relerr(elbo0, g["elbo0"]) < 1e-13
    assert relerr(Et0, g["Elogtheta0"]) < 1e-12 and relerr(Eb0, g["Elogbeta0"]) < 1e-12
    gam, (S, Et1) = m.e_step(X, Et0, Eb0)
    assert relerr(gam, g["gamma1"]) < 1e-13
    assert relerr(S, g["sstats1"]) < 1e-12
    assert relerr(Et1, g["Elogtheta1"]) < 1e-11
    lam, Eb1 = m.m_step(S)
    assert relerr(lam, g["lam1"]) < 1e-13
    elbo1, _ = m.approx_elbo(X, False, Et1, Eb1)
    assert relerr(elbo1, g["elbo1"]) < 1e-13
    assert m.inner_iters == 2 * X.D  # the frozen-theta loop runs exactly twice per document


FITS = ["fit_small.npz", "fit_nonexch.npz", "fit_sampling.npz", "fit_early.npz", "fit_ragged.npz"]


@pytest.mark.parametrize("name", FITS)
@pytest.mark.parametrize("engine", ["c", "numpy"])
def test_fit_traces(golden_dir, name, engine):
    g = _load(golden_dir, name)
    X = _corpus(g)
    kw = json.loads(str(g["kw"]))
    exch = g["alpha0"].ndim == 0
    m = OracleLDA(int(g["K"]), exchangeable_prior=exch, engine=engine)
    if not exch:
        np.testing.assert_array_equal(np.asarray(m.alpha), g["alpha0"])  # same legacy RNG stream
    if "D_population" in g:
        m.D_population = int(g["D_population"])
    perps = m.fit(X, **kw)
    assert len(perps) == len(g["perplexities"])
    assert relerr(perps, g["perplexities"]) < 1e-11
    assert relerr(m.gamma, g["gamma"]) < 1e-11
    assert relerr(m.lam, g["lam"]) < 1e-11
    assert relerr(m.alpha, g["alpha"]) < 1e-11
    assert relerr(m.eta, g["eta"]) < 1e-11
    assert m.inner_iters == int(g["inner_iters"])


def test_fit_cfg1_readme(golden_dir):
    """README config: 100 docs x 20 tokens from doc_generator, K=5, default fit (100 iterations + hyper update)."""
    g = _load(golden_dir, "fit_cfg1.npz")
    X = _corpus(g)
    np.testing.assert_array_equal(X.to_dense(), g["X"])
    m = OracleLDA(5, engine="c")
    perps = m.fit(X)
    assert len(perps) == 102
    assert relerr(perps, g["perplexities"]) < 1e-10
    assert relerr(m.gamma, g["gamma"]) < 1e-9
    assert relerr(m.lam, g["lam"]) < 1e-9
    assert relerr(m.alpha, g["alpha"]) < 1e-10 and relerr(m.eta, g["eta"]) < 1e-10


def test_fit_cfg2_slice(golden_dir):
    g = _load(golden_dir, "fit_cfg2_slice.npz")
    X = _corpus(g)
    m = OracleLDA(20, engine="c")
    perps = m.fit(X, **json.loads(str(g["kw"])))
    assert relerr(perps, g["perplexities"]) < 1e-12
    assert relerr(m.gamma, g["gamma"]) < 1e-12 and relerr(m.lam, g["lam"]) < 1e-12


def test_vocab_order_and_counts(golden_dir):
    """Vocabulary ids = order of first appearance; counts equal the reference's dense matrix (notebook cells 13-17)."""
    with open(os.path.join(golden_dir, "cfg1_docs.json")) as f:
        j = json.load(f)
    vocab, ids, offs = orc.build_vocab(j["sample_docs_list"])
    assert vocab == j["vocab_to_idx"] and list(vocab) == list(j["vocab_to_idx"])
    g = _load(golden_dir, "fit_cfg1.npz")
    X = CSRCorpus.from_token_ids(ids, offs, len(vocab))
    np.testing.assert_array_equal(X.rowptr, g["rowptr"])
    np.testing.assert_array_equal(X.colidx, g["colidx"])
    np.testing.assert_array_equal(X.counts, g["counts"])


def test_estep_iteration_cap_warns_like_reference(golden_dir):
    """step=0: the loop body runs once, `it > step` holds, every document is reported capped (lda_model.py:258)."""
    g = _load(golden_dir, "estep_small.npz")
    X = _corpus(g)
    for engine in ("c", "numpy"):
        m = OracleLDA(int(g["K"]), engine=engine)
        m.gamma, m.lam = g["gamma0"].copy(), g["lam0"].copy()
        gam, _ = m.e_step(X, g["Elogtheta0"], g["Elogbeta0"], step=0)
        assert m.capped_docs == X.D and m.inner_iters == X.D
        assert relerr(gam, g["gamma1"]) < 1e-13


def test_online_step_reduces_to_batch_step():
    """OracleLDA.partial_fit (online VB, not in the reference) with rho = 1 and |B| = D is one fixed-point EM step."""
    import lda_oracle as orc

    rng = np.random.default_rng(5)
    X = orc.CSRCorpus.from_dense(rng.poisson(0.3, size=(40, 30)).astype(np.float64))
    a = orc.OracleLDA(4, inner_fixed_point=True)
    a.partial_fit(X, total_docs=X.D, tau0=1.0, kappa=0.7)  # rho_0 = 1
    b = orc.OracleLDA(4, inner_fixed_point=True)
    b.fit(X, em_num_step=1, update_hyper_parms=False)
    assert np.max(np.abs(a.lam - b.lam) / np.abs(b.lam)) < 1e-13
    assert np.max(np.abs(a.gamma - b.gamma) / np.abs(b.gamma)) < 1e-13
    lam1 = a.lam.copy()
    a.partial_fit(X, total_docs=4 * X.D, tau0=1.0, kappa=0.7)  # rho_1 = 2^-0.7: a proper blend
    assert a.t_online == 2 and not np.allclose(a.lam, lam1)


def test_nonexchangeable_update_alpha_first_iterations(golden_dir):
    """The reference's non-exchangeable update_alpha (broken upstream: it diverges at its cap) — its first Newton
    iterations, pinned against the real reference."""
    import lda_oracle as orc

    z = np.load(os.path.join(golden_dir, "hyper_nonexch.npz"))
    X = orc.CSRCorpus(z["rowptr"], z["colidx"], z["counts"], int(z["V"]))
    o = orc.OracleLDA(int(z["K"]), exchangeable_prior=False)
    o.fit(X, em_num_step=2, update_hyper_parms=False)
    assert np.max(np.abs(o.alpha - z["alpha0"])) == 0
    assert np.max(np.abs(o.gamma - z["gamma"]) / z["gamma"]) < 1e-11
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        o.update_alpha(step=int(z["newton_steps"]))
    assert len(w) == int(z["n_warnings"])
    assert o.alpha.shape == z["alpha"].shape
    assert np.max(np.abs(o.alpha - z["alpha"]) / np.abs(z["alpha"])) < 1e-10
