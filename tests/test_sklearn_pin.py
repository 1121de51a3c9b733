"""External pin of the two modes the reference does not define: the TRUE per-document fixed point
(`inner_fixed_point=True`) and online variational Bayes (`partial_fit`, §8f row 4; the reference
only lists it as roadmap, readme.md:131-136).  Both are what scikit-learn's LatentDirichletAllocation implements
(Hoffman, Blei & Bach 2010), and scikit-learn 1.9 is in the image — so they are pinned to IT, not to a restatement
written for this repo:

* E-step:  sklearn.decomposition._lda._update_doc_distribution (random_state=None: gamma starts at 1)
* online M-step:  LatentDirichletAllocation._em_step(batch_update=False) (with its E-step's random gamma start
  switched off — `random_init=False` — since a random start cannot be reproduced by another implementation).

What makes a tight comparison possible: scikit-learn's Cython helpers (`_online_lda_fast.pyx`: psi, the two
`_dirichlet_expectation_*`, mean_change) are the code the reference's cutils.pyx was taken from — the SAME AS-103
digamma, so the 2.4e-9 AS-103-vs-exact gap does not enter.  The two remaining, accounted-for differences:
  1. the normaliser's guard is `+ eps` = 2.2e-16 there and `+ 1e-100` here (tuotuo/lda_model.py:242): relative effect
     eps / (theta.beta) <= ~1e-11 on these corpora (theta.beta >= ~1e-5);
  2. scikit-learn stops on mean|dgamma| < tol, the reference's loop on ||dgamma||_1 <= 1e-5 (:236): identical
     iteration sequences with tol = 1e-5 / K (up to an exact tie);
  3. the iteration caps differ by one (`range(max_iter)` there, `it <= step` here, :236): the tests give both a cap
     (1000) that no document reaches — started from gamma = 1 a document needs ~200 passes.
Gate: 1e-9 relative on gamma / lambda / statistics, the north-star tolerance.

CPU part (not gpu): the ORACLE's restatement of both modes against scikit-learn — which pins what the GPU tests in
test_gpu_parity.py compare the kernels with.  GPU part: the kernels against scikit-learn directly.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from lda_oracle import CSRCorpus, OracleLDA, dirichlet_expectation_2d

sk_lda = pytest.importorskip("sklearn.decomposition._lda")
from sklearn.decomposition import LatentDirichletAllocation  # noqa: E402

TOL = 1e-9


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def corpus(seed, D, V, K, L):
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(seed, D, V, min(K, 12), L, ragged=True)
    X = sp.csr_matrix((ct.astype(np.float64), ci, rp), shape=(D, V))
    return (rp, ci, ct), X


def lambda0(K, V, seed=3):
    return np.random.RandomState(seed).gamma(100.0, 0.01, (K, V))


def sklearn_estep(X, lam, alpha, K, tol_l1=1e-5, max_iter=1000):
    exp_topic_word = np.exp(sk_lda._dirichlet_expectation_2d(lam))
    return sk_lda._update_doc_distribution(X, exp_topic_word, alpha, max_iter, tol_l1 / K, True, None)


CASES = [(10, 300, 80, 60), (50, 900, 60, 90), (200, 700, 24, 120)]


@pytest.mark.parametrize("K,V,D,L", CASES)
@pytest.mark.parametrize("engine", ["c", "numpy"])
def test_oracle_fixed_point_estep_is_sklearns(K, V, D, L, engine):
    (rp, ci, ct), X = corpus(40 + K, D, V, K, L)
    lam = lambda0(K, V)
    g_sk, s_sk = sklearn_estep(X, lam, 1.0, K)
    o = OracleLDA(K, engine=engine, inner_fixed_point=True)
    o.gamma = np.ones((D, K))
    Elogbeta = dirichlet_expectation_2d(lam, engine)
    gamma, (S, _) = o.e_step(CSRCorpus(rp, ci, ct, V), dirichlet_expectation_2d(o.gamma, engine), Elogbeta, 1000, 1e-5)
    assert o.capped_docs == 0
    assert o.inner_iters > 3 * D  # a real fixed-point iteration, not the reference's two frozen passes
    assert relerr(gamma, g_sk) < TOL
    assert relerr(S, s_sk * np.exp(Elogbeta)) < TOL  # the oracle returns sstats * exp(Elogbeta) (lda_model.py:264)


def _sklearn_online(K, V, total, offset, decay, lam):
    m = LatentDirichletAllocation(n_components=K, doc_topic_prior=1.0, topic_word_prior=1.0, learning_method="online",
                                  learning_offset=offset, learning_decay=decay, total_samples=total,
                                  max_doc_update_iter=1000, mean_change_tol=1e-5 / K, random_state=0)
    m._init_latent_vars(V, dtype=np.float64)
    m.components_ = lam.copy()
    m.exp_dirichlet_component_ = np.exp(sk_lda._dirichlet_expectation_2d(m.components_))
    orig = m._e_step
    m._e_step = lambda X, cal_sstats, random_init, parallel=None: orig(X, cal_sstats=cal_sstats, random_init=False,
                                                                       parallel=parallel)
    return m


@pytest.mark.parametrize("K,V", [(12, 300), (60, 500)])
def test_oracle_partial_fit_is_sklearns_online_step(K, V):
    total, offset, decay = 4000, 2.0, 0.6
    lam = lambda0(K, V, seed=9)
    sk = _sklearn_online(K, V, total, offset, decay, lam)
    o = OracleLDA(K, inner_fixed_point=True)
    o.lam, o.t_online = lam.copy(), 0
    for i, D in enumerate((70, 40, 95)):
        (rp, ci, ct), X = corpus(200 + i, D, V, K, 50)
        sk._em_step(X, total_samples=total, batch_update=False, parallel=None)
        # n_batch_iter_ starts at 1 there, t at 0 here: tau0 = learning_offset + 1
        o.partial_fit(CSRCorpus(rp, ci, ct, V), total_docs=total, tau0=offset + 1.0, kappa=decay, gamma_init=1.0,
                      e_num_step=1000)
        assert relerr(o.lam, sk.components_) < TOL, i
    g_sk, _ = sk._e_step(X, cal_sstats=False, random_init=False)
    assert g_sk.shape == o.gamma.shape  # (and the last minibatch's gamma under the topics BEFORE the last update:)
    # o.gamma was computed before the last M-step; recompute sklearn's the same way
    sk2 = _sklearn_online(K, V, total, offset, decay, lam)
    o2 = OracleLDA(K, inner_fixed_point=True)
    o2.lam, o2.t_online = lam.copy(), 0
    (rp, ci, ct), X = corpus(200, 70, V, K, 50)
    g2, _ = sk2._e_step(X, cal_sstats=False, random_init=False)
    o2.partial_fit(CSRCorpus(rp, ci, ct, V), total_docs=total, tau0=offset + 1.0, kappa=decay, gamma_init=1.0,
                      e_num_step=1000)
    assert relerr(o2.gamma, g2) < TOL


# ------------------------------------------------------------------------------------------------------------------
# the kernels against scikit-learn directly
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("K,V,D,L", CASES + [(1000, 1100, 10, 100)])
def test_gpu_fixed_point_estep_is_sklearns(K, V, D, L):
    """ttlda_estep_fixed_point_f64 + the statistics pass, through the reference-shaped e_step() call, started from
    gamma = 1 and scikit-learn's exp(E[log beta])."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    (rp, ci, ct), X = corpus(40 + K, D, V, K, L)
    lam = lambda0(K, V)
    g_sk, s_sk = sklearn_estep(X, lam, 1.0, K)
    m = LDASmoothed(K, verbose=False, inner_fixed_point=True)
    m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=0, update_hyper_parms=False)
    Elogbeta = dirichlet_expectation_2d(lam)
    gamma, (upd, _) = m.e_step(None, dirichlet_expectation_2d(np.ones((D, K))), Elogbeta, step=1000, threshold=1e-5)
    assert m.inner_iterations > 3 * D
    assert relerr(gamma, g_sk) < TOL
    assert relerr(upd, s_sk * np.exp(Elogbeta)) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("K,V", [(12, 300), (100, 800)])
def test_gpu_partial_fit_is_sklearns_online_step(K, V):
    """LDASmoothed.partial_fit (ttlda_estep_fixed_point_f64, statistics, ttlda_mstep_online_f64) == three consecutive
    LatentDirichletAllocation._em_step(batch_update=False) calls from the same lambda."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    total, offset, decay = 4000, 2.0, 0.6
    m = LDASmoothed(K, verbose=False)
    # the model's own lambda_0 (seed 42, Gamma(100, .01)) — lda_model.py:356-357 — handed to scikit-learn
    np.random.seed(42)
    lam = np.random.gamma(shape=100, scale=0.01, size=(K, V))
    sk = _sklearn_online(K, V, total, offset, decay, lam)
    for i, D in enumerate((70, 40, 95)):
        (rp, ci, ct), X = corpus(200 + i, D, V, K, 50)
        g_sk, _ = sk._e_step(X, cal_sstats=False, random_init=False)
        sk._em_step(X, total_samples=total, batch_update=False, parallel=None)
        m.partial_fit(Document.from_csr(rp, ci, ct, V), total_docs=total, tau0=offset + 1.0, kappa=decay, gamma_init=1.0,
                      e_num_step=1000)
        assert relerr(m._gamma_, g_sk) < TOL, i
        assert relerr(m._lambda_, sk.components_) < TOL, i
