"""oracle/lda_oracle.py — CPU oracle for TuoTuo's batch-VI LDA hot path (``LDASmoothed.fit``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, never by the product package ``tuotuo_b200``.

Parity pinning: the reference has no tests or golden vectors.  This oracle is pinned against
the reference itself, executed in the build container: ``tests/golden/*.npz`` hold traces of the real
``tuotuo.lda_model.LDASmoothed.fit`` / ``tuotuo.cutils`` / ``tuotuo.document.Document`` produced by
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks every function here against them.

Two engines restate the same algorithm (all citations relative to ``/root/reference/``):

* ``engine="numpy"`` — a literal per-document NumPy/SciPy restatement (loops kept, SciPy ``logsumexp``,
  ``gammaln``, ``psi``, ``polygamma`` exactly where the reference calls them).  Small cases only.
* ``engine="c"``     — ``oracle/lda_oracle.c`` (gcc + OpenMP) through ctypes.  Scales to bench samples; it is
  the timed ``cpu_baseline`` ("port").

Layouts follow the reference: gamma (D,K), lambda/Elogbeta (K,V), corpus as CSR with ascending word ids per
row (what ``np.nonzero(X[d,:])`` yields, ``tuotuo/lda_model.py:226``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import warnings
from dataclasses import dataclass, field

import numpy as np
from scipy.special import gammaln, logsumexp, polygamma, psi

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "liblda_oracle.so")
SRC_PATH = os.path.join(HERE, "lda_oracle.c")

EULER = 0.577215664901532860606512090082402431
SEED = 42  # tuotuo/lda_model.py:19


# ----------------------------------------------------------------------------------------------------------
# C engine
# ----------------------------------------------------------------------------------------------------------
def build_c_oracle(force: bool = False) -> str:
    """gcc -O2 -fopenmp oracle/lda_oracle.c -> oracle/_build/liblda_oracle.so"""
    if (not force and os.path.exists(LIB_PATH)
            and (not os.path.exists(SRC_PATH) or os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC_PATH))):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
         SRC_PATH, "-o", LIB_PATH, "-lm"]
    )
    return LIB_PATH


_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        dp = ctypes.POINTER(ctypes.c_double)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64 = ctypes.c_int64
        lib.ttor_psi_as103.restype = ctypes.c_double
        lib.ttor_psi_as103.argtypes = [ctypes.c_double]
        lib.ttor_dirichlet_expectation_2d.restype = None
        lib.ttor_dirichlet_expectation_2d.argtypes = [dp, i64, i64, dp]
        lib.ttor_dirichlet_expectation_1d.restype = None
        lib.ttor_dirichlet_expectation_1d.argtypes = [dp, ctypes.c_double, dp, i64]
        lib.ttor_e_step.restype = i64
        lib.ttor_e_step.argtypes = [i64p, i32p, i32p, i64, i64, i64, dp, dp, dp, dp, dp, dp, i64,
                                    ctypes.c_double, i64p, ctypes.c_int]
        lib.ttor_m_step.restype = None
        lib.ttor_m_step.argtypes = [dp, ctypes.c_double, i64, i64, dp, dp]
        lib.ttor_elbo.restype = ctypes.c_double
        lib.ttor_elbo.argtypes = [i64p, i32p, i32p, i64, i64, i64, dp, dp, dp, dp, dp, ctypes.c_double,
                                  ctypes.c_double, ctypes.c_double, dp]
        lib.ttor_num_threads.restype = ctypes.c_int
        lib.ttor_set_num_threads.argtypes = [ctypes.c_int]
        _LIB = lib
    return _LIB


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def c_num_threads() -> int:
    return int(_lib().ttor_num_threads())


def c_set_num_threads(n: int) -> None:
    _lib().ttor_set_num_threads(int(n))


# ----------------------------------------------------------------------------------------------------------
# AS-103 digamma and Dirichlet expectation (tuotuo/cutils.pyx)
# ----------------------------------------------------------------------------------------------------------
def psi_as103(x) -> np.ndarray:
    """Vectorised NumPy restatement of ``psi`` (tuotuo/cutils.pyx:75-93), same operation order."""
    x = np.array(x, dtype=np.float64, copy=True)
    out = np.zeros_like(x)
    small = x <= 1e-6  # cutils.pyx:76-78
    with np.errstate(divide="ignore", invalid="ignore"):
        small_val = -EULER - 1.0 / x
        xs = np.where(small, 10.0, x)
        while True:  # cutils.pyx:83-85
            m = xs < 6
            if not m.any():
                break
            out = np.where(m, out - 1.0 / xs, out)
            xs = np.where(m, xs + 1, xs)
        r = 1.0 / xs  # cutils.pyx:89-92
        out = out + (np.log(xs) - 0.5 * r)
        r = r * r
        out = out - r * ((1.0 / 12.0) - r * ((1.0 / 120.0) - r * (1.0 / 252.0)))
    return np.where(small, small_val, out)


def dirichlet_expectation_2d(arr: np.ndarray, engine: str = "c") -> np.ndarray:
    """``_dirichlet_expectation_2d`` (tuotuo/cutils.pyx:41-68): psi~(a_ij) - psi~(sum_j a_ij), serial row sum."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    assert arr.ndim == 2
    if engine == "c":
        out = np.empty_like(arr)
        _lib().ttor_dirichlet_expectation_2d(_p(arr, ctypes.c_double), arr.shape[0], arr.shape[1],
                                             _p(out, ctypes.c_double))
        return out
    # serial left-to-right accumulation like cutils.pyx:60-62 (np.cumsum accumulates sequentially)
    tot = np.cumsum(arr, axis=1)[:, -1] if arr.shape[1] else np.zeros(arr.shape[0])
    return psi_as103(arr) - psi_as103(tot)[:, None]


def dirichlet_expectation_1d(doc_topic: np.ndarray, prior: float, out: np.ndarray, engine: str = "c") -> None:
    """``_dirichlet_expectation_1d`` (tuotuo/cutils.pyx:11-38): in-place ``doc_topic += prior``; out = exp(E[log])."""
    assert doc_topic.dtype == np.float64 and out.dtype == np.float64
    if engine == "c":
        _lib().ttor_dirichlet_expectation_1d(_p(doc_topic, ctypes.c_double), float(prior),
                                             _p(out, ctypes.c_double), doc_topic.shape[0])
        return
    total = 0.0
    for i in range(doc_topic.shape[0]):
        doc_topic[i] = doc_topic[i] + prior
        total += doc_topic[i]
    out[:] = np.exp(psi_as103(doc_topic) - psi_as103(total))


def np_clip_for_exp(x, sub: float = 1e-05):
    """tuotuo/utils.py:20-26."""
    if x > 0:
        x = max(sub, x)
    elif x < 0:
        x = min(sub, x)
    return x


# ----------------------------------------------------------------------------------------------------------
# corpus helpers
# ----------------------------------------------------------------------------------------------------------
@dataclass
class CSRCorpus:
    """Documents as CSR: rowptr int64 (D+1), colidx int32 (ascending per row), counts int32."""
    rowptr: np.ndarray
    colidx: np.ndarray
    counts: np.ndarray
    V: int

    @property
    def D(self) -> int:
        return len(self.rowptr) - 1

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])

    @property
    def tokens(self) -> int:
        return int(self.counts.sum(dtype=np.int64))

    @staticmethod
    def from_dense(X: np.ndarray) -> "CSRCorpus":
        X = np.asarray(X)
        D, V = X.shape
        rows, cols = np.nonzero(X)  # row-major order => ascending cols within a row
        counts = X[rows, cols]
        rowptr = np.zeros(D + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        rowptr = np.cumsum(rowptr)
        return CSRCorpus(rowptr.astype(np.int64), cols.astype(np.int32), counts.astype(np.int32), int(V))

    def to_dense(self) -> np.ndarray:
        X = np.zeros((self.D, self.V), dtype=np.float64)
        for d in range(self.D):
            b, e = self.rowptr[d], self.rowptr[d + 1]
            X[d, self.colidx[b:e]] = self.counts[b:e]
        return X

    @staticmethod
    def from_token_ids(token_ids: np.ndarray, doc_offsets: np.ndarray, V: int) -> "CSRCorpus":
        """Token stream -> per-document sorted unique (word, count): the bag-of-words of utils.py:177-206."""
        D = len(doc_offsets) - 1
        rp = [0]
        ci, ct = [], []
        for d in range(D):
            u, c = np.unique(token_ids[doc_offsets[d]:doc_offsets[d + 1]], return_counts=True)
            ci.append(u)
            ct.append(c)
            rp.append(rp[-1] + len(u))
        return CSRCorpus(np.asarray(rp, dtype=np.int64),
                         (np.concatenate(ci) if ci else np.zeros(0)).astype(np.int32),
                         (np.concatenate(ct) if ct else np.zeros(0)).astype(np.int32), int(V))


def build_vocab(docs_tokens):
    """Vocabulary ids by order of first appearance (tuotuo/utils.py:177-206: dict insertion order of
    get_vocab_from_docs, enumerated by get_np_wct).  Returns (vocab_to_idx, token_ids, doc_offsets)."""
    vocab = {}
    ids = []
    offs = [0]
    for doc in docs_tokens:
        for w in doc:
            j = vocab.get(w)
            if j is None:
                j = len(vocab)
                vocab[w] = j
            ids.append(j)
        offs.append(len(ids))
    return vocab, np.asarray(ids, dtype=np.int32), np.asarray(offs, dtype=np.int64)


# ----------------------------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------------------------
@dataclass
class OracleLDA:
    """Restatement of ``LDASmoothed`` (tuotuo/lda_model.py:49-564) over a CSR corpus."""
    K: int
    exchangeable_prior: bool = True
    engine: str = "c"
    inner_fixed_point: bool = False  # NOT the reference: restates the CUDA library's opt-in true-fixed-point mode
    alpha: object = field(init=False)
    eta: object = field(init=False)

    def __post_init__(self):
        np.random.seed(SEED)  # lda_model.py:71
        if self.exchangeable_prior:
            self.alpha = 1  # :73
        else:
            self.alpha = np.random.gamma(shape=100, scale=0.01, size=(1, self.K))  # :75
        self.eta = 1  # :83
        self.D_population = None
        self.inner_iters = 0
        self.capped_docs = 0

    # -- helpers -------------------------------------------------------------------------------------------
    def _alpha_vec(self) -> np.ndarray:
        return np.ascontiguousarray(np.broadcast_to(np.asarray(self.alpha, dtype=np.float64).reshape(-1), (self.K,)))

    def _alpha_sum(self) -> float:
        # np.sum(real_dist_prior) at lda_model.py:43: scalar prior -> itself; (1,K) array -> sum of entries
        return float(np.sum(self.alpha))

    # -- lda_model.py:25-46 --------------------------------------------------------------------------------
    @staticmethod
    def _prior_term_numpy(E, p, q):
        t = np.sum((p - q) * E)
        t += np.sum(gammaln(q) - gammaln(p))
        t += np.sum(gammaln(np.sum(p)) - gammaln(np.sum(q, 1)))
        return t

    # -- lda_model.py:89-165 -------------------------------------------------------------------------------
    def approx_elbo(self, X: CSRCorpus, sampling=False, Elogtheta=None, Elogbeta=None):
        if Elogtheta is None:
            Elogtheta = dirichlet_expectation_2d(self.gamma, self.engine)  # :116
        if Elogbeta is None:
            Elogbeta = dirichlet_expectation_2d(self.lam, self.engine)  # :120
        D = X.D
        scale = (self.D_population / D) if sampling else 1.0  # :152-153 (AttributeError in the reference if unset)
        if self.engine == "c":
            parts = np.zeros(3)
            av = self._alpha_vec()
            elbo = _lib().ttor_elbo(
                _p(X.rowptr, ctypes.c_int64), _p(X.colidx, ctypes.c_int32), _p(X.counts, ctypes.c_int32),
                D, self.K, X.V, _p(self.gamma, ctypes.c_double), _p(self.lam, ctypes.c_double),
                _p(Elogtheta, ctypes.c_double), _p(Elogbeta, ctypes.c_double),
                _p(av, ctypes.c_double), self._alpha_sum(), float(self.eta), float(scale),
                _p(parts, ctypes.c_double))
            self.last_parts = parts
            self.last_elbo = float(elbo)
            return float(elbo), (Elogtheta, Elogbeta)
        elbo = 0
        for d in range(D):  # :125-139
            b, e = X.rowptr[d], X.rowptr[d + 1]
            ids = X.colidx[b:e]
            cts = X.counts[b:e].astype(np.float64)
            lps = np.zeros(len(ids))
            for i in range(len(ids)):
                lps[i] = logsumexp(Elogtheta[d, :] + Elogbeta[:, ids[i]])
            elbo += np.dot(cts, lps)
        elbo += self._prior_term_numpy(Elogtheta, self.alpha, self.gamma)  # :142-148
        if sampling:
            elbo = elbo * self.D_population / D
        elbo += self._prior_term_numpy(Elogbeta, self.eta, self.lam)  # :157-163
        self.last_elbo = float(elbo)
        return float(elbo), (Elogtheta, Elogbeta)

    # -- lda_model.py:167-194 ------------------------------------------------------------------------------
    def approx_perplexity(self, X: CSRCorpus, sampling=False, Elogtheta=None, Elogbeta=None):
        elbo, logs = self.approx_elbo(X, sampling, Elogtheta, Elogbeta)
        total = np.float64(X.tokens)  # np.sum(X) in the reference: a zero token count yields inf/nan, not an exception
        if sampling:
            total = total * self.D_population / X.D  # :188
        temp = -elbo / total
        return float(np.exp(np_clip_for_exp(temp))), logs

    # -- lda_model.py:197-270 ------------------------------------------------------------------------------
    def e_step(self, X: CSRCorpus, Elogtheta, Elogbeta, step=100, threshold=1e-05):
        K, V, D = self.K, X.V, X.D
        if self.engine == "c":
            eB = np.exp(Elogbeta)
            S = np.empty((K, V))
            Et_out = np.empty((D, K))
            iters = ctypes.c_int64(0)
            av = self._alpha_vec()
            capped = _lib().ttor_e_step(
                _p(X.rowptr, ctypes.c_int64), _p(X.colidx, ctypes.c_int32), _p(X.counts, ctypes.c_int32),
                D, K, V, _p(av, ctypes.c_double), _p(np.ascontiguousarray(Elogtheta), ctypes.c_double),
                _p(eB, ctypes.c_double), _p(self.gamma, ctypes.c_double), _p(S, ctypes.c_double),
                _p(Et_out, ctypes.c_double), int(step), float(threshold), ctypes.byref(iters),
                int(self.inner_fixed_point))
            self.inner_iters += iters.value
            self.capped_docs += int(capped)
            return self.gamma, (S, Et_out)
        S = np.zeros((K, V))
        for d in range(D):
            l1 = np.inf
            b, e = X.rowptr[d], X.rowptr[d + 1]
            ids = X.colidx[b:e]
            cts = X.counts[b:e].astype(np.float64)[:, None]  # (w,1) :227
            eT = np.exp(Elogtheta[d, :, None])  # (K,1) :229,232
            eB = np.exp(Elogbeta[:, ids])  # (K,w) :230,233
            it = 0
            while l1 > threshold and it <= step:  # :236
                g_old = np.copy(self.gamma[d])
                norm = np.dot(eT.T, eB) + 1e-100  # :242
                self.gamma[d, :] = (self.alpha + eT.T * np.dot(cts.T / norm, eB.T)).ravel()  # :248
                l1 = np.linalg.norm(self.gamma[d] - g_old, ord=1)
                if self.inner_fixed_point:  # opt-in mode only: refresh exp(E[log theta_d]) inside the loop
                    eT = np.exp(dirichlet_expectation_2d(self.gamma[d:d + 1], "numpy")).T
                it += 1
            if self.inner_fixed_point:
                norm = np.dot(eT.T, eB) + 1e-100
            self.inner_iters += it
            if it > step:
                self.capped_docs += 1
            S[:, ids] += np.outer(eT, cts.T / norm)  # :262
        Et_out = dirichlet_expectation_2d(self.gamma, self.engine)  # :256 (see lda_oracle.c note)
        S *= np.exp(Elogbeta)  # :264
        return self.gamma, (S, Et_out)

    # -- lda_model.py:273-283 ------------------------------------------------------------------------------
    def m_step(self, S):
        if self.engine == "c":
            lam = np.empty_like(S)
            Eb = np.empty_like(S)
            _lib().ttor_m_step(_p(np.ascontiguousarray(S), ctypes.c_double), float(self.eta), S.shape[0],
                               S.shape[1], _p(lam, ctypes.c_double), _p(Eb, ctypes.c_double))
            self.lam = lam
            return lam, Eb
        self.lam = self.eta + S
        return self.lam, dirichlet_expectation_2d(self.lam, self.engine)

    # -- lda_model.py:285-316 ------------------------------------------------------------------------------
    def em_step(self, X, Elogtheta, Elogbeta, step=100):
        self.gamma, (S, Elogtheta) = self.e_step(X, Elogtheta, Elogbeta, step)
        self.lam, Elogbeta = self.m_step(S)
        return Elogtheta, Elogbeta

    # -- lda_model.py:318-432 ------------------------------------------------------------------------------
    def fit(self, X: CSRCorpus, sampling=False, threshold=1e-05, em_num_step=100, e_num_step=100,
            update_hyper_parms=True, lambda_init: np.ndarray | None = None):
        K, V, D = self.K, X.V, X.D
        self.M, self.V = D, V
        if lambda_init is None:
            np.random.seed(SEED)  # :356
            self.lam = np.random.gamma(shape=100, scale=0.01, size=(K, V)).astype(np.float64, copy=False)  # :357
        else:
            self.lam = np.array(lambda_init, dtype=np.float64, copy=True)
        self.gamma = (self.alpha + np.ones((D, K), dtype=np.float64) * V / K).astype(np.float64, copy=False)  # :364
        self.gamma = np.ascontiguousarray(self.gamma)
        perplexities = []
        perp, logs = self.approx_perplexity(X, sampling)  # :373
        perplexities.append(perp)
        delta = np.inf
        Elogtheta, Elogbeta = logs
        for _ in range(em_num_step):  # :384
            if delta < threshold:  # :386
                return perplexities
            orig = perp
            Elogtheta, Elogbeta = self.em_step(X, logs[0], logs[1], step=e_num_step)  # :394
            perp, logs = self.approx_perplexity(X, sampling, Elogtheta, Elogbeta)  # :403
            perplexities.append(perp)
            delta = orig - perp
        if update_hyper_parms:  # :414
            self.update_alpha()
            self.update_eta()
            perp, logs = self.approx_perplexity(X, sampling, Elogtheta, Elogbeta)  # stale Elog, new alpha/eta
            perplexities.append(perp)
        return perplexities

    # -- NOT in the reference (readme.md:131-136 roadmap): restates LDASmoothed.partial_fit of the CUDA build ----------
    def partial_fit(self, X: CSRCorpus, total_docs=None, tau0=1.0, kappa=0.7, e_num_step=100, gamma_init=None,
                    e_threshold=1e-05):
        """Online VB (Hoffman, Blei & Bach 2010, Alg. 2) on one minibatch; returns the minibatch perplexity bound under
        the converged gamma and the topics before the update (document part scaled by total_docs / |B|)."""
        assert self.inner_fixed_point, "online steps solve each document to its fixed point"
        K, V, D = self.K, X.V, X.D
        total = int(total_docs) if total_docs else D
        if not hasattr(self, "t_online"):
            np.random.seed(SEED)
            self.lam = np.random.gamma(shape=100, scale=0.01, size=(K, V)).astype(np.float64, copy=False)
            self.t_online = 0
        self.M, self.V = D, V
        if gamma_init is None:
            self.gamma = np.ascontiguousarray(self.alpha + np.ones((D, K), dtype=np.float64) * V / K)
        else:  # e.g. 1.0: scikit-learn's start without random_init (tests/test_sklearn_pin.py)
            self.gamma = np.full((D, K), float(gamma_init))
        Elogtheta = dirichlet_expectation_2d(self.gamma, self.engine)
        Elogbeta = dirichlet_expectation_2d(self.lam, self.engine)
        self.gamma, (S, Elogtheta) = self.e_step(X, Elogtheta, Elogbeta, e_num_step, e_threshold)
        self.D_population = total
        perp, _ = self.approx_perplexity(X, True, Elogtheta, Elogbeta)
        rho = (tau0 + self.t_online) ** (-kappa)
        self.lam = (1.0 - rho) * self.lam + rho * (self.eta + (total / D) * S)
        self.t_online += 1
        return perp

    # -- lda_model.py:435-509 ------------------------------------------------------------------------------
    def update_alpha(self, step=500, threshold=1e-07):
        it = 0
        while it <= step:
            sum_alpha = np.sum(self.alpha) if isinstance(self.alpha, int) else self.alpha * self.K  # :446
            psg = psi(self.gamma)
            if isinstance(self.alpha, np.ndarray):
                g = self.M * (psi(sum_alpha) - psi(self.alpha)) + \
                    (psg - psi(np.sum(self.gamma, axis=1).reshape(-1, 1))).sum(axis=0)
                h = -self.M * polygamma(1, self.alpha)
                z = self.M * polygamma(1, sum_alpha)
                c = np.sum(g / h) / ((1. / z) + np.sum(1. / h))
                update = (g - c) / h
            else:
                g = self.M * self.K * (psi(self.K * self.alpha) - psi(self.alpha)) + \
                    np.sum(psg - psi(np.sum(self.gamma, 1)).reshape(-1, 1))
                h = self.M * self.K * (self.K * polygamma(1, self.K * self.alpha) - polygamma(1, self.alpha))
                update = g / h
            new = self.alpha - update
            delta = np.sum(np.abs(new - self.alpha))
            self.alpha = new
            it += 1
            if delta < threshold:
                return
        warnings.warn(f"Update alphda Maximum iteration reached at step {it}")

    # -- lda_model.py:511-554 ------------------------------------------------------------------------------
    def update_eta(self, step=1000, threshold=1e-09):
        it = 0
        while it <= step:
            psl = psi(self.lam)
            g = self.K * self.V * (psi(self.V * self.eta) - psi(self.eta)) + \
                np.sum(psl - psi(np.sum(self.lam, 1)).reshape(-1, 1))
            h = self.K * self.V * (self.V * polygamma(1, self.V * self.eta) - polygamma(1, self.eta))
            new = self.eta - g / h
            if new < 0:
                raise ValueError(f"Dirichlet Parameter become < 0 at iteration {it}")
            delta = np.abs(new - self.eta)
            self.eta = new
            it += 1
            if delta < threshold:
                return
        warnings.warn(f"Update Eta: Maximum iteration reached at step {it}")
