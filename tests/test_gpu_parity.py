"""GPU parity tests: the sm_100a kernels, called through the C ABI (tuotuo_b200._native -> libttlda.so), against
(a) the golden traces of the REAL reference (tests/golden/*.npz) and (b) the oracle on seeded inputs.

Tolerances are north_star's: <= 1e-9 relative on gamma / lambda, <= 1e-8 on every perplexity; indices bit-exact.
"""
import json
import os
import warnings

import numpy as np
import pytest
import torch

import lda_oracle as orc
from lda_oracle import CSRCorpus, OracleLDA

pytestmark = pytest.mark.gpu

TOL_STATE = 1e-9
TOL_PERP = 1e-8


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if not a.size:
        return 0.0
    with np.errstate(invalid="ignore"):
        d = np.where(a == b, 0.0, np.abs(a - b) / np.maximum(np.abs(b), 1e-300))  # equal infinities agree
    return float(np.max(d))


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def nv():
    from tuotuo_b200 import _native

    _native.load()
    _native.check_device(0)
    return _native


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def dpad(a, ld, dev):
    """(rows, K) host array -> (rows, ld) device table with zero padding."""
    t = torch.zeros((a.shape[0], ld), dtype=torch.float64, device=dev)
    t[:, :a.shape[1]] = torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    return t


# ------------------------------------------------------------------------------------------------------------------
def test_dirichlet_expectation_known_answers(golden_dir, nv, dev):
    g = _load(golden_dir, "psi_as103.npz")
    for key_in, key_out in (("rows", "de_rows"), ("wide", "de_wide")):
        a = g[key_in]
        ld = nv.ld_for(a.shape[1])
        t = dpad(a, ld, dev)
        elog, ex = torch.empty_like(t), torch.empty_like(t)
        nv.dirichlet_expectation(t, a.shape[1], elog, ex)
        got = elog[:, :a.shape[1]].cpu().numpy()
        # op-for-op AS-103: only FMA-contraction / libm last-bit differences are allowed
        assert np.max(np.abs(got - g[key_out])) <= 1e-13 * max(1.0, np.max(np.abs(g[key_out])))
        np.testing.assert_allclose(ex[:, :a.shape[1]].cpu().numpy(), np.exp(g[key_out]), rtol=1e-12)
        assert float(ex[:, a.shape[1]:].abs().sum()) == 0.0
    # long rows take the CTA-per-row kernel
    rs = np.random.RandomState(3)
    a = rs.gamma(1.0, 1.0, size=(3, 5003))
    t = dpad(a, nv.ld_for(5003), dev)
    elog = torch.empty_like(t)
    nv.dirichlet_expectation(t, 5003, elog, None)
    assert relerr(elog[:, :5003].cpu().numpy(), orc.dirichlet_expectation_2d(a)) < 1e-11
    # 1-d in-place form
    d1 = torch.as_tensor(g["d1_in"].copy()).to(dev)
    out = torch.empty_like(d1)
    nv.dirichlet_expectation_1d(d1, float(g["d1_prior"]), out)
    np.testing.assert_allclose(d1.cpu().numpy(), g["d1_after"], rtol=0, atol=0)
    assert relerr(out.cpu().numpy(), g["d1_out"]) < 1e-13


def test_cutils_shim(golden_dir):
    from tuotuo.cutils import _dirichlet_expectation_1d, _dirichlet_expectation_2d

    g = _load(golden_dir, "psi_as103.npz")
    assert relerr(_dirichlet_expectation_2d(g["wide"]), g["de_wide"]) < 1e-12
    out32 = _dirichlet_expectation_2d(g["wide"].astype(np.float32))
    assert out32.dtype == np.float32 and np.max(np.abs(out32 - g["de_wide32"])) < 2e-5
    with pytest.raises(TypeError):
        _dirichlet_expectation_2d(np.ones((2, 2), dtype=np.int64))
    d1, out = g["d1_in"].copy(), np.empty_like(g["d1_in"])
    _dirichlet_expectation_1d(d1, float(g["d1_prior"]), out)
    np.testing.assert_array_equal(d1, g["d1_after"])
    assert relerr(out, g["d1_out"]) < 1e-13


def test_psi_exact_matches_scipy(nv, dev):
    from scipy.special import psi

    x = np.concatenate([np.logspace(-6, 6, 4001), np.linspace(0.5, 12, 2001)])
    t = torch.as_tensor(x).to(dev)
    out = torch.empty_like(t)
    nv.psi_exact(t, out)
    ref = psi(x)
    assert np.max(np.abs(out.cpu().numpy() - ref) / np.maximum(1.0, np.abs(ref))) < 5e-15


# ------------------------------------------------------------------------------------------------------------------
def test_bow_to_csr_and_csc_bit_exact(golden_dir, nv, dev):
    from tuotuo_b200.corpus import DeviceCorpus

    rs = np.random.RandomState(5)
    V, D = 173, 211
    lens = rs.poisson(30, size=D)
    lens[[0, 7, 8, D - 1]] = 0  # empty documents, including first and last
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = (rs.zipf(1.3, size=int(offs[-1])) % V).astype(np.int32)
    ref = CSRCorpus.from_token_ids(tok, offs, V)
    c = DeviceCorpus.from_token_ids(torch.as_tensor(tok).to(dev), torch.as_tensor(offs).to(dev), V)
    c.build_csc()
    rp, ci, ct = c.host_csr()
    np.testing.assert_array_equal(rp, ref.rowptr)
    np.testing.assert_array_equal(ci, ref.colidx)
    np.testing.assert_array_equal(ct, ref.counts)
    assert c.tokens == int(offs[-1]) and c.nnz == ref.nnz
    # CSC mirror == scipy's transpose, documents ascending inside every word
    import scipy.sparse as sp

    m = sp.csr_matrix((ref.counts, ref.colidx, ref.rowptr), shape=(D, V)).tocsc()
    m.sort_indices()
    np.testing.assert_array_equal(c.colptr.cpu().numpy(), m.indptr)
    np.testing.assert_array_equal(c.docidx.cpu().numpy()[:c.nnz], m.indices)
    np.testing.assert_array_equal(c.ccounts.cpu().numpy()[:c.nnz], m.data)
    # document-blocked mirror: virtual word b*V + w lists the postings of w inside document block b
    nb, bd = 4, 53
    c.build_csc(nb, bd)
    cp, di, cc = c.colptr.cpu().numpy(), c.docidx.cpu().numpy()[:c.nnz], c.ccounts.cpu().numpy()[:c.nnz]
    assert len(cp) == nb * V + 1 and cp[0] == 0 and cp[-1] == c.nnz
    full = sp.csr_matrix((ref.counts, ref.colidx, ref.rowptr), shape=(D, V))
    for b in range(nb):
        blk = full[b * bd:min(D, (b + 1) * bd)].tocsc()
        blk.sort_indices()
        np.testing.assert_array_equal(np.diff(cp[b * V:(b + 1) * V + 1]), np.diff(blk.indptr))
        np.testing.assert_array_equal(di[cp[b * V]:cp[(b + 1) * V]], blk.indices + b * bd)
        np.testing.assert_array_equal(cc[cp[b * V]:cp[(b + 1) * V]], blk.data)
    # golden: the reference's Document on the README corpus
    j = json.load(open(os.path.join(golden_dir, "cfg1_docs.json")))
    from tuotuo.document import Document

    d = Document({int(k): v for k, v in j["docs"].items()})
    g = _load(golden_dir, "fit_cfg1.npz")
    np.testing.assert_array_equal(d.doc_vocab_count_array, g["X"])
    assert not d.doc_vocab_count_array.flags["C_CONTIGUOUS"] and d.doc_vocab_count_array.dtype == np.float64
    # empty corpus
    e = DeviceCorpus.from_token_ids(torch.zeros(0, dtype=torch.int32, device=dev),
                                    torch.zeros(4, dtype=torch.int64, device=dev), 9)
    e.build_csc()
    assert e.nnz == 0 and e.rowptr.cpu().tolist() == [0, 0, 0, 0] and e.colptr.cpu().tolist() == [0] * 10


# ------------------------------------------------------------------------------------------------------------------
def _single_step(nv, dev, X: CSRCorpus, K, gamma0, lam0, alpha, mode, nblocks=1, handoff=False):
    """theta/beta refresh -> estep -> sstats -> mstep -> tokens, all through the C ABI; returns host results."""
    from tuotuo_b200.corpus import DeviceCorpus

    c = DeviceCorpus.from_host_csr(X.rowptr, X.colidx, X.counts, X.V, dev)
    c.build_csc(nblocks)
    ld = nv.ld_for(K)
    D, V = X.D, X.V
    gamma = dpad(gamma0, ld, dev)
    lam = torch.zeros((V, ld), dtype=torch.float64, device=dev)
    nv.transpose_pad(torch.as_tensor(np.ascontiguousarray(lam0)).to(dev), lam)
    eT0 = torch.zeros((D, ld), dtype=torch.float64, device=dev)
    eT1 = torch.zeros_like(eT0)
    eB = torch.zeros_like(lam)
    S = torch.zeros_like(lam)
    lam_new = torch.zeros_like(lam)
    colsum = torch.zeros(ld, dtype=torch.float64, device=dev)
    part = torch.zeros((5, nv.NPART), dtype=torch.float64, device=dev)
    iters = torch.zeros(2, dtype=torch.int64, device=dev)
    ws = torch.empty(nv.reduce_workspace_bytes(), dtype=torch.uint8, device=dev)
    ws_m = torch.empty(nv.mstep_workspace_bytes(V, ld), dtype=torch.uint8, device=dev)
    ws_s = torch.empty(nv.sstats_workspace_bytes(c.nnz, V, ld, c.nblocks), dtype=torch.uint8, device=dev)
    av = torch.as_tensor(np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=np.float64).reshape(-1), (K,)))).to(dev)
    asum = float(np.sum(alpha))
    nv.theta_refresh(gamma, K, av, asum, eT0, part[0], ws)
    nv.beta_refresh(lam, K, 1.0, eB, None, colsum, part[1], ws_m)
    nv.elbo_tokens(c.rowptr, c.colidx, c.counts, K, eT0, eB, part[2], ws)
    elbo0 = float(part[0, 0] + part[1, 0] + part[2, 0])
    ratio = torch.full((max(c.nnz, 1),), float("nan"), dtype=torch.float64, device=dev) if handoff else None
    nv.estep(c.rowptr, c.colidx, c.counts, K, V, av, asum, eT0, eB, gamma, gamma, eT1, 100, 1e-5, iters, part[3], ws,
             c.csr2csc if handoff else None, ratio)
    if mode == "atomic":
        nv.sstats_atomic(c.rowptr, c.colidx, c.counts, K, V, eT0, eB, S)
    else:
        nv.sstats_csc(c.colptr, c.docidx, c.ccounts, V, K, c.nnz, c.nblocks, eT0, eB, S, ws_s, ratio)
    upd = torch.empty((K, V), dtype=torch.float64, device=dev)
    nv.transpose_unpad(S * eB, K, upd)
    elog_b = torch.zeros_like(lam)
    nv.mstep(S, K, 1.0, eB, lam_new, elog_b, colsum, part[1], ws_m)
    nv.elbo_tokens(c.rowptr, c.colidx, c.counts, K, eT1, eB, part[2], ws)
    elbo1 = float(part[3, 1] + part[1, 0] + part[2, 0])
    lam_out = torch.empty((K, V), dtype=torch.float64, device=dev)
    nv.transpose_unpad(lam_new, K, lam_out)
    eb_out = torch.empty((K, V), dtype=torch.float64, device=dev)
    nv.transpose_unpad(elog_b, K, eb_out)
    return dict(elbo0=elbo0, elbo1=elbo1, gamma=gamma[:, :K].cpu().numpy(), sstats=upd.cpu().numpy(),
                lam=lam_out.cpu().numpy(), elogbeta=eb_out.cpu().numpy(), eT1=eT1[:, :K].cpu().numpy(),
                iters=iters.cpu().numpy(), tok_in=float(part[3, 0]), count=float(part[3, 2]), pad=float(
                    gamma[:, K:].abs().sum() + eT1[:, K:].abs().sum() + lam_new[:, K:].abs().sum() + eB[:, K:].abs().sum()))


@pytest.mark.parametrize("mode,nblocks,handoff", [("csc", 1, False), ("csc", 3, False), ("csc", 200, False),
                                                  ("csc", 1, True), ("csc", 5, True), ("atomic", 1, False)])
def test_single_em_step_vs_reference(golden_dir, nv, dev, mode, nblocks, handoff):
    g = _load(golden_dir, "estep_small.npz")
    X = CSRCorpus(g["rowptr"], g["colidx"], g["counts"], int(g["V"]))
    r = _single_step(nv, dev, X, int(g["K"]), g["gamma0"], g["lam0"], 1, mode, nblocks, handoff)
    assert relerr(r["elbo0"], g["elbo0"]) < 1e-12
    assert relerr(r["gamma"], g["gamma1"]) < TOL_STATE
    assert relerr(r["sstats"], g["sstats1"]) < TOL_STATE
    assert relerr(r["lam"], g["lam1"]) < TOL_STATE
    assert np.max(np.abs(r["elogbeta"] - g["Elogbeta1"])) < 1e-11
    assert relerr(r["eT1"], np.exp(g["Elogtheta1"])) < 1e-11
    assert relerr(r["elbo1"], g["elbo1"]) < 1e-12
    assert r["iters"][0] == 2 * X.D and r["iters"][1] == 0  # the reference's loop: exactly two passes per document
    assert r["count"] == X.tokens and r["pad"] == 0.0


FITS = ["fit_small.npz", "fit_nonexch.npz", "fit_sampling.npz", "fit_early.npz", "fit_ragged.npz", "fit_cfg2_slice.npz"]


@pytest.mark.parametrize("name", FITS)
def test_fit_traces_vs_reference(golden_dir, name):
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    g = _load(golden_dir, name)
    kw = json.loads(str(g["kw"]))
    exch = g["alpha0"].ndim == 0
    doc = Document.from_csr(g["rowptr"], g["colidx"], g["counts"], int(g["V"]))
    m = LDASmoothed(int(g["K"]), exchangeable_prior=exch, verbose=False)
    if "D_population" in g:
        m.D_population = int(g["D_population"])
    perps = m.fit(doc, return_perplexities=True, **kw)
    assert len(perps) == len(g["perplexities"]) and all(isinstance(p, np.float64) for p in perps)
    assert relerr(perps, g["perplexities"]) < TOL_PERP
    assert m._gamma_.shape == g["gamma"].shape and m._lambda_.shape == g["lam"].shape
    assert relerr(m._gamma_, g["gamma"]) < TOL_STATE
    assert relerr(m._lambda_, g["lam"]) < TOL_STATE
    assert relerr(m._alpha_, g["alpha"]) < 1e-10 and relerr(m._eta_, g["eta"]) < 1e-10


def test_fit_readme_config(golden_dir, capsys):
    """BASELINE.json configs[0]: doc_generator M=100 L=20 -> Document -> LDASmoothed(5).fit, 100 iterations + hyper."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    j = json.load(open(os.path.join(golden_dir, "cfg1_docs.json")))
    g = _load(golden_dir, "fit_cfg1.npz")
    lda = LDASmoothed(num_topics=5)
    perps = lda.fit({int(k): v for k, v in j["docs"].items()}, sampling=False, verbose=True, return_perplexities=True)
    out = capsys.readouterr().out
    assert "Var Inf - Word Dirichlet prior, Lambda\n(5, 40)" in out and "Var Inf - Topic Dirichlet prior, Gamma\n(100, 5)" in out
    assert f"Init perplexity = {perps[0]}" in out and f"End perplexity = {perps[-1]}" in out
    assert len(perps) == 102
    assert relerr(perps, g["perplexities"]) < TOL_PERP
    assert relerr(lda._gamma_, g["gamma"]) < TOL_STATE and relerr(lda._lambda_, g["lam"]) < TOL_STATE
    assert relerr(lda._alpha_, g["alpha"]) < 1e-9 and relerr(lda._eta_, g["eta"]) < 1e-9
    assert lda.M == 100 and lda.V == 40 and lda.train_doc.idx_to_vocab[0] == list(j["vocab_to_idx"])[0]
    lda.get_top_k_words(3)
    # lda_model.py:556-564: per topic, the 3 largest lambda entries in ASCENDING order of lambda, printed as "Top i -> word"
    want = ""
    for k in range(5):
        want += f"Topic {k}\n"
        for i, idx in enumerate(np.argsort(g["lam"][k, :])[-3:]):
            want += f"Top {i + 1} -> {list(j['vocab_to_idx'])[idx]}\n"
        want += "\n"
    assert capsys.readouterr().out == want
    assert lda.fit(lda.train_doc, em_num_step=1) is None  # return_perplexities defaults to False


@pytest.mark.parametrize("K,V,D,L", [(100, 2000, 300, 120), (500, 1500, 64, 200), (1000, 1200, 40, 150),
                                      (37, 700, 500, 60), (64, 500, 200, 50), (130, 900, 100, 80), (257, 600, 50, 90)])
def test_fit_vs_oracle_many_K(K, V, D, L):
    """Seeded corpora at topic counts covering every register-tiling variant (KP = 1..16), vs the C oracle."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(100 + K, D, V, min(K, 40), L, ragged=True)
    X = CSRCorpus(rp, ci, ct, V)
    o = OracleLDA(K, engine="c")
    po = o.fit(X, em_num_step=3)
    m = LDASmoothed(K, verbose=False)
    pg = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, return_perplexities=True)
    assert relerr(pg, po) < TOL_PERP  # (tiny corpora with K*V >> tokens overflow exp() to inf on both sides)
    assert relerr(m._last_elbo, o.last_elbo) < 1e-11
    assert relerr(m._gamma_, o.gamma) < TOL_STATE and relerr(m._lambda_, o.lam) < TOL_STATE
    assert relerr(m._alpha_, o.alpha) < 1e-9 and relerr(m._eta_, o.eta) < 1e-9


def test_atomic_mode_matches_csc_mode():
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(9, 400, 800, 20, 70, ragged=True)
    doc = Document.from_csr(rp, ci, ct, 800)
    a = LDASmoothed(24, verbose=False, sstats_mode="atomic")
    pa = a.fit(doc, em_num_step=4, return_perplexities=True)
    b = LDASmoothed(24, verbose=False, sstats_mode="csc")
    pb = b.fit(doc, em_num_step=4, return_perplexities=True)
    assert relerr(pa, pb) < 1e-12 and relerr(a._lambda_, b._lambda_) < 1e-11 and relerr(a._gamma_, b._gamma_) < 1e-11
    # the two-pass path is bit-reproducible run to run
    c = LDASmoothed(24, verbose=False, sstats_mode="csc")
    pc = c.fit(doc, em_num_step=4, return_perplexities=True)
    assert pb == pc and np.array_equal(b._lambda_, c._lambda_) and np.array_equal(b._gamma_, c._gamma_)


def test_reference_shaped_step_api(golden_dir):
    """e_step / m_step / em_step / approx_elbo / approx_perplexity keep the reference's argument and return shapes."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    g = _load(golden_dir, "estep_small.npz")
    X = CSRCorpus(g["rowptr"], g["colidx"], g["counts"], int(g["V"]))
    doc = Document.from_csr(X.rowptr, X.colidx, X.counts, X.V)
    m = LDASmoothed(int(g["K"]), verbose=False)
    m.fit(doc, em_num_step=0, update_hyper_parms=False)
    Xd = doc.doc_vocab_count_array
    elbo0, (Et0, Eb0) = m.approx_elbo(Xd)
    assert relerr(elbo0, g["elbo0"]) < 1e-12 and Et0.shape == (X.D, m.K) and Eb0.shape == (m.K, X.V)
    assert np.max(np.abs(Et0 - g["Elogtheta0"])) < 1e-11 and np.max(np.abs(Eb0 - g["Elogbeta0"])) < 1e-11
    gam, (S, Et1) = m.e_step(Xd, Et0, Eb0)
    assert relerr(gam, g["gamma1"]) < TOL_STATE and relerr(S, g["sstats1"]) < TOL_STATE
    assert np.max(np.abs(Et1 - g["Elogtheta1"])) < 1e-11
    lam, Eb1 = m.m_step(S)
    assert relerr(lam, g["lam1"]) < TOL_STATE and np.max(np.abs(Eb1 - g["Elogbeta1"])) < 1e-11
    perp, _ = m.approx_perplexity(Xd, False, Et1, Eb1)
    assert relerr(perp, np.exp(-g["elbo1"] / X.tokens)) < TOL_PERP


def test_estep_iteration_cap_statistics(golden_dir, nv, dev):
    """max_iter = 0 reproduces the reference's `it > step` warning condition for every document."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    g = _load(golden_dir, "estep_small.npz")
    doc = Document.from_csr(g["rowptr"], g["colidx"], g["counts"], int(g["V"]))
    m = LDASmoothed(int(g["K"]), verbose=False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m.fit(doc, em_num_step=1, e_num_step=0, update_hyper_parms=False)
    assert any("Maximum iteration reached" in str(x.message) for x in w)
    assert relerr(m._gamma_, g["gamma1"]) < TOL_STATE


def test_invariants_at_scale(nv, dev):
    """Size-independent properties on a corpus far beyond what the oracle can check (200k docs x 300 tokens, K=100):
    sum_k (gamma_dk - alpha) = document length; sum_wk sstats*eB = token count; perplexity decreases."""
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import synthetic_document

    doc = synthetic_document("cfg3", device=dev, docs=200_000)
    m = LDASmoothed(100, verbose=False)
    perps = m.fit(doc, em_num_step=3, update_hyper_parms=False, return_perplexities=True)
    st = m._st
    c = st.corpus
    assert c.tokens == 200_000 * 300 and c.nblocks > 1  # 160 MB of theta rows => document-blocked statistics pass
    lens = torch.zeros(c.D, dtype=torch.float64, device=dev)
    rows = torch.repeat_interleave(torch.arange(c.D, device=dev), torch.diff(c.rowptr))
    lens.index_add_(0, rows, c.counts.to(torch.float64))
    assert float(((st.gamma[st.cur][:, :100].sum(1) - 100.0) - lens).abs().max()) < 1e-9
    tot = float((st.lam[:, :100] - 1.0).sum())  # lambda - eta = sstats * eB
    assert abs(tot - c.tokens) / c.tokens < 1e-12
    assert perps[0] > perps[1] > perps[2] > perps[3]
    assert st.terms["count"] == c.tokens


def test_cfg3_shaped_slice_vs_oracle(monkeypatch, dev):
    """A slice with the headline configuration's own shape (K=100, V=50k, 300-token documents from the bench's
    generator) through the production path — blocked mirror, ticket-ordered statistics, ratio hand-off, DMMA group
    reduction — against the C oracle, hyper update included."""
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import synthetic_document

    monkeypatch.setenv("TTLDA_DOC_BLOCKS", "4")
    doc = synthetic_document("cfg3", device=dev, docs=3000)
    m = LDASmoothed(100, verbose=False)
    p = m.fit(doc, em_num_step=3, return_perplexities=True)
    assert m._st.corpus.nblocks == 4 and m._st.ratio is not None
    rp, ci, ct = doc.host_csr()
    o = OracleLDA(100)
    po = o.fit(CSRCorpus(rp, ci, ct, 50000), em_num_step=3)
    assert len(p) == len(po) and relerr(p, po) < TOL_PERP
    assert relerr(m._gamma_, o.gamma) < TOL_STATE and relerr(m._lambda_, o.lam) < TOL_STATE
    assert relerr(m._alpha_, o.alpha) < TOL_STATE and relerr(m._eta_, o.eta) < TOL_STATE


@pytest.mark.parametrize("seed", range(6))
def test_corpus_kernels_random_ragged(nv, dev, seed):
    """bow -> CSR -> (blocked) CSC mirror -> per-block mirror -> widen on random ragged inputs (empty documents, unused
    vocabulary, long rows), bit-exact against NumPy."""
    from tuotuo_b200.corpus import DeviceCorpus

    rs = np.random.RandomState(1000 + seed)
    D, V = int(rs.randint(1, 400)), int(rs.randint(1, 70000 if seed % 2 else 300))
    lens = rs.poisson(rs.choice([0.5, 5, 60]), size=D)
    lens[rs.rand(D) < 0.2] = 0
    tok = rs.randint(0, V, size=int(lens.sum())).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    c = DeviceCorpus.from_token_ids(torch.as_tensor(tok).to(dev), torch.as_tensor(offs).to(dev), V)
    rp, ci, ct = c.host_csr()
    # CSR vs numpy
    exp_rp, exp_ci, exp_ct = [0], [], []
    for d in range(D):
        u, n = np.unique(tok[offs[d]:offs[d + 1]], return_counts=True)
        exp_ci.append(u)
        exp_ct.append(n)
        exp_rp.append(exp_rp[-1] + len(u))
    exp_ci = np.concatenate(exp_ci) if exp_ci else np.zeros(0)
    exp_ct = np.concatenate(exp_ct) if exp_ct else np.zeros(0)
    assert np.array_equal(rp, exp_rp) and np.array_equal(ci, exp_ci) and np.array_equal(ct, exp_ct)
    nnz = len(ci)
    # blocked mirror vs numpy (stable sort by (block, word))
    nb = int(rs.randint(1, 5))
    bd = max(1, -(-D // nb))
    nb = max(1, -(-D // bd))
    c.build_csc(nb, bd)
    docs = np.repeat(np.arange(D), np.diff(rp))
    key = (docs // bd) * V + ci
    order = np.argsort(key, kind="stable")
    colptr = np.searchsorted(key[order], np.arange(nb * V + 1), side="left")
    assert np.array_equal(c.colptr.cpu().numpy(), colptr)
    assert np.array_equal(c.docidx.cpu().numpy()[:nnz], docs[order])
    assert np.array_equal(c.ccounts.cpu().numpy()[:nnz], ct[order])
    inv = np.empty(nnz, dtype=np.int64)
    inv[order] = np.arange(nnz)
    assert np.array_equal(c.csr2csc.cpu().numpy()[:nnz].astype(np.int64), inv)
    # the same mirror one block at a time
    i32 = dict(dtype=torch.int32, device=dev)
    blk = [torch.full((nb * V + 1,), -1, dtype=torch.int64, device=dev), torch.full((max(nnz, 1),), -1, **i32),
           torch.full((max(nnz, 1),), -1, **i32), torch.full((max(nnz, 1),), -1, **i32)]
    ws = torch.empty(nv.csr_to_csc_workspace_bytes(nnz, D, V, 1), dtype=torch.uint8, device=dev)
    for b in range(nb):
        d0, d1 = min(b * bd, D), min((b + 1) * bd, D)
        s, e = int(rp[d0]), int(rp[d1])
        nv.csr_to_csc_block(c.rowptr[d0:d1 + 1], c.colidx, c.counts, V, d0, s, e - s, blk[0][b * V:b * V + V + 1],
                            blk[1], blk[2], ws, blk[3])
    assert torch.equal(blk[0], c.colptr)
    for a, b_ in zip(blk[1:], (c.docidx, c.ccounts, c.csr2csc)):
        assert torch.equal(a[:nnz], b_[:nnz])
    # widen
    if V <= 65536 and nnz:
        out = torch.empty(nnz, **i32)
        nv.widen_i32(torch.as_tensor(ci.astype(np.uint16).view(np.int16)).to(dev), out)
        assert np.array_equal(out.cpu().numpy(), ci)
        out8 = torch.empty(nnz, **i32)
        nv.widen_i32(torch.as_tensor(np.minimum(ct, 255).astype(np.uint8)).to(dev), out8)
        assert np.array_equal(out8.cpu().numpy(), np.minimum(ct, 255))


def test_blocked_and_unblocked_statistics_agree(monkeypatch):
    """The document-blocked CSC mirror only reorders the (fixed-order) summation: results agree to rounding."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(21, 700, 900, 20, 60, ragged=True)
    res = {}
    for nb in ("1", "7"):
        monkeypatch.setenv("TTLDA_DOC_BLOCKS", nb)
        doc = Document.from_csr(rp, ci, ct, 900)
        m = LDASmoothed(50, verbose=False)
        p = m.fit(doc, em_num_step=3, return_perplexities=True)
        assert m._st.corpus.nblocks == int(nb)
        res[nb] = (p, m._lambda_.copy(), m._gamma_.copy())
    assert relerr(res["1"][0], res["7"][0]) < 1e-13
    assert relerr(res["1"][1], res["7"][1]) < 1e-12 and relerr(res["1"][2], res["7"][2]) < 1e-12


@pytest.mark.parametrize("K,kw,blocks", [(100, "32", "1"), (100, "32", "3"), (300, "128", "2"), (1000, "128", "4"),
                                         (1000, "256", "1"), (36, "16", "2")])
def test_topic_windowed_statistics_vs_oracle(monkeypatch, K, kw, blocks):
    """ttlda_sstats_csc_windowed_f64 (the form cfg5-class shapes run: one walk of the blocked mirror per window of kw
    topics): same fit as the unwindowed pass and as the oracle; run-to-run bitwise reproducible."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    V, D, L = 1300, 150, 80
    rp, ci, ct = numpy_corpus(4000 + K, D, V, min(K, 20), L, ragged=True)
    monkeypatch.setenv("TTLDA_DOC_BLOCKS", blocks)
    monkeypatch.setenv("TTLDA_KWINDOW", "0")
    plain = LDASmoothed(K, verbose=False)
    pp = plain.fit(Document.from_csr(rp, ci, ct, V), em_num_step=2, return_perplexities=True)
    assert plain._st.corpus.kwindow == 0
    monkeypatch.setenv("TTLDA_KWINDOW", kw)
    runs = []
    for _ in range(2):
        m = LDASmoothed(K, verbose=False)
        p = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=2, return_perplexities=True)
        assert m._st.corpus.kwindow == int(kw) and m._st.corpus.nblocks == int(blocks)
        runs.append((p, m._gamma_.copy(), m._lambda_.copy()))
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    assert relerr(runs[0][2], plain._lambda_) < 1e-13 and relerr(runs[0][1], plain._gamma_) < 1e-13
    o = OracleLDA(K, engine="c")
    po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=2)
    assert relerr(runs[0][0], po) < TOL_PERP and relerr(runs[0][0], pp) < 1e-12
    assert relerr(runs[0][1], o.gamma) < TOL_STATE and relerr(runs[0][2], o.lam) < TOL_STATE


@pytest.mark.parametrize("K,V,D,L,handoff", [(20, 120, 200, 40, "1"), (100, 3000, 250, 120, "1"), (100, 3000, 250, 120, "0"),
                                              (60, 900, 150, 80, "1"), (250, 1500, 80, 150, "1"), (500, 1200, 40, 200, "1")])
def test_hot_word_cache_estep_vs_oracle(monkeypatch, K, V, D, L, handoff):
    """Opt-in E-step with the most frequent words' exp(E[log beta]) rows in shared memory (ttlda_estep_hot_f64 on the
    stream built by ttlda_build_estream): the stream is a per-document stable partition of the CSR row (checked bit for
    bit), the fit equals the oracle's, and the streamed step runs the same kernel (bitwise equal to the resident one)."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    monkeypatch.setenv("TTLDA_HOT", "1")
    monkeypatch.setenv("TTLDA_RATIO_HANDOFF", handoff)
    monkeypatch.setenv("TTLDA_DOC_BLOCKS", "2")
    rp, ci, ct = numpy_corpus(7000 + K, D, V, min(K, 20), L, ragged=True)
    m = LDASmoothed(K, verbose=False)
    p = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, return_perplexities=True)
    st, c = m._st, m._st.corpus
    assert st.hot and c.hot_words is not None
    H = c.hot_words.numel()
    assert H == min(V, m._st.corpus.hot_words.numel()) and 0 < c.hot_share <= 1.0
    # the stream: hot entries (slot | 0x80000000) first, in CSR order, then the cold ones, in CSR order
    e_idx, e_cnt = c.e_idx.cpu().numpy(), c.e_cnt.cpu().numpy()
    slot = c.slot_of_word.cpu().numpy()
    hot_words = c.hot_words.cpu().numpy()
    for d in (0, 1, D // 2, D - 1):
        b, e = rp[d], rp[d + 1]
        w, cnt = ci[b:e], ct[b:e]
        is_hot = slot[w] >= 0
        want_idx = np.concatenate([(slot[w[is_hot]].astype(np.int64) | 0x80000000).astype(np.uint32).view(np.int32),
                                   w[~is_hot]])
        assert np.array_equal(e_idx[b:e], want_idx) and np.array_equal(e_cnt[b:e], np.concatenate([cnt[is_hot], cnt[~is_hot]]))
        assert np.array_equal(hot_words[slot[w[is_hot]]], w[is_hot])
    o = OracleLDA(K, engine="c")
    po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=3)
    assert relerr(p, po) < TOL_PERP and relerr(m._gamma_, o.gamma) < TOL_STATE and relerr(m._lambda_, o.lam) < TOL_STATE
    assert m.inner_iterations == 2 * D
    # streamed step == resident iteration, bit for bit, with the cache on both sides
    a = LDASmoothed(K, verbose=False)
    a.fit(Document.from_csr(rp, ci, ct, V), em_num_step=2, update_hyper_parms=False)
    b_ = LDASmoothed(K, verbose=False)
    b_.fit(Document.from_csr(rp, ci, ct, V), em_num_step=0, update_hyper_parms=False)
    h = [torch.as_tensor(x).pin_memory() for x in (rp, ci, ct)]
    for _ in range(2):
        b_.em_step_from_host(*h)
    assert np.array_equal(a._gamma_, b_._gamma_) and np.array_equal(a._lambda_, b_._lambda_)


@pytest.mark.parametrize("K,V,N", [(20, 97, 3), (100, 1000, 8), (500, 333, 2), (7, 5, 8)])
def test_sharded_mstep_slices_equal_full_mstep(nv, dev, K, V, N):
    """ttlda_mstep_slice_lambda_f64 / ttlda_mstep_slice_beta_f64 — the per-rank halves of the sharded M-step, here with
    the N "ranks" emulated on one GPU: N source statistics tables summed in order inside phase A (what the peer-mapped
    NVLink loads do), every slice's exp(E[log beta]) rows stored into N destination tables in phase B (the NVLink
    stores) — against ttlda_mstep_f64 on the pre-summed table."""
    g = torch.Generator(device="cpu").manual_seed(11 * K + V)
    ld = nv.ld_for(K)
    Vs = -(-V // N)
    Vp = Vs * N
    parts = [torch.zeros((Vp, ld), dtype=torch.float64) for _ in range(N)]
    for t in parts:
        t[:V, :K] = torch.rand((V, K), generator=g, dtype=torch.float64) * 3
    eB0 = torch.zeros((Vp, ld), dtype=torch.float64)
    eB0[:V, :K] = torch.rand((V, K), generator=g, dtype=torch.float64) + 0.01
    parts = [t.to(dev) for t in parts]
    eta = 0.7
    # reference: full M-step on the sum
    S = torch.stack(parts).sum(0)[:V].contiguous()
    eB_ref, lam_ref = eB0[:V].clone().to(dev), torch.zeros((V, ld), dtype=torch.float64, device=dev)
    colsum_ref, part_ref = torch.zeros(ld, dtype=torch.float64, device=dev), torch.zeros(nv.NPART, dtype=torch.float64, device=dev)
    ws = torch.empty(nv.mstep_workspace_bytes(V, ld), dtype=torch.uint8, device=dev)
    nv.mstep(S, K, eta, eB_ref, lam_ref, None, colsum_ref, part_ref, ws)
    # sharded: every "rank" has its own copy of the old eB, its lambda table, and receives every slice's new eB rows
    eBs = [eB0.clone().to(dev) for _ in range(N)]
    lams = [torch.zeros((Vp, ld), dtype=torch.float64, device=dev) for _ in range(N)]
    colparts = torch.zeros((N, ld), dtype=torch.float64, device=dev)
    for r in range(N):
        w0 = r * Vs
        n = max(0, min(Vs, V - w0))
        nv.mstep_slice_lambda([t.data_ptr() for t in parts], w0, w0, n, K, eta, eBs[r], lams[r], colparts[r], ws)
    beta_total = 0.0
    colsum = torch.zeros(ld, dtype=torch.float64, device=dev)
    for r in range(N):
        w0 = r * Vs
        n = max(0, min(Vs, V - w0))
        pr = torch.zeros(nv.NPART, dtype=torch.float64, device=dev)
        nv.mstep_slice_beta(colparts, w0, n, K, eta, lams[r], [t.data_ptr() for t in eBs], r == 0, colsum, pr, ws)
        beta_total += float(pr[0].item())
    lam = torch.zeros((V, ld), dtype=torch.float64, device=dev)
    for r in range(N):
        w0 = r * Vs
        n = max(0, min(Vs, V - w0))
        lam[w0:w0 + n] = lams[r][w0:w0 + n]
    assert relerr(lam.cpu().numpy(), lam_ref.cpu().numpy()) < 1e-14
    assert relerr(colsum.cpu().numpy()[:K], colsum_ref.cpu().numpy()[:K]) < 1e-14
    for t in eBs:  # every destination table received every slice
        assert relerr(t[:V].cpu().numpy(), eB_ref.cpu().numpy()) < 1e-12
        assert torch.equal(t, eBs[0]) and float(t[V:].abs().sum()) == 0.0
    assert abs(beta_total - float(part_ref[0].item())) <= 1e-12 * abs(float(part_ref[0].item()))


def _random_rows(seed, D, V, lo, hi, empty_every=0):
    """CSR with `lo..hi` distinct, ascending word ids per document (uniform over the vocabulary), optional empty rows."""
    rs = np.random.RandomState(seed)
    rows = []
    for d in range(D):
        n = 0 if (empty_every and d % empty_every == 0) else int(rs.randint(lo, hi + 1))
        rows.append(np.sort(rs.choice(V, size=min(n, V), replace=False)).astype(np.int32))
    rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    ci = np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)
    ct = rs.randint(1, 9, size=len(ci)).astype(np.int32)
    return rp, ci, ct


@pytest.mark.parametrize("case", ["ragged", "three_blocks", "tiny", "one_cta_per_sm", "long_documents", "many_slices",
                                  "long_slices", "very_long_slices", "longer_than_a_stage", "vocabulary_too_large"])
def test_slice_walk_mirror_equals_sorted_mirror(monkeypatch, dev, case):
    """The default mirror — the hand-written counting sort of csrc/corpus.cu (per-slice in-order walk with 16-bit
    shared-memory counters fed by a cp.async stage, scan over slices, placement) — builds the same colptr / docidx /
    ccounts / csr2csc, bit for bit, as the generic CUB pair sort (TTLDA_CSC_WALK=0) and as a stable NumPy argsort: ragged
    and empty documents, several document blocks, a vocabulary that leaves one CTA per SM, documents longer than one walk
    step (> 256 entries) and longer than a whole stage buffer, slices of more than one 256-document offsets chunk, many
    small slices (generic placement), and a vocabulary beyond one counter table (V = 120 000: two word ranges, one walk
    CTA per (slice, range))."""
    from tuotuo_b200.corpus import DeviceCorpus
    from tuotuo_b200.synth import numpy_corpus

    blocks = 1
    if case == "ragged":
        D, V = 700, 900
        rp, ci, ct = numpy_corpus(1000, D, V, 8, 30, ragged=True)
    elif case == "three_blocks":
        D, V, blocks = 2500, 70_000, 3
        rp, ci, ct = numpy_corpus(2800, D, V, 8, 30, ragged=True)
    elif case == "tiny":
        D, V, blocks = 40, 12, 2
        rp, ci, ct = numpy_corpus(340, D, V, 8, 30, ragged=True)
    elif case == "one_cta_per_sm":
        D, V = 65, 102_660
        rp, ci, ct = _random_rows(5, D, V, 0, 400, empty_every=7)
    elif case == "long_documents":
        D, V = 90, 5_000
        rp, ci, ct = _random_rows(6, D, V, 300, 1700, empty_every=11)
    elif case == "many_slices":  # 296 slices of 68 documents; V = 50k as at cfg3
        D, V = 20_000, 50_000
        rp, ci, ct = _random_rows(7, D, V, 0, 12, empty_every=5)
    elif case == "longer_than_a_stage":  # V = 100k leaves ~3 500-entry stage buffers: longer documents go straight from global
        D, V = 40, 100_000
        rp, ci, ct = _random_rows(10, D, V, 0, 6000, empty_every=6)
    elif case in ("long_slices", "very_long_slices"):
        # 541 documents per slice: more than one 256-document offsets chunk in the walk; 2 365 per slice: more than one
        # 2 048-document chunk in the placement
        D, V = (160_000 if case == "long_slices" else 700_000), 200
        rs = np.random.RandomState(9)
        n = rs.randint(0, 4, size=D)
        start, stride = rs.randint(0, V, size=D), rs.randint(1, 50, size=D)
        k = np.arange(3)[None, :]
        w = np.sort(np.where(k < n[:, None], (start[:, None] + k * stride[:, None]) % V, V + k), axis=1)  # distinct per row
        rp = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        ci = w[w < V].astype(np.int32)
        ct = rs.randint(1, 9, size=len(ci)).astype(np.int32)
    else:
        D, V = 300, 120_000
        rp, ci, ct = _random_rows(8, D, V, 0, 40)
    out = []
    for walk in ("0", "1"):
        monkeypatch.setenv("TTLDA_CSC_WALK", walk)
        c = DeviceCorpus.from_host_csr(rp, ci, ct, V, dev)
        c.build_csc(blocks)
        out.append([t.clone() for t in (c.colptr, c.docidx[:c.nnz], c.ccounts[:c.nnz], c.csr2csc[:c.nnz])])
    for a, b in zip(*out):
        assert torch.equal(a, b)
    docs = np.repeat(np.arange(D), np.diff(rp))
    bd = -(-D // blocks)
    order = np.argsort((docs // bd) * V + ci, kind="stable")
    assert np.array_equal(out[1][1].cpu().numpy(), docs[order])
    assert np.array_equal(out[1][2].cpu().numpy(), ct[order])
    assert np.array_equal(np.argsort(out[1][3].cpu().numpy().astype(np.int64), kind="stable"), order)


def test_slice_walk_mirror_redoes_unordered_rows_with_atomics(monkeypatch, dev):
    """The fast walk updates its 16-bit counters without atomics, which is only right for rows that hold every word once
    (what np.nonzero gives the reference, what Document.from_csr checks).  It verifies that on the fly (strictly ascending
    rows); a block that fails is redone by the atomic form of the same walk, so ANY row order / repetition still gets a
    complete, in-bounds mirror: (a) rows that repeat words — the repeats take distinct ranks, every mirror slot holds the
    right document; (b) unsorted rows of distinct words — bit-identical to the generic pair sort."""
    from tuotuo_b200.corpus import DeviceCorpus

    V, D = 30, 50
    rs = np.random.RandomState(3)
    rows = [np.sort(rs.randint(0, V, size=rs.randint(0, 25))).astype(np.int32) for _ in range(D)]  # with repeats
    rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    ci = np.concatenate(rows).astype(np.int32)
    ct = np.ones(len(ci), np.int32)
    c = DeviceCorpus.from_host_csr(rp, ci, ct, V, dev)
    c.build_csc(1)
    docs = np.repeat(np.arange(D), np.diff(rp))
    order = np.argsort(ci, kind="stable")
    assert np.array_equal(c.docidx[:c.nnz].cpu().numpy(), docs[order])
    pos = c.csr2csc[:c.nnz].cpu().numpy().astype(np.int64)
    assert np.array_equal(np.sort(pos), np.arange(len(ci)))  # a permutation: every slot is written by one entry
    assert np.array_equal(ci[np.argsort(pos)], ci[order])

    V, D = 4000, 900
    rp, ci, ct = _random_rows(12, D, V, 0, 700, empty_every=9)
    for d in range(0, D, 3):  # every third row in random order
        seg = slice(rp[d], rp[d + 1])
        perm = rs.permutation(rp[d + 1] - rp[d])
        ci[seg], ct[seg] = ci[seg][perm], ct[seg][perm]
    out = []
    for walk in ("0", "1"):
        monkeypatch.setenv("TTLDA_CSC_WALK", walk)
        c = DeviceCorpus.from_host_csr(rp, ci, ct, V, dev)
        c.build_csc(2)
        out.append([t.clone() for t in (c.colptr, c.docidx[:c.nnz], c.ccounts[:c.nnz], c.csr2csc[:c.nnz])])
    for a, b in zip(*out):
        assert torch.equal(a, b)


def test_cfg5_class_shapes_plan_the_windowed_pass():
    """The planner picks the windowed pass where rows are long (cfg4: K = 500; cfg5: K = 1000, where whole-row blocks
    could not be afforded at all) and keeps whole rows at cfg3."""
    from tuotuo_b200.corpus import DeviceCorpus

    def plan(D, V, ld, ok=True):
        c = DeviceCorpus.__new__(DeviceCorpus)
        c.D, c.V = D, V
        nb, bd = c.plan_blocks(ld, windowed_ok=ok)
        return c.kwindow, nb, bd

    assert plan(1_250_000, 100_000, 1000) == (128, 25, 50_000)
    assert plan(1_250_000, 100_000, 1000, ok=False)[:2] == (0, 1)
    assert plan(1_000_000, 50_000, 100)[:2] == (0, 16)
    assert plan(300_000, 102_660, 500)[:2] == (128, 6)
    assert plan(300_000, 102_660, 500, ok=False)[:2] == (0, 23)


@pytest.mark.parametrize("G,K", [("2", 30), ("4", 60), ("8", 100), ("16", 100), ("16", 250), ("32", 300)])
def test_every_group_width_is_exact(monkeypatch, G, K):
    """TTLDA_GROUP forces the lanes-per-K-vector mapping; every width must give the same answer as the oracle."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    V = 500
    rp, ci, ct = numpy_corpus(33, 150, V, 12, 45, ragged=True)
    o = OracleLDA(K, engine="c")
    po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=2)
    monkeypatch.setenv("TTLDA_GROUP", G)
    m = LDASmoothed(K, verbose=False)
    pg = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=2, return_perplexities=True)
    assert relerr(pg, po) < TOL_PERP and relerr(m._gamma_, o.gamma) < TOL_STATE and relerr(m._lambda_, o.lam) < TOL_STATE


@pytest.mark.parametrize("K,V,D,L", [(10, 300, 120, 50), (100, 1200, 150, 90), (300, 900, 40, 120), (1000, 1100, 12, 100)])
def test_inner_fixed_point_mode_vs_oracle(K, V, D, L):
    """Opt-in true fixed point (NOT the reference's behaviour): exp(E[log theta_d]) refreshed inside the
    per-document loop until ||dgamma||_1 <= 1e-5.  Checked against the oracle's restatement of the same mode."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(900 + K, D, V, min(K, 25), L, ragged=True)
    o = OracleLDA(K, engine="c", inner_fixed_point=True)
    po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=3)
    m = LDASmoothed(K, verbose=False, inner_fixed_point=True)
    pg = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, return_perplexities=True)
    assert o.inner_iters > 4 * 3 * D  # the loop really iterates (the frozen-theta reference makes 2 passes)
    assert relerr(m._last_elbo, o.last_elbo) < 1e-10 and relerr(pg, po) < TOL_PERP
    assert relerr(m._gamma_, o.gamma) < 1e-8 and relerr(m._lambda_, o.lam) < 1e-8
    # and it is a different (better-converged) answer than the reference-parity mode: higher ELBO after 3 EM steps
    # (ELBOs, not perplexities: exp() overflows to inf on the tiny K*V >> tokens corpora used here)
    fp = LDASmoothed(K, verbose=False, inner_fixed_point=True)
    fp.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, update_hyper_parms=False)
    ref = LDASmoothed(K, verbose=False)
    ref.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, update_hyper_parms=False)
    assert fp._last_elbo > ref._last_elbo and fp.inner_iterations > ref.inner_iterations == 2 * D


def test_ratio_handoff_and_recompute_agree(monkeypatch):
    """The statistics pass fed with the E-step's c/N ratios equals the pass that recomputes the normaliser."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(55, 500, 700, 20, 60, ragged=True)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TTLDA_RATIO_HANDOFF", flag)
        m = LDASmoothed(64, verbose=False)
        p = m.fit(Document.from_csr(rp, ci, ct, 700), em_num_step=3, return_perplexities=True)
        assert (m._st.ratio is not None) == (flag == "1")
        res[flag] = (p, m._lambda_.copy(), m._gamma_.copy())
    assert relerr(res["1"][0], res["0"][0]) < 1e-13
    assert relerr(res["1"][1], res["0"][1]) < 1e-12 and relerr(res["1"][2], res["0"][2]) < 1e-12


@pytest.mark.parametrize("blocks", [1, 3, 8])
@pytest.mark.parametrize("handoff", ["1", "0"])
def test_em_step_from_host_equals_resident_iteration(monkeypatch, blocks, handoff):
    """Streaming entry point: the corpus arrives from pinned host memory one document block at a time, each block
    mirrored / E-stepped / accumulated while the next one is in flight; results equal the resident-corpus iterations
    of fit() — bit for bit (same segments, same summation order)."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    monkeypatch.setenv("TTLDA_DOC_BLOCKS", str(blocks))
    monkeypatch.setenv("TTLDA_RATIO_HANDOFF", handoff)
    rp, ci, ct = numpy_corpus(61, 900, 600, 20, 50, ragged=True)
    ref = LDASmoothed(40, verbose=False)
    pr = ref.fit(Document.from_csr(rp, ci, ct, 600), em_num_step=3, update_hyper_parms=False, return_perplexities=True)
    m = LDASmoothed(40, verbose=False)
    m.fit(Document.from_csr(rp, ci, ct, 600), em_num_step=0, update_hyper_parms=False)
    c = m._st.corpus
    assert c.nblocks == blocks
    h = [torch.as_tensor(a).pin_memory() for a in (rp, ci, ct)]
    mirror = [t.clone() for t in (c.colptr, c.docidx, c.ccounts, c.csr2csc)]
    # scribble over the device corpus and its mirror: the streamed copy must be what the iteration reads
    for t in (c.colidx, c.counts, c.docidx, c.ccounts, c.csr2csc, c.colptr):
        t.fill_(-1)
    ps = [m.em_step_from_host(*h) for _ in range(3)]
    c = m._st.corpus  # the streamed step works in a private copy of the corpus buffers, owned by the fit state
    # block-by-block mirror == one-call blocked mirror (with the ratio hand-off the mirrored counts are not rebuilt)
    got = [c.colptr, c.docidx, c.csr2csc if handoff == "1" else c.ccounts]
    want = [mirror[0], mirror[1], mirror[3] if handoff == "1" else mirror[2]]
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    assert c.ccounts_valid == (handoff == "0")
    assert relerr(ps, pr[:3]) < 1e-13  # perplexity of the incoming state of each iteration
    assert np.array_equal(m._gamma_, ref._gamma_) and np.array_equal(m._lambda_, ref._lambda_)
    assert m.inner_iterations == 2 * 900


@pytest.mark.parametrize("ci_dtype,ct_dtype", [(np.uint16, np.uint8), (np.int32, np.uint16), (np.uint16, np.int32)])
def test_em_step_from_host_compact_formats(monkeypatch, ci_dtype, ct_dtype):
    """16-bit word ids / 8- or 16-bit counts in host memory are widened on the device (ttlda_widen_i32): same bits
    as the int32 form.  Word ids above 32767 and counts above 127 exercise the unsigned interpretation."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    monkeypatch.setenv("TTLDA_DOC_BLOCKS", "3")
    V = 40000
    rp, ci, ct = numpy_corpus(67, 400, V, 10, 120, ragged=True)
    ci = (ci.astype(np.int64) * 7 % V).astype(np.int32)  # spread ids over [0, V), then restore the per-row order
    for d in range(len(rp) - 1):
        seg = slice(rp[d], rp[d + 1])
        u, first = np.unique(ci[seg], return_index=True)
        assert len(u) == rp[d + 1] - rp[d]
        ct[seg] = ct[seg][first]
        ci[seg] = u
    ct = ct.copy()
    ct[::5] = 200  # > 127
    assert ci.max() > 32767
    res = []
    for a_ci, a_ct in ((ci, ct), (ci.astype(ci_dtype), ct.astype(ct_dtype))):
        m = LDASmoothed(24, verbose=False)
        m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=0, update_hyper_parms=False)
        h = [torch.as_tensor(rp).pin_memory()]
        for a in (a_ci, a_ct):  # torch has no from_numpy for uint16 everywhere: hand over the bit pattern
            t = torch.as_tensor(a.view(np.int16) if a.dtype == np.uint16 else a)
            h.append(t.pin_memory())
        ps = [m.em_step_from_host(*h) for _ in range(2)]
        res.append((ps, m._gamma_.copy(), m._lambda_.copy()))
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


def test_block_entry_points_equal_one_call_forms(nv, dev):
    """ttlda_csr_to_csc_block / ttlda_sstats_csc_block_f64 + ttlda_sstats_fold_f64 over every block reproduce
    ttlda_csr_to_csc / ttlda_sstats_csc_f64 with nblocks > 1 exactly (blocks that do not start on the chunk grid,
    an empty block, with and without the c/N ratios)."""
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(7, 1300, 300, 5, 60, ragged=True)
    D, V, K = 1300, 300, 20
    ld = nv.ld_for(K)
    bd = 300
    nb = -(-D // bd)
    lens = np.diff(rp)
    lens[bd:2 * bd] = 0  # a second corpus whose block 1 is empty
    rp2 = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    keep = np.ones(len(ci), bool)
    keep[rp[bd]:rp[2 * bd]] = False
    ci2, ct2 = ci[keep], ct[keep]
    for rowptr, colidx, counts in ((rp, ci, ct), (rp2, ci2, ct2)):
        nnz = len(colidx)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
        d_rp, d_ci, d_ct = t(rowptr), t(colidx), t(counts)
        i32 = dict(dtype=torch.int32, device=dev)
        ref = [torch.empty(nb * V + 1, dtype=torch.int64, device=dev), torch.empty(nnz, **i32), torch.empty(nnz, **i32),
               torch.empty(nnz, **i32)]
        ws = torch.empty(nv.csr_to_csc_workspace_bytes(nnz, D, V, nb), dtype=torch.uint8, device=dev)
        nv.csr_to_csc(d_rp, d_ci, d_ct, D, V, nnz, nb, bd, ref[0], ref[1], ref[2], ws, ref[3])
        blk = [torch.full_like(x, -7) for x in ref]
        for b in range(nb):
            d0, d1 = b * bd, min((b + 1) * bd, D)
            s, e = int(rowptr[d0]), int(rowptr[d1])
            nv.csr_to_csc_block(d_rp[d0:d1 + 1], d_ci, d_ct, V, d0, s, e - s, blk[0][b * V:b * V + V + 1], blk[1], blk[2],
                                ws, blk[3])
        for a, b_ in zip(ref, blk):
            assert torch.equal(a, b_)

        g = torch.Generator(device="cpu").manual_seed(3)
        eT = torch.zeros((D, ld), dtype=torch.float64)
        eT[:, :K] = torch.rand((D, K), generator=g, dtype=torch.float64) + 0.1
        eB = torch.zeros((V, ld), dtype=torch.float64)
        eB[:, :K] = torch.rand((V, K), generator=g, dtype=torch.float64) + 0.1
        eT, eB = eT.to(dev), eB.to(dev)
        ratio = (torch.rand(max(nnz, 1), generator=g, dtype=torch.float64) + 0.5).to(dev)
        ws_s = torch.empty(nv.sstats_workspace_bytes(nnz, V, ld, nb), dtype=torch.uint8, device=dev)
        for r in (None, ratio):
            S_ref = torch.empty((V, ld), dtype=torch.float64, device=dev)
            nv.sstats_csc(ref[0], ref[1], ref[2], V, K, nnz, nb, eT, eB, S_ref, ws_s, r)
            Sb = torch.full((nb, V, ld), float("nan"), dtype=torch.float64, device=dev)
            for b in range(nb):
                d0, d1 = b * bd, min((b + 1) * bd, D)
                s, e = int(rowptr[d0]), int(rowptr[d1])
                nv.sstats_csc_block(ref[0][b * V:b * V + V + 1], ref[1], ref[2], V, K, s, e - s, eT, eB, Sb[b], ws_s, r)
            S = torch.empty_like(S_ref)
            nv.sstats_fold(Sb, S)
            assert torch.equal(S, S_ref)
            # serial form: the blocks accumulate into ONE table, no per-block buffers — same bits again
            S2 = torch.full_like(S_ref, float("nan"))
            for b in range(nb):
                d0, d1 = b * bd, min((b + 1) * bd, D)
                s, e = int(rowptr[d0]), int(rowptr[d1])
                nv.sstats_csc_block(ref[0][b * V:b * V + V + 1], ref[1], ref[2], V, K, s, e - s, eT, eB, S2, ws_s, r,
                                    accumulate=b > 0)
            assert torch.equal(S2, S_ref)


@pytest.mark.parametrize("handoff", ["1", "0"])
def test_serial_block_statistics_equal_folded(monkeypatch, handoff):
    """Serial-block mode (one launch per document block accumulating in place: the cfg5-class form without per-block
    buffers) gives the bits of the folded one-launch form, in fit() and in the streamed step."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    monkeypatch.setenv("TTLDA_RATIO_HANDOFF", handoff)
    monkeypatch.setenv("TTLDA_DOC_BLOCKS", "5")
    rp, ci, ct = numpy_corpus(71, 700, 900, 15, 70, ragged=True)
    res = {}
    for serial in ("0", "1"):
        monkeypatch.setenv("TTLDA_SERIAL_BLOCKS", serial)
        m = LDASmoothed(36, verbose=False)
        p = m.fit(Document.from_csr(rp, ci, ct, 900), em_num_step=3, return_perplexities=True)
        assert m._st.corpus.serial_blocks == (serial == "1") and m._st.corpus.nblocks == 5
        h = [torch.as_tensor(a).pin_memory() for a in (rp, ci, ct)]
        p.append(m.em_step_from_host(*h))
        res[serial] = (p, m._lambda_.copy(), m._gamma_.copy())
    assert res["0"][0] == res["1"][0]
    assert np.array_equal(res["0"][1], res["1"][1]) and np.array_equal(res["0"][2], res["1"][2])


@pytest.mark.parametrize("K,V", [(12, 300), (100, 800)])
def test_partial_fit_online_vs_oracle(K, V):
    """Online VB (partial_fit): three minibatches of different sizes, topics persisting across calls, against the
    oracle's restatement — gamma, lambda, minibatch perplexity bound; then the topics keep serving approx_perplexity."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    m = LDASmoothed(K, verbose=False)
    o = OracleLDA(K, inner_fixed_point=True)
    total = 5000
    for i, D in enumerate((150, 90, 220)):
        rp, ci, ct = numpy_corpus(100 + i, D, V, min(K, 10), 60, ragged=True)
        p = m.partial_fit(Document.from_csr(rp, ci, ct, V), total_docs=total, tau0=2.0, kappa=0.6, return_perplexity=True)
        po = o.partial_fit(CSRCorpus(rp, ci, ct, V), total_docs=total, tau0=2.0, kappa=0.6)
        assert relerr(p, po) < 1e-8
        assert relerr(m._gamma_, o.gamma) < 1e-9 and relerr(m._lambda_, o.lam) < 1e-9
    assert m._online_t == 3 and m._gamma_.shape == (220, K)
    # asynchronous form (no perplexity, nothing read back) gives the same update
    m2 = LDASmoothed(K, verbose=False)
    for i, D in enumerate((150, 90, 220)):
        rp, ci, ct = numpy_corpus(100 + i, D, V, min(K, 10), 60, ragged=True)
        assert m2.partial_fit(Document.from_csr(rp, ci, ct, V), total_docs=total, tau0=2.0, kappa=0.6) is None
    assert np.array_equal(m2._lambda_, m._lambda_)
    with pytest.raises(ValueError):
        m.partial_fit(Document.from_csr(rp, ci, ct, V), kappa=0.5)
    with pytest.raises(ValueError):
        m.partial_fit(Document.from_csr(rp, ci, ct, V + 1))


def test_nonexchangeable_update_alpha_vs_reference(golden_dir):
    """update_alpha with a (1,K) array prior: device per-topic digamma statistics (ttlda_hyper_colstats_f64) + the
    reference's Newton step on the host, against the real reference's first iterations; and fit() now runs through
    the hyper update with such a prior like the reference does."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    z = np.load(os.path.join(golden_dir, "hyper_nonexch.npz"))
    V, K = int(z["V"]), int(z["K"])
    m = LDASmoothed(K, exchangeable_prior=False, verbose=False)
    m.fit(Document.from_csr(z["rowptr"], z["colidx"], z["counts"], V), em_num_step=2, update_hyper_parms=False)
    assert relerr(m._gamma_, z["gamma"]) < 1e-9
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m.update_alpha(step=int(z["newton_steps"]))
    assert len(w) == int(z["n_warnings"])
    assert m._alpha_.shape == z["alpha"].shape and relerr(m._alpha_, z["alpha"]) < 1e-9
    o = OracleLDA(K, exchangeable_prior=False)
    o.fit(CSRCorpus(z["rowptr"], z["colidx"], z["counts"], V), em_num_step=2, update_hyper_parms=False)
    o.update_alpha(step=int(z["newton_steps"]))
    assert relerr(m._alpha_, o.alpha) < 1e-10


EDGE = [
    # (K, V, rows as lists of (word, count))
    (1, 1, [[(0, 3)]]),                                   # one topic, one word, one document
    (1, 5, [[(0, 1), (4, 2)], [], [(2, 7)]]),             # K = 1 with an empty document
    (3, 4, [[], [], []]),                                 # nothing but empty documents (nnz = 0)
    (7, 9, [[(8, 1)]] * 40),                              # identical one-word documents: one long posting list
    (5, 3, [[(0, 1000000), (1, 1), (2, 1)]]),             # a huge count
    (33, 70, [[(w, 1 + (w * d) % 3) for w in range(d % 70, 70, 1 + d % 5)] for d in range(65)]),  # ld=36, ragged
    (4, 2000, [[(w, 1) for w in range(0, 2000, 3)]] * 3),  # rows far longer than a 32-item batch
]


@pytest.mark.parametrize("K,V,rows", EDGE)
@pytest.mark.parametrize("fixed_point", [False, True])
def test_edge_cases_vs_oracle(K, V, rows, fixed_point):
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    rp = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64)
    ci = np.array([w for r in rows for w, _ in r], dtype=np.int32)
    ct = np.array([c for r in rows for _, c in r], dtype=np.int32)
    o = OracleLDA(K, engine="c", inner_fixed_point=fixed_point)
    m = LDASmoothed(K, verbose=False, inner_fixed_point=fixed_point)
    if ct.sum() == 0:
        # an all-empty corpus: the reference divides by a zero token count; both sides must agree on nan/inf
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=2, update_hyper_parms=False)
            pg = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=2, update_hyper_parms=False,
                       return_perplexities=True)
        assert len(pg) == len(po) and relerr(m._gamma_, o.gamma) < TOL_STATE and relerr(m._lambda_, o.lam) < TOL_STATE
        return
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=3, update_hyper_parms=False)
        pg = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, update_hyper_parms=False, return_perplexities=True)
    assert len(pg) == len(po)
    if abs(o.last_elbo) > 1e-6:
        assert relerr(m._last_elbo, o.last_elbo) < 1e-10 and relerr(pg, po) < TOL_PERP
    else:  # K = V = 1: the ELBO is exactly 0 and the reference's exponent clip (utils.py:20-26) is discontinuous there
        assert abs(m._last_elbo - o.last_elbo) < 1e-11
    tol = 1e-8 if fixed_point else TOL_STATE
    assert relerr(m._gamma_, o.gamma) < tol and relerr(m._lambda_, o.lam) < tol


@pytest.mark.parametrize("K", [70, 100, 128])
def test_tma_gather4_variants_are_exact(monkeypatch, K):
    """TTLDA_TMA=1: K-vector rows fed by the TMA engine (cp.async.bulk.tensor ... tile::gather4 + mbarrier pipeline)
    in the E-step, the ELBO pass and the statistics pass; same answers as the oracle."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    V = 900
    rp, ci, ct = numpy_corpus(700 + K, 300, V, 25, 70, ragged=True)
    o = OracleLDA(K, engine="c")
    po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=3)
    monkeypatch.setenv("TTLDA_TMA", "1")
    m = LDASmoothed(K, verbose=False)
    pg = m.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, return_perplexities=True)
    assert relerr(pg, po) < TOL_PERP and relerr(m._gamma_, o.gamma) < TOL_STATE and relerr(m._lambda_, o.lam) < TOL_STATE
    monkeypatch.setenv("TTLDA_FUSED_REFRESH", "1")  # MODE 1 kernel as well
    m2 = LDASmoothed(K, verbose=False)
    p2 = m2.fit(Document.from_csr(rp, ci, ct, V), em_num_step=3, return_perplexities=True)
    assert relerr(p2, po) < TOL_PERP and relerr(m2._gamma_, o.gamma) < TOL_STATE
