"""GPU parity at the shapes BASELINE.json names (configs[1], [3], [4]) and the N > 1 path on real GPUs.

* cfg2 IN FULL (10 000 documents x 200 tokens, K=20, V=5k): default `fit` for 3 EM iterations + hyper update against
  the C oracle (which is pinned to the real reference by tests/test_oracle_golden.py);
* a cfg4-shaped slice: K=500 WITH V=102 660, Poisson(330) lengths, a document-blocked mirror;
* a cfg5-shaped slice: K=1000 WITH V=100 000, `sampling=True`, D_population != D;
* world_size 2 on two GPUs (skipped on a one-GPU box): Document.shard + NCCL all-reduce vs the oracle.
Tolerances are north_star's: <= 1e-9 on gamma / lambda / alpha / eta, <= 1e-8 on every perplexity.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from lda_oracle import CSRCorpus, OracleLDA

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_STATE, TOL_PERP = 1e-9, 1e-8


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        d = np.where(a == b, 0.0, np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    return float(np.max(d)) if d.size else 0.0


def host_corpus(config, docs, seed):
    """The bench generator's documents, counted on the host (the oracle's own counting): identical input for both."""
    from tuotuo_b200.synth import CONFIGS, synthetic_token_stream

    D, L, K, V, pois = CONFIGS[config]
    tok, offs = synthetic_token_stream(docs, L, K, V, device="cpu", seed=seed, poisson=pois)
    return CSRCorpus.from_token_ids(tok.numpy(), offs.numpy(), V), tok, offs, K, V


def check(m, o, p, po):
    assert len(p) == len(po) and relerr(p, po) < TOL_PERP
    assert relerr(m._gamma_, o.gamma) < TOL_STATE and relerr(m._lambda_, o.lam) < TOL_STATE
    assert relerr(m._alpha_, o.alpha) < TOL_STATE and relerr(m._eta_, o.eta) < TOL_STATE


def test_cfg2_full_fit_vs_oracle():
    """BASELINE.json configs[1] at its full size, through Document.from_token_ids (ttlda_bow_to_csr counts on the GPU):
    default `fit` for 3 EM iterations, then the hyper-parameter step.  On this corpus the reference's Newton step drives
    eta below zero at its first iteration and raises (tuotuo/lda_model.py:539-540) — the oracle does, so must the GPU
    build, with alpha updated to the same value before that."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    X, tok, offs, K, V = host_corpus("cfg2", 10_000, seed=1234)
    assert X.D == 10_000 and X.tokens == 10_000 * 200
    doc = Document.from_token_ids(tok.numpy(), offs.numpy(), V)
    m = LDASmoothed(K, verbose=False)
    p = m.fit(doc, em_num_step=3, update_hyper_parms=False, return_perplexities=True)
    rp, ci, ct = doc.host_csr()
    assert np.array_equal(rp, X.rowptr) and np.array_equal(ci, X.colidx) and np.array_equal(ct, X.counts)  # bit-exact
    o = OracleLDA(K, engine="c")
    po = o.fit(X, em_num_step=3, update_hyper_parms=False)
    check(m, o, p, po)
    assert m.inner_iterations == 2 * X.D  # the reference's frozen-theta loop: exactly two passes per document
    # the hyper-parameter step of the default fit (lda_model.py:414-424)
    o.update_alpha()
    m.update_alpha()
    assert relerr(m._alpha_, o.alpha) < TOL_STATE
    try:
        o.update_eta()
        oracle_raises = False
    except ValueError:
        oracle_raises = True
    if oracle_raises:
        with pytest.raises(ValueError, match="Dirichlet Parameter become < 0"):
            m.update_eta()
    else:
        m.update_eta()
    assert relerr(m._eta_, o.eta) < TOL_STATE


def test_cfg4_shaped_slice_vs_oracle(monkeypatch):
    """K=500 together with V=102 660 (G=32 lane groups, 411 MB beta tables), ragged Poisson(330) documents, three
    document blocks in the mirror (per-block partial statistics + fold)."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    monkeypatch.setenv("TTLDA_DOC_BLOCKS", "3")
    X, tok, offs, K, V = host_corpus("cfg4", 600, seed=77)
    assert (K, V) == (500, 102_660) and len(set(np.diff(offs.numpy()).tolist())) > 10
    m = LDASmoothed(K, verbose=False)
    p = m.fit(Document.from_csr(X.rowptr, X.colidx, X.counts, V), em_num_step=2, return_perplexities=True)
    assert m._st.corpus.nblocks == 3 and m._st.ratio is not None
    o = OracleLDA(K, engine="c")
    po = o.fit(X, em_num_step=2)
    check(m, o, p, po)


def test_cfg5_shaped_slice_sampling_vs_oracle():
    """K=1000 together with V=100 000 (G=32, E=16: the widest instantiation), `sampling=True` with a population
    (10M documents) different from the number of documents fitted (lda_model.py:152-153,187-188)."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed

    X, tok, offs, K, V = host_corpus("cfg5", 300, seed=78)
    assert (K, V) == (1000, 100_000)
    doc = Document.from_csr(X.rowptr, X.colidx, X.counts, V, num_doc_population=10_000_000)
    m = LDASmoothed(K, verbose=False)
    p = m.fit(doc, sampling=True, em_num_step=2, return_perplexities=True)
    assert m.D_population == 10_000_000
    o = OracleLDA(K, engine="c")
    o.D_population = 10_000_000
    po = o.fit(X, sampling=True, em_num_step=2)
    check(m, o, p, po)
    # and the unsampled perplexities differ: the rescale is really applied
    m2 = LDASmoothed(K, verbose=False)
    p2 = m2.fit(doc, sampling=False, em_num_step=1, update_hyper_parms=False, return_perplexities=True)
    assert abs(p2[0] - p[0]) / p[0] > 1e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_match_oracle(tmp_path):
    """tests/run_multigpu_check.py under torch.distributed.run on 2 GPUs: K in {20, 100, 500}, Document.shard, NCCL
    all-reduce of the statistics, rank 0 vs the CPU oracle at the north-star tolerances, lambda bitwise equal on both."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ)
    env.pop("TTLDA_DOC_BLOCKS", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "run_multigpu_check.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["multigpu_check"] == "ok" and rec["world"] == 2
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "multigpu_check_n2.json"), "w") as f:
        f.write(line + "\n")


def test_fit_on_a_non_current_device():
    """LDASmoothed(device=...) is honoured whatever the process's current device is: every launch goes to the device the
    tensors live on (_native._call).  With one GPU this checks the explicit-index path; with two, a fit on cuda:1 while
    cuda:0 is current."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    target = torch.device("cuda", torch.cuda.device_count() - 1)
    torch.cuda.set_device(0)
    rp, ci, ct = numpy_corpus(3, 90, 300, 8, 50, ragged=True)
    m = LDASmoothed(16, verbose=False, device=target)
    p = m.fit(Document.from_csr(rp, ci, ct, 300), em_num_step=2, return_perplexities=True)
    assert torch.cuda.current_device() == 0 and m._st.gamma[0].device == target
    o = OracleLDA(16, engine="c")
    po = o.fit(CSRCorpus(rp, ci, ct, 300), em_num_step=2)
    check(m, o, p, po)


def test_malformed_corpora_raise_value_error():
    """Indices that would drive out-of-bounds gathers are refused at ingestion (a NumPy IndexError in the reference)."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    rp, ci, ct = numpy_corpus(5, 40, 200, 6, 30)
    V = 200
    bad = ci.copy()
    bad[7] = V
    with pytest.raises(ValueError):
        Document.from_csr(rp, bad, ct, V)
    bad = rp.copy()
    bad[5] = bad[6] + 1
    with pytest.raises(ValueError):
        Document.from_csr(bad, ci, ct, V)
    with pytest.raises(ValueError):
        Document.from_csr(rp, ci, -ct, V)
    # token-id streams are checked on the device (ttlda_validate_csr) before they become sort keys
    tok = np.array([0, 3, V + 5, 2], dtype=np.int32)
    with pytest.raises(ValueError):
        LDASmoothed(4, verbose=False).fit(Document.from_token_ids(tok, np.array([0, 2, 4]), V), em_num_step=1)
    with pytest.raises(ValueError):
        LDASmoothed(4, verbose=False).fit(Document.from_token_ids(tok[:2], np.array([0, 3, 2]), V), em_num_step=1)


def test_streamed_step_validates_and_does_not_alias_the_document():
    """em_step_from_host works in buffers owned by the fit state: the Document's cached device corpus (and a later
    fit() on it) is untouched; a malformed streamed corpus raises ValueError before anything is committed."""
    from tuotuo.document import Document
    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    K, V, D = 24, 400, 120
    rp, ci, ct = numpy_corpus(21, D, V, 8, 40)
    rp2, ci2, ct2 = numpy_corpus(22, D, V, 8, 40)  # fixed length => same D; nnz differs
    doc = Document.from_csr(rp, ci, ct, V)
    m = LDASmoothed(K, verbose=False)
    m.fit(doc, em_num_step=1, update_hyper_parms=False)
    n = min(len(ci), len(ci2))
    if len(ci2) > len(ci):  # keep the streamed corpus within the fitted buffers: drop the tail of the last document
        ci2, ct2 = ci2[:n], ct2[:n]
        rp2 = np.minimum(rp2, n)
    t = lambda a: torch.as_tensor(a).pin_memory()
    lam_before = m._lambda_.copy()
    m.em_step_from_host(t(rp2), t(ci2.astype(np.int32)), t(ct2.astype(np.int32)))
    assert not np.array_equal(m._lambda_, lam_before)
    a, b, c = doc.host_csr()
    assert np.array_equal(a, rp) and np.array_equal(b, ci) and np.array_equal(c, ct)
    dc = doc.device_corpus(m._st.dev)
    assert np.array_equal(dc.colidx.cpu().numpy(), ci) and m._st.corpus is not dc
    # malformed: a word id >= V in compact and in int32 form; a decreasing rowptr
    lam_before = m._lambda_.copy()
    badci = ci2.astype(np.int32).copy()
    badci[3] = V + 1
    with pytest.raises(ValueError):
        m.em_step_from_host(t(rp2), t(badci), t(ct2.astype(np.int32)))
    with pytest.raises(ValueError):
        m.em_step_from_host(t(rp2), t(badci.astype(np.uint16).view(np.int16)), t(ct2.astype(np.uint8)))
    badrp = rp2.copy()
    badrp[4] = badrp[5] + 2
    with pytest.raises(ValueError):
        m.em_step_from_host(t(badrp), t(ci2.astype(np.int32)), t(ct2.astype(np.int32)))
    assert np.array_equal(m._lambda_, lam_before)  # nothing was committed
    # and the model still steps on its fitted corpus afterwards
    p, committed = m.em_iteration()
    assert committed and np.isfinite(p)


def test_partial_fit_keeps_one_vocabulary_and_fitted_topics():
    """Raw-text minibatches are encoded against the vocabulary of the first corpus (not re-numbered by first
    appearance), and partial_fit after fit() updates the FITTED topics instead of re-seeding lambda."""
    from tuotuo.document import Document
    from tuotuo.generator import doc_generator
    from tuotuo.lda_model import LDASmoothed

    gen = doc_generator(M=60, L=25, topic_prior=torch.full((5,), 0.3, dtype=torch.double))
    docs = gen.generate_doc()
    keys = list(docs)
    first = {k: docs[k] for k in keys[:30]}
    second = {k: " ".join(reversed(docs[k].split(" "))) for k in keys[30:]}  # another first-appearance order
    m = LDASmoothed(5, verbose=False)
    m.fit(Document(first), em_num_step=3, update_hyper_parms=False)
    lam_fit = m._lambda_.copy()
    vocab = dict(m.train_doc.vocab_to_idx)
    m.partial_fit(second, total_docs=60, tau0=64.0, kappa=0.7)
    assert m.train_doc.vocab_to_idx == vocab
    lam = m._lambda_
    assert lam.shape == lam_fit.shape
    # rho = 64^-0.7 ~ 0.054: the fitted topics moved a little, they were not replaced by a fresh Gamma(100, .01) draw
    np.random.seed(42)
    fresh = np.random.gamma(100, 0.01, lam.shape)
    assert 0 < np.linalg.norm(lam - lam_fit) < 0.5 * np.linalg.norm(fresh - lam_fit)
    m.update_alpha()  # M is known on this path too


@pytest.mark.parametrize("cfg", ["cfg3", "cfg4"])
def test_full_size_properties(cfg):
    """BASELINE.json configs[2] and configs[3] at FULL size (1M documents x 300 tokens, K = 100, V = 50k; 300k documents x
    Poisson(330), K = 500, V = 102 660) — where no oracle can follow — through size-independent properties of the path:
      * the mirror is the transpose of the CSR: csr2csc is a permutation, inside every (block, word) segment the documents
        ascend, every entry's word matches its segment, the counts add up;
      * sum_k (gamma_dk - alpha) = length of document d, for every document (lda_model.py:248 sums phi over k to 1);
      * sum_wk (lambda_wk - eta) = number of tokens (lda_model.py:262-264); the token count of the ELBO record agrees;
      * the perplexity bound decreases over the first EM iterations;
      * a second fit of the same corpus reproduces gamma, lambda and every perplexity bit for bit."""
    import torch

    from tuotuo.lda_model import LDASmoothed
    from tuotuo_b200.synth import CONFIGS, synthetic_document

    dev = torch.device("cuda", 0)
    D, L, K, V, pois = CONFIGS[cfg]
    doc = synthetic_document(cfg, device=dev)
    m = LDASmoothed(K, verbose=False)
    perps = m.fit(doc, em_num_step=3, update_hyper_parms=False, return_perplexities=True)
    st = m._st
    c = st.corpus
    assert c.D == D and c.V == V and c.nblocks > 1
    if not pois:
        assert c.tokens == D * L
    # --- mirror = transpose
    nnz = c.nnz
    pos = c.csr2csc[:nnz].to(torch.int64)
    seen = torch.zeros(nnz, dtype=torch.bool, device=dev)
    seen[pos] = True
    assert bool(seen.all())  # a permutation
    del seen
    seg = torch.searchsorted(c.colptr, torch.arange(nnz, device=dev), right=True) - 1  # (block, word) segment of a slot
    word_of_slot = (seg % V).to(torch.int32)
    assert torch.equal(word_of_slot[pos], c.colidx[:nnz])  # every entry sits in its own word's segment of its block
    rows = torch.repeat_interleave(torch.arange(c.D, device=dev, dtype=torch.int32), torch.diff(c.rowptr))
    assert torch.equal(c.docidx[:nnz][pos], rows)  # ... and the slot names its document
    same = seg[1:] == seg[:-1]
    assert bool((c.docidx[1:nnz][same] > c.docidx[:nnz - 1][same]).all())  # documents ascend inside a segment
    del seg, word_of_slot, same, pos
    # --- gamma rows and lambda total
    lens = torch.zeros(c.D, dtype=torch.float64, device=dev)
    lens.index_add_(0, rows.to(torch.int64), c.counts[:nnz].to(torch.float64))
    alpha = float(np.asarray(m._alpha_).reshape(-1)[0])
    assert float(((st.gamma[st.cur][:, :K].sum(1) - K * alpha) - lens).abs().max()) < 1e-8
    m._sync_lambda()
    tot = float((st.lam[:V, :K] - float(m._eta_)).sum())
    assert abs(tot - c.tokens) / c.tokens < 1e-11
    assert st.terms["count"] == c.tokens
    assert perps[0] > perps[1] > perps[2] > perps[3]
    # --- run-to-run reproducibility
    g1, l1 = st.gamma[st.cur].clone(), st.lam.clone()
    m2 = LDASmoothed(K, verbose=False)
    perps2 = m2.fit(doc, em_num_step=3, update_hyper_parms=False, return_perplexities=True)
    assert list(perps2) == list(perps)
    assert torch.equal(m2._st.gamma[m2._st.cur], g1) and torch.equal(m2._st.lam, l1)
