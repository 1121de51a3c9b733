"""world_size-2 tests of the multi-rank HOST logic on CPU (gloo): document sharding, the all-reduce of the K x V
statistics and of the ELBO partials, identical convergence decisions on every rank, gathering gamma.  The CUDA
entry points are replaced by the test-only stand-ins of tests/cpu_kernels.py (the product has no CPU path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, kw, out_dir, mode):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import cpu_kernels
    from tuotuo_b200.document import Document
    from tuotuo_b200.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    cpu_kernels.install()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rp, ci, ct = numpy_corpus(77, 180, 400, 10, 50, ragged=True)
        rp = rp.copy()
        full = Document.from_csr(rp, ci, ct, 400, num_doc_population=5000)
        shard = full.shard(rank, world)
        m = LDASmoothed(K, verbose=False, device="cpu", sstats_mode=mode)
        perps = m.fit(shard, return_perplexities=True, **kw)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), perps=np.array(perps), gamma=m._gamma_, lam=m._lambda_,
                 alpha=np.float64(m._alpha_), eta=np.float64(m._eta_), M=m.M, local_docs=shard.M,
                 doc_range=np.array(shard.doc_range))
    finally:
        dist.destroy_process_group()


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("kw,mode", [(dict(em_num_step=4), "csc"), (dict(em_num_step=3, sampling=True), "csc"),
                                     (dict(em_num_step=30, threshold=2.0), "atomic")])
def test_two_ranks_equal_one_rank(tmp_path, kw, mode):
    from lda_oracle import CSRCorpus, OracleLDA
    from tuotuo_b200.synth import numpy_corpus

    K, world = 12, 2
    mp.spawn(_worker, args=(world, _free_port(), K, kw, str(tmp_path), mode), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    # single-process oracle on the whole corpus
    rp, ci, ct = numpy_corpus(77, 180, 400, 10, 50, ragged=True)
    o = OracleLDA(K, engine="c")
    if kw.get("sampling"):
        o.D_population = 5000  # Document.num_doc_population
    po = o.fit(CSRCorpus(rp, ci, ct, 400), **kw)
    for i in range(world):
        assert len(r[i]["perps"]) == len(po)  # same number of iterations: identical convergence decision
        assert relerr(r[i]["perps"], po) < 1e-10
        assert relerr(r[i]["lam"], o.lam) < 1e-10
        assert relerr(r[i]["gamma"], o.gamma) < 1e-10  # _gamma_ gathers all shards in rank order
        assert relerr(r[i]["alpha"], o.alpha) < 1e-10 and relerr(r[i]["eta"], o.eta) < 1e-10
        assert int(r[i]["M"]) == 180
    np.testing.assert_array_equal(r[0]["perps"], r[1]["perps"])  # bitwise identical across ranks
    np.testing.assert_array_equal(r[0]["lam"], r[1]["lam"])
    assert int(r[0]["local_docs"]) + int(r[1]["local_docs"]) == 180
    assert r[0]["doc_range"][1] == r[1]["doc_range"][0]
    if "threshold" in kw:
        assert len(po) < kw["em_num_step"]  # the early-return path was taken, on both ranks alike


def test_single_process_cpu_standins_match_oracle(monkeypatch):
    """The stand-ins themselves (used above) agree with the oracle through the real driver, world_size 1."""
    import cpu_kernels
    from lda_oracle import CSRCorpus, OracleLDA
    from tuotuo_b200 import _native
    from tuotuo_b200.document import Document
    from tuotuo_b200.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    for n in cpu_kernels.NAMES:
        monkeypatch.setattr(_native, n, getattr(cpu_kernels, n))
    rp, ci, ct = numpy_corpus(5, 60, 200, 6, 40, ragged=True)
    m = LDASmoothed(7, verbose=False, device="cpu")
    pg = m.fit(Document.from_csr(rp, ci, ct, 200), em_num_step=3, return_perplexities=True)
    o = OracleLDA(7, engine="c")
    po = o.fit(CSRCorpus(rp, ci, ct, 200), em_num_step=3)
    assert relerr(pg, po) < 1e-11 and relerr(m._gamma_, o.gamma) < 1e-11 and relerr(m._lambda_, o.lam) < 1e-11
