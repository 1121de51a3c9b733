"""N-GPU == oracle check, to be launched with torchrun on a multi-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_multigpu_check.py

Every rank fits its contiguous document shard (Document.shard) with the CUDA kernels; the K x V statistics are
summed with one NCCL all-reduce per EM iteration.  Rank 0 compares perplexities / gamma / lambda / alpha / eta with
the single-process CPU oracle on the whole corpus (<= 1e-9 / 1e-8, north-star tolerances) and prints one JSON line.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tuotuo_b200.document import Document
    from tuotuo_b200.lda_model import LDASmoothed
    from tuotuo_b200.synth import numpy_corpus

    results = {}
    for K, V, D, L, steps in [(20, 800, 600, 80, 4), (100, 1500, 400, 100, 3), (500, 1200, 96, 150, 2)]:
        rp, ci, ct = numpy_corpus(500 + K, D, V, min(K, 30), L, ragged=True)
        full = Document.from_csr(rp, ci, ct, V)
        m = LDASmoothed(K, verbose=False, device=dev)
        perps = m.fit(full.shard(rank, world), em_num_step=steps, return_perplexities=True)
        gam, lam = m._gamma_, m._lambda_
        if rank == 0:
            from lda_oracle import CSRCorpus, OracleLDA

            o = OracleLDA(K, engine="c")
            po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=steps)
            e = dict(perp=relerr(perps, po), gamma=relerr(gam, o.gamma), lam=relerr(lam, o.lam),
                     alpha=relerr(m._alpha_, o.alpha), eta=relerr(m._eta_, o.eta), n_perp=len(perps))
            assert e["perp"] < 1e-8 and e["gamma"] < 1e-9 and e["lam"] < 1e-9 and e["alpha"] < 1e-9 and e["eta"] < 1e-9, e
            assert gam.shape == (D, K)
            results[f"K{K}"] = e
        # lambda must be bitwise identical on every rank (same reduced buffer, same M-step)
        h = torch.tensor([float(np.sum(lam)), float(perps[-1])], dtype=torch.float64, device=dev)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        assert all(torch.equal(hs[0], x) for x in hs), "ranks disagree"
    if rank == 0:
        print(json.dumps({"multigpu_check": "ok", "world": world, "max_rel_err": results}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
