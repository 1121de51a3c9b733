"""oracle/ref_time.py — time the REAL reference's fit() path on this machine's host cores.

TEST / BENCH INFRASTRUCTURE ONLY (run as a subprocess by ``bench.py``; never imported by ``tuotuo_b200``).

    python oracle/ref_time.py --corpus corpus.npz --K 100 --steps 2 --warmup 1 [--full-fit]

Loads the unmodified reference (``/root/reference/tuotuo`` here, the copy staged by ``oracle/build_ref.py`` in
``oracle/_ref/tuotuo`` on the GPU box; its ``cutils`` is the module compiled from the reference's own .pyx) and runs
the body of ``LDASmoothed.fit``'s loop (tuotuo/lda_model.py:384-409: ``em_step`` + ``approx_perplexity``) on a dense
count matrix built from the CSR arrays in ``corpus.npz`` (rowptr, colidx, counts, V) — or, with ``--full-fit``, the
reference's ``fit()`` itself.  The reference is single-threaded Python/NumPy/SciPy: cores used = 1.
Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import io
import contextlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--corpus", required=True)
    ap.add_argument("--K", type=int, required=True)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--full-fit", action="store_true")
    ap.add_argument("--em-num-step", type=int, default=100)
    a = ap.parse_args()

    import numpy as np

    from ref_harness import DenseCorpus, import_reference, reference_root

    lda_model, _, _, _ = import_reference()
    z = np.load(a.corpus)
    rp, ci, ct, V = z["rowptr"], z["colidx"], z["counts"], int(z["V"])
    D = len(rp) - 1
    X = np.zeros((D, V), dtype=np.float64)
    X[np.repeat(np.arange(D), np.diff(rp)), ci] = ct
    tokens = float(X.sum())
    sink = io.StringIO()
    out = {"root": reference_root(), "docs": D, "V": V, "K": a.K, "doc_tokens": tokens, "cores": 1}
    with contextlib.redirect_stdout(sink):
        m = lda_model.LDASmoothed(num_topics=a.K, verbose=False)
        if a.full_fit:
            t0 = time.perf_counter()
            perps = m.fit(DenseCorpus(X), em_num_step=a.em_num_step, return_perplexities=True)
            dt = time.perf_counter() - t0
            iters = max(1, len(perps) - 2)  # initial + one per EM iteration + the one after the hyper update
            out.update(kind="fit", seconds=dt, em_iterations=iters, perplexity_first=float(perps[0]),
                       perplexity_last=float(perps[-1]), tokens_per_s=tokens * iters / dt)
        else:
            # fit()'s own initialisation (tuotuo/lda_model.py:340-378), then its loop body, timed per iteration
            m.train_doc, m.M, m.V = DenseCorpus(X), D, V
            np.random.seed(42)
            m._lambda_ = np.random.gamma(shape=100, scale=0.01, size=(a.K, V))
            m._gamma_ = m._alpha_ + np.ones((D, a.K), dtype=float) * V / a.K
            perp, logs = m.approx_perplexity(X, False)
            times, perps = [], [float(perp)]
            for i in range(a.warmup + a.steps):
                t0 = time.perf_counter()
                et, eb = m.em_step(X, logs[0], logs[1])  # :403
                perp, logs = m.approx_perplexity(X, False, et, eb)  # :404-409
                dt = time.perf_counter() - t0
                perps.append(float(perp))
                if i >= a.warmup:
                    times.append(dt)
            sec = sum(times) / len(times)
            out.update(kind="em_iteration", seconds_per_iteration=sec, steps=a.steps, warmup=a.warmup,
                       perplexities=perps[:4], tokens_per_s=tokens / sec)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
