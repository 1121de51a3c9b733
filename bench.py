#!/usr/bin/env python
"""bench.py — doc-tokens/sec per full variational-inference EM iteration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is ONE full EM iteration of the hot path over the resident corpus, made through the PUBLIC call
`LDASmoothed.em_iteration()` — the body of `LDASmoothed.fit`'s loop, which `fit` itself calls: E-step (gamma, with the
ELBO token term of the incoming state as a by-product of the same dot products), the device->host read of the
perplexity terms and the host-side convergence test, K x V sufficient statistics, [their exchange across ranks, fused
into the vocabulary-sharded M-step, when N > 1], theta refresh, M-step (lambda + beta refresh); every step yields one
perplexity.  Workload at N=1: BASELINE.json configs[2] "synthetic 1M docs x 300 tokens, K=100, V=50k" (cfg3), the
configuration the roofline target is quoted on; for N > 1 every rank holds its own cfg3-sized shard of documents drawn
from the SAME generative model (weak scaling: rank 0's shard is the N=1 corpus), exp(E[log beta]) is replicated and lambda
sharded over the vocabulary.

Prints ONE JSON line (rank 0).  `value` is the device-timed whole-job throughput with the corpus resident in HBM;
`e2e` repeats the measurement through `LDASmoothed.em_step_from_host` with the CSR corpus coming from pinned HOST
memory every step (H2D copy + CSC rebuild + EM iteration + D2H of the result inside the timed region).
`roofline` describes the dominant kernel against a gather ceiling MEASURED IN THIS RUN (ttlda_probe_gather_f64: the same
traversal over the same table and index stream, loads only) and against the HBM copy peak through the DRAM bytes of the
committed ncu capture.  `extra_configs` holds short measurements of the other BASELINE.json shapes: at N=1 cfg4 and one
1.25M-document shard of cfg5 (the HBM-class shapes); at N>1 cfg4 STRONG-scaled over the ranks (`Document.shard`, 408 MB
all-reduce) and, at N=8, cfg5 (8 x 1.25M documents, sampling=True, 800 MB all-reduce).  At N>1 the line also carries
`multigpu_parity`: shards fitted on the N GPUs against the single-process CPU oracle, checked before the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "doc_tokens_per_sec_per_em_iteration"
UNIT = "doc-tokens/s"
L2_BYTES = 126e6


# ----------------------------------------------------------------------------------------------------------------------
# byte models
# ----------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(nnz, D, K, V):
    """Gather-model bytes per launch of each kernel and per EM iteration: every nonzero touches one K-vector of fp64
    per dense operand (the ELBO token term is fused into the E-step => the 16K-per-nonzero variant)."""
    kb = 8 * K
    est = nnz * (kb + 8 + 12) + D * kb * 4 + (D + 1) * 8  # gather eB row + (colidx,count) + (csr2csc, ratio out); eT in x2, gamma in/out
    ref = D * kb * 2                                       # theta refresh pass: gamma in, eT out
    sst = nnz * (kb + 12) + V * kb + (V + 1) * 8         # gather eT row + (docidx, ratio in); S out
    mst = V * kb * 6                                       # S, eB in; lambda out; lambda in; eB out (+colsum pass)
    return {"ttlda_estep_f64": est, "ttlda_sstats_csc_f64": sst, "ttlda_mstep_f64": mst,
            "ttlda_theta_refresh_f64": ref, "iteration": est + sst + mst + ref}


def compulsory_bytes(nnz, D, K, V, nblocks):
    """Unique bytes each kernel cannot avoid moving across HBM when every table it gathers from is read once
    (perfect cache): the floor `roofline.traffic` is compared with."""
    kb = 8 * K
    est = nnz * (8 + 4 + 8) + D * kb * 3 + V * kb + (D + 1) * 8   # CSR + map in, ratio out; eT in, gamma in/out; eB once
    sst = nnz * 12 + D * kb + V * kb + (V + 1) * 8 * nblocks       # docidx + ratio in; eT once; S out
    if nblocks > 1:
        sst += 2 * nblocks * V * kb                                 # per-block partial statistics: written, folded
    return {"ttlda_estep_f64": est, "ttlda_sstats_csc_f64": sst, "ttlda_mstep_f64": V * kb * 5,
            "ttlda_theta_refresh_f64": D * kb * 2}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic(kernel, config):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json), else None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(config, {}).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------------------------
# CPU baselines: (i) the REAL reference (staged copy, oracle/ref_time.py, 1 core) and (ii) the oracle port
# (oracle/lda_oracle.c, gcc + OpenMP, all host threads), each on a bounded sample
# ----------------------------------------------------------------------------------------------------------------------
def cpu_port_iteration_time(rp, ci, ct, K, V, steps, warmup):
    """Seconds per full EM iteration (E-step + M-step + ELBO/perplexity) of the CPU oracle on the given CSR sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from lda_oracle import CSRCorpus, OracleLDA, c_num_threads, c_set_num_threads

    # all the host threads this process may use — torchrun exports OMP_NUM_THREADS=1, which must not throttle the
    # CPU baseline
    c_set_num_threads(host_threads())
    X = CSRCorpus(rp, ci, ct, V)
    m = OracleLDA(K, engine="c")
    m.M, m.V = X.D, V
    np.random.seed(42)
    m.lam = np.random.gamma(100., 0.01, (K, V))
    m.gamma = np.ascontiguousarray(np.full((X.D, K), 1.0 + V / K))
    _, logs = m.approx_perplexity(X)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        et, eb = m.em_step(X, logs[0], logs[1])
        _, logs = m.approx_perplexity(X, False, et, eb)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), c_num_threads(), X.tokens


def real_reference_time(rp, ci, ct, K, V, steps, warmup, full_fit=False, em_num_step=100, timeout=900):
    """The unmodified reference (oracle/_ref/tuotuo staged by oracle/build_ref.py, or /root/reference where it exists)
    run in a subprocess on the given CSR sample; returns oracle/ref_time.py's JSON record or {"unavailable": why}."""
    import numpy as np

    script = os.path.join(ROOT, "oracle", "ref_time.py")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "corpus.npz")
        np.savez(path, rowptr=rp, colidx=ci, counts=ct, V=V)
        cmd = [sys.executable, script, "--corpus", path, "--K", str(K), "--steps", str(steps), "--warmup", str(warmup)]
        if full_fit:
            cmd += ["--full-fit", "--em-num-step", str(em_num_step)]
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        except subprocess.TimeoutExpired:
            return {"unavailable": f"reference run exceeded {timeout} s"}
    if r.returncode != 0:
        return {"unavailable": (r.stderr.strip().splitlines() or ["reference run failed"])[-1][:200]}
    return json.loads(r.stdout.strip().splitlines()[-1])


def host_sample(config, docs, seed=1234):
    """First `docs` documents of the configuration's generator as host CSR (NumPy), via the oracle's counting."""
    import numpy as np
    from tuotuo_b200.synth import CONFIGS, synthetic_token_stream

    D, L, K, V, pois = CONFIGS[config]
    tok, offs = synthetic_token_stream(min(D, docs), L, K, V, device="cpu", seed=seed, poisson=pois)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from lda_oracle import CSRCorpus

    X = CSRCorpus.from_token_ids(tok.numpy(), offs.numpy(), V)
    return X.rowptr, X.colidx, X.counts, K, V


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.

    `value` = the REAL reference (unmodified tuotuo/lda_model.py + its compiled cutils, single-threaded: cores = 1) on a
    bounded document sample of the same workload, every step one pass of fit()'s loop body; when no copy of the
    reference is available, the oracle port (kind "port", all host threads).  The port's number on a 100k-document
    sample and the real reference at cfg1 (complete fit) and on a cfg2 prefix are reported beside it."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import numpy as np
    from tuotuo_b200.synth import CONFIGS, numpy_corpus

    D, L, K, V, pois = CONFIGS[args.config]
    # (ii) the port on >= 100k documents: the K*V M-step is then amortised as at full size
    rp, ci, ct, _, _ = host_sample(args.config, args.cpu_docs)
    n_port = len(rp) - 1
    sec, threads, tokens = cpu_port_iteration_time(rp, ci, ct, K, V, 2, 1)
    port = {"value": tokens / sec, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": sec * 1e3,
            "sample": f"first {n_port} of {D} documents of {args.config}, 2 timed EM iterations of oracle/lda_oracle.c "
                      f"(gcc -O2 -fopenmp)"}
    # (i) the real reference on a sample it can finish: ~7e3-2e4 doc-tokens/s, dense M x V matrix
    n_ref = min(n_port, args.ref_docs)
    e = int(rp[n_ref])
    rec = real_reference_time(rp[:n_ref + 1], ci[:e], ct[:e], K, V, args.steps, args.warmup)
    runs = {}
    if "unavailable" not in rec:
        val, ms, kind, cores = rec["tokens_per_s"], rec["seconds_per_iteration"] * 1e3, "reference", 1
        sample = (f"first {n_ref} of {D} documents of {args.config} (K={K}, V={V}), {args.steps} timed passes of the loop "
                  f"body of the unmodified reference's fit() (em_step + approx_perplexity), single-threaded")
        tokens_ref = rec["doc_tokens"]
        # cfg1 in full and a cfg2 prefix
        c1 = numpy_corpus(11, 100, 40, 5, 20)
        r1 = real_reference_time(*c1, 5, 40, 0, 0, full_fit=True)
        runs["cfg1_full_fit"] = r1
        if not args.quick_reference:
            rp2, ci2, ct2, K2, V2 = host_sample("cfg2", 2000)
            runs["cfg2_prefix_2000_docs"] = real_reference_time(rp2, ci2, ct2, K2, V2, 1, 0)
    else:
        val, ms, kind, cores, sample, tokens_ref = port["value"], port["ms_per_step"], "port", threads, port["sample"], tokens
        runs["real_reference"] = rec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config}: {sample}", "K": K, "V": V, "docs": n_ref if kind == "reference" else n_port,
                       "doc_tokens": tokens_ref},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "host_cpus": os.cpu_count()},
            "cpu_port": port, "reference_runs": runs,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def barrier(ctx):
    import torch

    if ctx.world > 1:
        ctx.dist.barrier()
    torch.cuda.synchronize()


def allmax(ctx, x):
    import torch

    t = torch.tensor([x], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return float(t.item())


def allsum_ints(ctx, xs):
    import torch

    t = torch.tensor(list(xs), dtype=torch.int64, device=ctx.dev)
    if ctx.world > 1:
        ctx.dist.all_reduce(t)
    return [int(v) for v in t.tolist()]


def probe_gather_gbs(nv, table, idx, reps=3):
    """Measured ceiling of the row gather table[idx[i], :] (loads only, product traversal) in GB/s of algorithmic bytes."""
    import torch

    sink = torch.zeros(1, dtype=torch.float64, device=table.device)
    n = idx.numel()
    if n == 0:
        return None
    nv.probe_gather(table, idx, sink)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        nv.probe_gather(table, idx, sink)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return n * table.shape[1] * 8 / (ms * 1e-3) / 1e9, ms


def measure(ctx, name, doc, K, V, steps, warmup, *, sampling=False, e2e=False, fit_check=False, scaling="weak",
            traffic_key=None):
    """Fit state for `doc` on this rank, W warm-up + S timed `em_iteration()` calls, per-kernel CUDA-event times,
    collective / host-wait attribution, the roofline of the dominant kernel and (optionally) the streamed e2e step."""
    import numpy as np
    import torch

    nv, dist, dev, world = ctx.nv, ctx.dist, ctx.dev, ctx.world
    lda = ctx.LDASmoothed(K, verbose=False, device=dev, sstats_mode=ctx.args.sstats)
    # em_num_step=0: corpus upload/CSR+CSC build, lambda_0 / gamma_0 initialisation, initial perplexity (untimed)
    p0 = lda.fit(doc, sampling=sampling, em_num_step=0, update_hyper_parms=False, return_perplexities=True)
    st = lda._st
    c = st.corpus
    tokens_all, nnz_all, docs_all = allsum_ints(ctx, (c.tokens, c.nnz, c.D))

    perps = [float(p0[0])]
    for _ in range(warmup):
        perps.append(lda.em_iteration(None, sampling=sampling)[0])

    # ---- timed region: value --------------------------------------------------------------------------------------------
    sampler = ClockSampler(ctx.local_rank)
    nv.timing = {}
    lda.host_wait_s = 0.0
    k0 = nv.kernel_count
    barrier(ctx)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        perps.append(lda.em_iteration(None, sampling=sampling)[0])
    e1.record()
    barrier(ctx)
    clocks = sampler.stop()
    launches = nv.kernel_count - k0
    ms_local = e0.elapsed_time(e1)
    per_kernel = {n: sum(a.elapsed_time(b) for a, b in ev) / len(ev) for n, ev in nv.timing.items()}
    nv.timing = None
    ms_per_step = allmax(ctx, ms_local) / steps
    res = {"value": tokens_all / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
           "scaling": scaling, "docs": docs_all, "doc_tokens": tokens_all, "nnz": nnz_all, "K": K, "V": V,
           "docs_per_gpu": c.D, "doc_blocks": c.nblocks, "sampling": sampling, "gpu_launches": launches,
           "clocks": clocks, "kernel_ms": {k: round(v, 4) for k, v in per_kernel.items()},
           "host_wait_ms": round(lda.host_wait_s / steps * 1e3, 4),
           "step_gap_ms": round(ms_local / steps - sum(v for k, v in per_kernel.items() if k.startswith("ttlda_")), 4),
           "perplexity_trace": [round(p, 6) for p in perps[:6]]}
    res["_launches"] = launches

    # ---- collectives, timed alone on this rank's buffers (N > 1) -------------------------------------------------------
    if world > 1:
        # every rank's own kernel time per step: the step runs at the pace of the slowest GPU (the ranks meet at the
        # statistics exchange), so the spread of this list is what weak scaling loses before any communication cost
        cap = 1.0 if (clocks and "sw_power_cap" in clocks.get("reasons", [])) else 0.0
        mine = torch.tensor([sum(v for k, v in per_kernel.items() if k.startswith("ttlda_")), ms_local / steps,
                             float(clocks["sm_mhz"]) if clocks else 0.0, cap, float(c.nnz)],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        res["per_rank"] = {"kernel_ms": [round(float(t[0]), 3) for t in allr],
                           "step_ms": [round(float(t[1]), 3) for t in allr],
                           "sm_mhz_median": [float(t[2]) for t in allr],
                           "sw_power_cap_seen": [bool(t[3]) for t in allr],
                           "nnz": [int(t[4]) for t in allr],
                           "note": "every rank samples its own GPU's clocks during the timed region"}
        work = st.sstats.clone()
        for _ in range(2):
            dist.all_reduce(work)
        barrier(ctx)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(work)
        a1.record()
        barrier(ctx)
        ar_ms = allmax(ctx, a0.elapsed_time(a1)) / 5
        nbytes = work.numel() * 8
        res["allreduce"] = {"bytes": nbytes, "ms_alone": round(ar_ms, 4),
                            "bus_GBps": round(2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9, 1),
                            "exchange_to_mstep_ms": per_kernel.get("exchange_to_mstep"),
                            "mstep_mode": st.mstep_mode, "peer_error": getattr(st, "peer_error", None),
                            "nvswitch_multicast": bool(getattr(st, "S_mc", 0)),
                            "note": "ms_alone = an NCCL all-reduce of a table of this size by itself, for reference; the "
                                    "iteration itself exchanges the statistics in mode `mstep_mode`: peer = NVLink loads / "
                                    "stores fused into the sharded M-step kernels (their times are in kernel_ms), nccl = "
                                    "reduce-scatter + all-gather around them, replicated = all-reduce + full M-step; "
                                    "exchange_to_mstep_ms = from the end of the statistics pass to the M-step's first "
                                    "kernel (theta refresh + waiting for the slowest rank)"}
        del work

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------------
    ab = algorithmic_bytes(c.nnz, c.D, K, V)
    ab["ttlda_estep_hot_f64"] = ab["ttlda_estep_f64"]
    ab["ttlda_sstats_csc_windowed_f64"] = ab["ttlda_sstats_csc_f64"] + c.nnz * 12 * max(0, -(-st.ld // max(c.kwindow, 1)) - 1)
    cb = compulsory_bytes(c.nnz, c.D, K, V, 1 if c.serial_blocks else c.nblocks)
    hbm, hbm_src = measured_peaks()
    hot = {k: v for k, v in per_kernel.items() if k in ab}
    dom = max(hot, key=hot.get)
    dom_ms = hot[dom]
    achieved = ab[dom] / (dom_ms * 1e-3) / 1e9
    ld = st.ld
    # which memory serves the dominant kernel's gather: the exp(E[log beta]) table (E-step) or one document block's
    # window of exp(E[log theta]) (statistics pass) against the 126 MB L2
    if dom == "ttlda_estep_f64":
        table_bytes, table, idx, what = V * ld * 8, st.eB, c.colidx, "exp(E[log beta]) [V][ld], index stream = CSR colidx"
    elif dom == "ttlda_sstats_csc_f64":
        table_bytes = min(c.D, c.block_docs if c.nblocks > 1 else c.D) * ld * 8
        table, idx, what = st.eT[st.cur], c.docidx, "exp(E[log theta]) [D][ld], index stream = mirror docidx"
    else:
        table_bytes, table, idx, what = 0, None, None, "streaming kernel"
    bound = "l2" if table is not None and table_bytes <= 0.6 * L2_BYTES else "hbm"
    roof = {"bound": bound, "kernel": dom, "achieved": achieved, "unit": "GB/s",
            "algorithmic_bytes_per_launch": ab[dom], "kernel_ms": dom_ms,
            "gathered_table_bytes": table_bytes, "hbm_peak": hbm, "hbm_peak_source": hbm_src,
            "compulsory_bytes": cb.get(dom)}
    probe = probe_gather_gbs(nv, table, idx[:c.nnz]) if table is not None else None
    if probe is not None:
        roof["peak"], roof["peak_source"] = probe[0], (
            f"measured in this run: ttlda_probe_gather_f64 — loads-only gather with the product traversal over the same "
            f"table and index stream ({what}); {probe[1]:.3f} ms per pass")
    else:
        roof["peak"], roof["peak_source"] = hbm, hbm_src
    roof["frac"] = achieved / roof["peak"]
    traffic = committed_traffic(dom, traffic_key) if traffic_key else None
    roof["traffic"] = traffic
    if traffic:
        roof["dram_frac"] = traffic / (dom_ms * 1e-3) / 1e9 / hbm
        roof["traffic_over_compulsory"] = traffic / cb[dom] if cb.get(dom) else None
    roof["note"] = ("achieved = gather-model bytes / kernel time; bound=l2: the gathered table (or document-block window) "
                    "fits the 126 MB L2, so the ceiling is the L2 -> SM path and `frac` is taken against the gather rate "
                    "measured above, never against the HBM copy peak; dram_frac = DRAM bytes of the committed ncu capture / "
                    "kernel time / hbm_peak")
    # every hot kernel, same arithmetic, for the record
    roof["all_kernels"] = {}
    for k, ms in hot.items():
        t = committed_traffic(k, traffic_key) if traffic_key else None
        roof["all_kernels"][k] = {"ms": round(ms, 4), "achieved_GBps": round(ab[k] / (ms * 1e-3) / 1e9, 1),
                                  "compulsory_GB": round(cb[k] / 1e9, 3) if k in cb else None,
                                  "traffic_GB": round(t / 1e9, 3) if t else None,
                                  "dram_frac": round(t / (ms * 1e-3) / 1e9 / hbm, 4) if t else None}
    if dom != "ttlda_estep_f64" and "ttlda_estep_f64" in hot:
        pe = probe_gather_gbs(nv, st.eB, c.colidx[:c.nnz])
        roof["all_kernels"]["ttlda_estep_f64"]["probe_GBps"] = round(pe[0], 1)
        roof["all_kernels"]["ttlda_estep_f64"]["frac_of_probe"] = round(roof["all_kernels"]["ttlda_estep_f64"]["achieved_GBps"] / pe[0], 4)
    if dom != "ttlda_sstats_csc_f64" and "ttlda_sstats_csc_f64" in hot and c.docidx is not None:
        ps = probe_gather_gbs(nv, st.eT[st.cur], c.docidx[:c.nnz])
        roof["all_kernels"]["ttlda_sstats_csc_f64"]["probe_GBps"] = round(ps[0], 1)
        roof["all_kernels"]["ttlda_sstats_csc_f64"]["frac_of_probe"] = round(roof["all_kernels"]["ttlda_sstats_csc_f64"]["achieved_GBps"] / ps[0], 4)
    roof["iteration_achieved_GBps"] = ab["iteration"] / (ms_per_step * 1e-3) / 1e9
    res["roofline"] = roof

    # ---- the same iterations through fit() itself: the headline cannot drift from what fit does ------------------------
    if fit_check:
        def timed_fit(n):
            barrier(ctx)
            t0 = time.perf_counter()
            out = lda.fit(doc, sampling=sampling, threshold=-float("inf"), em_num_step=n, update_hyper_parms=False,
                          return_perplexities=True)
            torch.cuda.synchronize()
            return out, (time.perf_counter() - t0) * 1e3

        n_fit = max(2, steps)
        _, ms0 = timed_fit(0)
        pf, msn = timed_fit(n_fit)
        res["fit_call"] = {"em_num_step": n_fit, "wall_ms_total": round(msn, 3), "wall_ms_em_num_step_0": round(ms0, 3),
                           "marginal_wall_ms_per_iteration": round((msn - ms0) / n_fit, 3),
                           "note": "LDASmoothed.fit(em_num_step=n, threshold=-inf, update_hyper_parms=False), host wall "
                                   "clock: state allocation + lambda_0 (host RNG) + n x em_iteration() + final perplexity "
                                   "pass; marginal = (fit(n) - fit(0)) / n, to be compared with ms_per_step",
                           "perplexities_equal_em_iteration": bool(np.allclose(pf[:min(n_fit, len(perps) - 1)],
                                                                               perps[1:1 + min(n_fit, len(perps) - 1)],
                                                                               rtol=1e-12))}
        st = lda._st
        c = st.corpus

    # ---- e2e: CSR corpus arrives from pinned host memory every step ---------------------------------------------------
    if e2e:
        # Host format: int64 rowptr + the narrowest unsigned types that hold the data losslessly (uint16 word ids when
        # V <= 65536, uint8 / uint16 counts), which em_step_from_host widens on the device; the plain int32 form is timed
        # next to it (e2e.int32).
        h_rp = c.rowptr.cpu().pin_memory()
        host_forms = {"int32": (c.colidx.cpu().pin_memory(), c.counts.cpu().pin_memory())}
        cmax = int(c.counts.max().item()) if c.nnz else 0
        ci_c = c.colidx.to(torch.int16) if V <= 65536 else c.colidx  # two's-complement truncation keeps the uint16 bits
        ct_c = c.counts.to(torch.uint8) if cmax <= 255 else (c.counts.to(torch.int16) if cmax <= 65535 else c.counts)
        compact = ci_c.dtype != torch.int32 or ct_c.dtype != torch.int32
        if compact:
            host_forms["compact"] = (ci_c.cpu().pin_memory(), ct_c.cpu().pin_memory())
        del ci_c, ct_c
        n_e2e = max(2, min(steps, 5))
        e2e_res = {}
        for form, (h_ci, h_ct) in host_forms.items():
            for _ in range(max(1, min(warmup, 2))):
                lda.em_step_from_host(h_rp, h_ci, h_ct, e_num_step=100)
            barrier(ctx)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(n_e2e):
                # public streaming API: per document block, pinned-host -> device upload overlapped with the previous
                # block's CSC mirror + E-step; then statistics, [all-reduce], M-step, perplexity terms -> host
                lda.em_step_from_host(h_rp, h_ci, h_ct, e_num_step=100)
            f1.record()
            barrier(ctx)
            ms_form = allmax(ctx, f0.elapsed_time(f1)) / n_e2e
            e2e_res[form] = {"ms_per_step": ms_form, "value": tokens_all / (ms_form * 1e-3),
                             "h2d_bytes_per_step": (h_rp.numel() * 8 + h_ci.numel() * h_ci.element_size()
                                                    + h_ct.numel() * h_ct.element_size()) * world,
                             "host_dtypes": f"rowptr int64, colidx {'uint16' if h_ci.element_size() == 2 else 'int32'}, "
                                            f"counts {({1: 'uint8', 2: 'uint16', 4: 'int32'})[h_ct.element_size()]}"}
        main = e2e_res["compact" if compact else "int32"]
        res["e2e"] = {"value": main["value"], "unit": UNIT, "h2d_bytes_per_step": main["h2d_bytes_per_step"],
                      "d2h_bytes_per_step": 56 * world, "ms_per_step": main["ms_per_step"], "steps": n_e2e,
                      "host_dtypes": main["host_dtypes"], "int32": e2e_res["int32"],
                      "what": "LDASmoothed.em_step_from_host: pinned-host CSR -> device one document block at a time; "
                              "per block CSC mirror + E-step overlap the next block's upload; then statistics, "
                              "theta refresh, M-step, perplexity terms -> host"}
    res["_corpus_bytes"] = c.nbytes() + 3 * c.D * K * 8
    res["_rank0_sample"] = (lda, c) if ctx.rank == 0 else None
    return res


def multigpu_parity(ctx):
    """Shards fitted on the N GPUs (NCCL all-reduce of the statistics) against the single-process CPU oracle on the
    whole corpus, K in {20, 100, 500}; also: lambda bitwise identical on every rank.  The oracle is the CHECKER here."""
    import numpy as np
    import torch

    from tuotuo_b200.document import Document
    from tuotuo_b200.synth import numpy_corpus

    def relerr(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

    out = {}
    for K, V, D, L, steps in [(20, 800, 600, 80, 4), (100, 1500, 400, 100, 3), (500, 1200, 96, 150, 2)]:
        rp, ci, ct = numpy_corpus(500 + K, D, V, min(K, 30), L, ragged=True)
        m = ctx.LDASmoothed(K, verbose=False, device=ctx.dev)
        perps = m.fit(Document.from_csr(rp, ci, ct, V).shard(ctx.rank, ctx.world), em_num_step=steps,
                      return_perplexities=True)
        gam, lam = m._gamma_, m._lambda_
        h = torch.tensor([float(np.sum(lam)), float(perps[-1])], dtype=torch.float64, device=ctx.dev)
        hs = [torch.zeros_like(h) for _ in range(ctx.world)]
        ctx.dist.all_gather(hs, h)
        same = all(torch.equal(hs[0], x) for x in hs)
        if ctx.rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            from lda_oracle import CSRCorpus, OracleLDA

            o = OracleLDA(K, engine="c")
            po = o.fit(CSRCorpus(rp, ci, ct, V), em_num_step=steps)
            e = dict(perplexity=relerr(perps, po), gamma=relerr(gam, o.gamma), lam=relerr(lam, o.lam),
                     alpha=relerr(m._alpha_, o.alpha), eta=relerr(m._eta_, o.eta), ranks_bitwise_equal=bool(same))
            e["ok"] = bool(e["perplexity"] < 1e-8 and max(e["gamma"], e["lam"], e["alpha"], e["eta"]) < 1e-9 and same
                           and gam.shape == (D, K))
            out[f"K{K}"] = e
    if ctx.rank == 0:
        out["ok"] = all(v["ok"] for v in out.values())
        out["what"] = (f"{ctx.world} ranks, Document.shard + NCCL all-reduce vs oracle/lda_oracle.c on the whole corpus; "
                       "max relative errors (gates: 1e-9 gamma/lambda/alpha/eta, 1e-8 perplexity)")
    return out


def strip(res):
    return {k: v for k, v in res.items() if not k.startswith("_")}


def run_ours(args):
    import torch
    import torch.distributed as dist

    ctx = Ctx()
    ctx.args = args
    ctx.rank, ctx.local_rank, ctx.world = dist_env()
    torch.cuda.set_device(ctx.local_rank)
    ctx.dev = torch.device("cuda", ctx.local_rank)
    ctx.dist = dist
    if ctx.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=ctx.dev)

    from tuotuo_b200 import _native as nv
    from tuotuo_b200.lda_model import LDASmoothed
    from tuotuo_b200.synth import CONFIGS, synthetic_document

    ctx.nv, ctx.LDASmoothed = nv, LDASmoothed
    rank, world, dev = ctx.rank, ctx.world, ctx.dev

    parity = multigpu_parity(ctx) if world > 1 and not args.no_parity else None

    # ---- headline: cfg3, one shard per GPU (weak scaling) ---------------------------------------------------------------
    D, L, K, V, pois = CONFIGS[args.config]
    if args.docs:
        D = args.docs
    if args.strong and world > 1:  # the SAME corpus on every rank, cut into contiguous nnz-balanced document shards
        doc = synthetic_document(args.config, device=dev, seed=1234, docs=D).shard(rank, world)
    else:
        # weak scaling: every rank holds its own documents of ONE corpus (same topics, rank 0's shard is the N = 1 corpus)
        doc = synthetic_document(args.config, device=dev, seed=1234, docs=D, doc_seed=(1234 + rank) if rank else None)
    main = measure(ctx, args.config, doc, K, V, args.steps, args.warmup, e2e=not args.no_e2e, fit_check=args.fit_check,
                   scaling="strong" if args.strong else "weak", traffic_key=args.config if not args.docs else None)
    main.setdefault("e2e", None)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:  # reported at N=1 only (bench contract)
        lda, c = main["_rank0_sample"]
        n = min(c.D, args.cpu_docs)
        rp = c.rowptr[:n + 1].cpu().numpy()
        e = int(rp[-1])
        ci, ct = c.colidx[:e].cpu().numpy(), c.counts[:e].cpu().numpy()
        sec, threads, tok = cpu_port_iteration_time(rp, ci, ct, K, V, 2, 1)
        cpu = {"value": tok / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n} of {c.D} documents of rank 0's {args.config} shard, 2 timed EM iterations of "
                         f"oracle/lda_oracle.c (gcc -O2 -fopenmp), {os.cpu_count()} host CPUs visible",
               "ms_per_step": sec * 1e3}
        nr = min(n, args.ref_docs)
        er = int(rp[nr])
        rec = real_reference_time(rp[:nr + 1], ci[:er], ct[:er], K, V, 1, 0, timeout=300)
        cpu["real_reference"] = ({"value": rec["tokens_per_s"], "unit": UNIT, "cores": 1, "kind": "reference",
                                  "ms_per_step": rec["seconds_per_iteration"] * 1e3,
                                  "sample": f"first {nr} documents of the same shard, 1 pass of the loop body of the "
                                            f"unmodified reference's fit() (oracle/ref_time.py)"}
                                 if "unavailable" not in rec else rec)
    main.pop("_rank0_sample", None)
    corpus_gb = main["_corpus_bytes"] / 1e9
    del doc
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json shapes (short runs) ----------------------------------------------------------------------
    extra = {}
    if not args.no_extra and not args.docs and args.config == "cfg3":
        xs, xw = args.extra_steps, 3
        if world == 1:
            plan = [("cfg4", dict(config="cfg4", docs=None, sampling=False, shard=False, scaling="weak", traffic="cfg4")),
                    ("cfg5_shard_1p25M", dict(config="cfg5", docs=1_250_000, sampling=True, shard=False, scaling="weak",
                                              traffic="cfg5_shard"))]
        else:
            plan = [("cfg4_strong", dict(config="cfg4", docs=None, sampling=False, shard=True, scaling="strong", traffic=None))]
            if world == 8:
                plan.append(("cfg5_x8", dict(config="cfg5", docs=1_250_000, sampling=True, shard=False, scaling="weak",
                                             traffic=None)))
        for name, p in plan:
            Dx, Lx, Kx, Vx, _ = CONFIGS[p["config"]]
            if p["shard"]:  # the SAME corpus on every rank, cut into contiguous nnz-balanced document shards
                dx = synthetic_document(p["config"], device=dev, seed=1234).shard(rank, world)
            else:
                dx = synthetic_document(p["config"], device=dev, seed=1234, docs=p["docs"],
                                        doc_seed=(1234 + rank) if rank else None)
                if p["sampling"]:
                    dx.num_doc_population = Dx  # the population the sampled ELBO is scaled to (lda_model.py:152-153)
            r = measure(ctx, name, dx, Kx, Vx, xs, xw, sampling=p["sampling"], scaling=p["scaling"],
                        traffic_key=p["traffic"])
            r.pop("_rank0_sample", None)
            extra[name] = strip(r)
            del dx, r
            torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.config}: {main['docs']} docs x ~{L} tokens, K={K}, V={V} "
                                       f"({main['docs'] // world} docs per GPU), batch VI EM iteration",
                           "docs": main["docs"], "doc_tokens": main["doc_tokens"], "nnz": main["nnz"], "K": K, "V": V,
                           "parallelism": f"doc-shard x{world}" + (
                               f" + K*V statistics exchange ({os.environ.get('TTLDA_MSTEP', 'peer')}: fused into the "
                               "vocabulary-sharded M-step over NVLink peer memory / NVSwitch multicast; nccl: "
                               "reduce-scatter + all-gather; replicated: all-reduce)" if world > 1 else ""),
                           "sstats": args.sstats, "doc_blocks": main["doc_blocks"],
                           "api": "LDASmoothed.em_iteration() — the loop body fit() calls",
                           "elbo": "token term fused into the E-step (16K-byte-per-nonzero variant, SURVEY 8d)",
                           "l2": "no explicit flush: each step streams the CSR/CSC arrays and the D*K theta tables "
                                 f"({corpus_gb:.1f} GB per GPU) >> 126 MB L2"},
                "clocks": main["clocks"], "gpu_launches": main["gpu_launches"],
                "e2e": main["e2e"], "roofline": main["roofline"], "cpu_baseline": cpu,
                "kernel_ms": main["kernel_ms"], "host_wait_ms": main["host_wait_ms"], "step_gap_ms": main["step_gap_ms"],
                "fit_call": main.get("fit_call"), "perplexity_trace": main["perplexity_trace"]}
        if "allreduce" in main:
            line["allreduce"] = main["allreduce"]
            line["per_rank"] = main.get("per_rank")
        if parity is not None:
            line["multigpu_parity"] = parity
        if extra:
            line["extra_configs"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--docs", type=int, default=0, help="override documents per GPU (debug)")
    ap.add_argument("--sstats", default="csc", choices=["csc", "atomic"])
    ap.add_argument("--cpu-docs", type=int, default=100000, help="documents in the bounded sample of the CPU port")
    ap.add_argument("--ref-docs", type=int, default=100, help="documents in the bounded sample of the real reference")
    ap.add_argument("--extra-steps", type=int, default=3, help="timed steps of each extra configuration")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg4 / cfg5 measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the N>1 shard-vs-oracle check")
    ap.add_argument("--strong", action="store_true", help="N>1: shard ONE corpus of the configuration over the ranks")
    ap.add_argument("--no-e2e", action="store_true", help="skip the streamed (host -> device) measurement")
    ap.add_argument("--no-fit-check", dest="fit_check", action="store_false")
    ap.add_argument("--quick-reference", action="store_true", help="--impl reference: skip the cfg2 prefix run")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
