#!/usr/bin/env python
"""bench.py — doc-tokens/sec per full variational-inference EM iteration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is ONE full EM iteration of the hot path over the resident corpus: E-step (gamma + theta refresh, with the
ELBO token term of the incoming state as a by-product of the same dot products), the device->host read of the
perplexity terms and the host-side convergence test, K x V sufficient statistics, [NCCL all-reduce of the
statistics when N > 1], M-step (lambda + beta refresh) — exactly the body of LDASmoothed.fit's loop; every step
yields one perplexity.  Workload at N=1: BASELINE.json configs[2] "synthetic 1M docs x 300 tokens, K=100, V=50k" (cfg3), the
configuration the 60 %-of-roofline target is quoted on; for N > 1 every rank holds its own cfg3-sized document
shard (weak scaling), lambda is replicated and the statistics are summed with one fp64 all-reduce per iteration.

Prints ONE JSON line (rank 0).  `value` is the device-timed whole-job throughput with the corpus resident in HBM;
`e2e` repeats the measurement through the public API with the CSR corpus coming from pinned HOST memory every
step (H2D copy + CSC rebuild + EM iteration + D2H of the result inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "doc_tokens_per_sec_per_em_iteration"
UNIT = "doc-tokens/s"


# ----------------------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(nnz, D, K, V):
    """Gather-model bytes.  Per launch of each kernel and per EM iteration."""
    kb = 8 * K
    est = nnz * (kb + 8 + 12) + D * kb * 4 + (D + 1) * 8  # gather eB row + (colidx,count) + (csr2csc, ratio out); eT in x2, gamma in/out
    ref = D * kb * 2                                       # theta refresh pass: gamma in, eT out
    sst = nnz * (kb + 12) + V * kb + (V + 1) * 8         # gather eT row + (docidx, ratio in); S out
    mst = V * kb * 6                                       # S, eB in; lambda out; lambda in; eB out (+colsum pass)
    # the ELBO token term is a by-product of the E-step (same dot product): no third gather => 16K bytes per nonzero
    return {"ttlda_estep_f64": est, "ttlda_sstats_csc_f64": sst, "ttlda_mstep_f64": mst,
            "ttlda_theta_refresh_f64": ref, "iteration": est + sst + mst + ref}


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
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
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


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (oracle/lda_oracle.c, gcc + OpenMP, all host threads) on a bounded sample
# ----------------------------------------------------------------------------------------------------------------------
def cpu_port_iteration_time(rp, ci, ct, K, V, steps, warmup):
    """Seconds per full EM iteration (E-step + M-step + ELBO/perplexity) of the CPU oracle on the given CSR sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from lda_oracle import CSRCorpus, OracleLDA, c_num_threads, c_set_num_threads

    # all the host threads this process may use — torchrun exports OMP_NUM_THREADS=1, which must not throttle the
    # CPU baseline
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    c_set_num_threads(ncpu)
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


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the Python reference itself
    cannot travel to the GPU box and manages ~7e3 doc-tokens/s)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, ROOT)
    from tuotuo_b200.synth import CONFIGS, synthetic_token_stream

    D, L, K, V, pois = CONFIGS[args.config]
    n = min(D, args.cpu_docs)
    tok, offs = synthetic_token_stream(n, L, K, V, device="cpu", seed=1234, poisson=pois)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from lda_oracle import CSRCorpus

    X = CSRCorpus.from_token_ids(tok.numpy(), offs.numpy(), V)
    sec, threads, tokens = cpu_port_iteration_time(X.rowptr, X.colidx, X.counts, K, V, args.steps, args.warmup)
    val = tokens / sec
    sample = f"{n} of {D} documents of {args.config} (same generator, K={K}, V={V}), full EM iteration incl. the K*V M-step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config}: {sample}", "K": K, "V": V, "docs": n, "doc_tokens": tokens},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from tuotuo_b200 import _native as nv
    from tuotuo_b200.lda_model import LDASmoothed
    from tuotuo_b200.synth import CONFIGS, synthetic_document

    D, L, K, V, pois = CONFIGS[args.config]
    if args.docs:
        D = args.docs
    doc = synthetic_document(args.config, device=dev, seed=1234 + rank, docs=D)
    lda = LDASmoothed(K, verbose=False, device=dev, sstats_mode=args.sstats)
    # em_num_step=0: corpus upload/CSR+CSC build, lambda_0 / gamma_0 initialisation, initial perplexity (untimed)
    p0 = lda.fit(doc, em_num_step=0, update_hyper_parms=False, return_perplexities=True)
    st = lda._st
    c = st.corpus
    tokens_local, nnz_local = c.tokens, c.nnz

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        # exactly the body of LDASmoothed.fit's loop: E-step (+ token term of the incoming state), 48-byte D2H and
        # perplexity / convergence test on the host, statistics, [all-reduce], M-step
        lda._estep_speculative(e_num_step=100)
        lda._collect(tok_row=0)
        p = lda._perplexity_from_terms(False)
        lda._commit_iteration()
        return p

    perps = [float(p0[0])]
    for _ in range(args.warmup):
        perps.append(one_step())

    # ---- timed region: value --------------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    nv.timing = {}
    k0 = nv.kernel_count
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        perps.append(one_step())
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = nv.kernel_count - k0
    ms = e0.elapsed_time(e1)
    per_kernel = {n: sum(a.elapsed_time(b) for a, b in ev) / len(ev) for n, ev in nv.timing.items()}
    nv.timing = None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([tokens_local, nnz_local, c.D], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms = float(t.item())
    tokens_all, nnz_all, docs_all = (int(x) for x in tot.tolist())
    ms_per_step = ms / args.steps
    value = tokens_all / (ms_per_step * 1e-3)

    # ---- e2e: CSR corpus arrives from pinned host memory every step ---------------------------------------------------
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
    n_e2e = max(2, min(args.steps, 5))
    e2e_res = {}
    for form, (h_ci, h_ct) in host_forms.items():
        def e2e_step():
            # public streaming API: per document block, pinned-host -> device upload overlapped with the previous
            # block's CSC mirror + E-step + statistics; then fold, [all-reduce], M-step, perplexity terms -> host
            return lda.em_step_from_host(h_rp, h_ci, h_ct, e_num_step=100)

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(n_e2e):
            e2e_step()
        f1.record()
        barrier()
        t2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_form = float(t2.item()) / n_e2e
        e2e_res[form] = {"ms_per_step": ms_form, "value": tokens_all / (ms_form * 1e-3),
                         "h2d_bytes_per_step": (h_rp.numel() * 8 + h_ci.numel() * h_ci.element_size()
                                                + h_ct.numel() * h_ct.element_size()) * world,
                         "host_dtypes": f"rowptr int64, colidx {'uint16' if h_ci.element_size() == 2 else 'int32'}, "
                                        f"counts {({1: 'uint8', 2: 'uint16', 4: 'int32'})[h_ct.element_size()]}"}
    e2e_main = e2e_res["compact" if compact else "int32"]
    e2e_ms, e2e_val, h2d_all = e2e_main["ms_per_step"], e2e_main["value"], e2e_main["h2d_bytes_per_step"]

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------------------
    ab = algorithmic_bytes(nnz_local, c.D, K, V)
    peak, peak_src = measured_peaks()
    hot = {k: v for k, v in per_kernel.items() if k in ab}
    dom = max(hot, key=hot.get)
    dom_ms = hot[dom]
    achieved = ab[dom] / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": committed_traffic(dom, args.config), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab[dom], "kernel_ms": dom_ms,
                "kernel_ms_all": {k: round(v, 4) for k, v in per_kernel.items()},
                "iteration_frac": ab["iteration"] / (ms_per_step * 1e-3) / 1e9 / peak,
                "note": "gathers of the 8*K*V-byte exp(E[log beta]) table are L2 hits when it fits the 126 MB L2"}
    if 8 * K * V <= 100e6:
        # the table the dominant kernel gathers from is L2-resident: the relevant ceiling is the L2 -> SM path,
        # 64 B/clk/SM (tools/membench2.cu level 0 / membench7.cu: 16.5-18.5 TB/s measured for this traversal)
        l2_peak = 64.0 * 148 * 1.92
        roofline["l2_gather"] = {"peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                                 "source": "64 B/clk/SM x 148 SMs x 1.92 GHz; tools/membench2.cu level 0 measures "
                                           "16.5-18.5 TB/s for the same traversal (DESIGN.md section 4)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:  # reported at N=1 only (bench contract)
        n = min(c.D, args.cpu_docs)
        rp = c.rowptr[:n + 1].cpu().numpy()
        e = int(rp[-1])
        sec, threads, tok = cpu_port_iteration_time(rp, c.colidx[:e].cpu().numpy(), c.counts[:e].cpu().numpy(), K, V,
                                                    2, 1)
        cpu = {"value": tok / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n} of {c.D} documents of rank 0's {args.config} shard, 2 timed EM iterations of "
                         f"oracle/lda_oracle.c (gcc -O2 -fopenmp), {os.cpu_count()} host CPUs visible",
               "ms_per_step": sec * 1e3}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.config}: {docs_all} docs x ~{L} tokens, K={K}, V={V} "
                                       f"({docs_all // world} docs per GPU), batch VI EM iteration",
                           "docs": docs_all, "doc_tokens": tokens_all, "nnz": nnz_all, "K": K, "V": V,
                           "parallelism": f"doc-shard x{world}" + (" + NCCL all-reduce(K*V f64)" if world > 1 else ""),
                           "sstats": args.sstats, "doc_blocks": c.nblocks,
                           "elbo": "token term fused into the E-step (16K-byte-per-nonzero variant, SURVEY 8d)",
                           "l2": "no explicit flush: each step streams the CSR/CSC arrays and the D*K theta tables "
                                 f"({(c.nbytes() + 3 * c.D * K * 8) / 1e9:.1f} GB per GPU) >> 126 MB L2"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": 48 * world,
                        "ms_per_step": e2e_ms, "steps": n_e2e, "host_dtypes": e2e_main["host_dtypes"],
                        "int32": e2e_res["int32"],
                        "what": "LDASmoothed.em_step_from_host: pinned-host CSR -> device one document block at a time; "
                                "per block CSC mirror + E-step overlap the next block's upload; then statistics, "
                                "theta refresh, M-step, perplexity terms -> host"},
                "roofline": roofline, "cpu_baseline": cpu,
                "perplexity_trace": [round(p, 6) for p in perps[:8]]}
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
    ap.add_argument("--cpu-docs", type=int, default=20000, help="documents in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
