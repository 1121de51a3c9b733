"""Where does one em_step_from_host go?  Per-C-call CUDA-event times of one streamed EM step at cfg3."""
import sys
import time

import torch

sys.path.insert(0, '/root/repo')
from tuotuo_b200 import _native as nv
from tuotuo_b200.lda_model import LDASmoothed
from tuotuo_b200.synth import synthetic_document

dev = torch.device('cuda', 0)
doc = synthetic_document('cfg3', device=dev)
lda = LDASmoothed(100, verbose=False, device=dev)
lda.fit(doc, em_num_step=0, update_hyper_parms=False)
c = lda._st.corpus
h = [x.cpu().pin_memory() for x in (c.rowptr, c.colidx, c.counts)]


def t(f, n=3):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def h2d():
    c.rowptr.copy_(h[0], non_blocking=True)
    c.colidx.copy_(h[1], non_blocking=True)
    c.counts.copy_(h[2], non_blocking=True)


b = sum(x.numel() * x.element_size() for x in h)
ms = t(h2d)
print("H2D", ms, "ms", b / ms / 1e6, "GB/s")
print("csc build (one call)", t(lambda: c.build_csc(c.nblocks, c.block_docs)))
print("em_step_from_host", t(lambda: lda.em_step_from_host(*h)))
nv.timing = {}
torch.cuda.synchronize()
t0 = time.perf_counter()
lda.em_step_from_host(*h)
torch.cuda.synchronize()
print("wall of the timed-call step (ms)", (time.perf_counter() - t0) * 1e3)
for k, v in nv.timing.items():
    ts = [a.elapsed_time(b_) for a, b_ in v]
    print(f"{k:32s} n={len(ts):2d} sum={sum(ts):7.3f} ms  each={[round(x, 2) for x in ts]}")
nv.timing = None
