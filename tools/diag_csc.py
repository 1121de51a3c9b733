"""ttlda_csr_to_csc_block at the cfg3 block size (62 500 documents, 12.7M entries), as the streamed step calls it (no mirrored
counts): CUDA-event time per call, slice walk against the CUB pair sort; run under ncu for the per-kernel split."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tuotuo_b200 import _native as nv
from tuotuo_b200.synth import synthetic_document

dev = torch.device('cuda', 0)
doc = synthetic_document('cfg3', device=dev, docs=125000)
c = doc.device_corpus(dev)
V = c.V
bd = 62500
rp = c.rowptr.cpu()
colptr = torch.empty(2 * V + 1, dtype=torch.int64, device=dev)
docidx = torch.empty(c.nnz, dtype=torch.int32, device=dev)
csr2csc = torch.empty(c.nnz, dtype=torch.int32, device=dev)
big = max(int(rp[bd]), int(rp[2 * bd]) - int(rp[bd]))
ws = torch.empty(nv.csr_to_csc_workspace_bytes(big, bd, V, 1), dtype=torch.uint8, device=dev)


def run():
    for b in range(2):
        d0, d1 = b * bd, (b + 1) * bd
        s, e = int(rp[d0]), int(rp[d1])
        nv.csr_to_csc_block(c.rowptr[d0:d1 + 1], c.colidx, c.counts, V, d0, s, e - s, colptr[b * V:b * V + V + 1], docidx,
                            None, ws, csr2csc)


reps = int(os.environ.get("DIAG_REPS", "5"))
run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
print("TTLDA_CSC_WALK =", os.environ.get("TTLDA_CSC_WALK", "(default: walk)"), " block entries", int(rp[bd]),
      " ms per block:", e0.elapsed_time(e1) / reps / 2)
