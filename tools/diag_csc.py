"""One ttlda_csr_to_csc_block call at cfg3 block size (run under ncu for the per-kernel split)."""
import sys

import torch

sys.path.insert(0, '/root/repo')
from tuotuo_b200 import _native as nv
from tuotuo_b200.synth import synthetic_document

dev = torch.device('cuda', 0)
doc = synthetic_document('cfg3', device=dev, docs=250000)
c = doc.device_corpus(dev)
V = c.V
bd = 125000
d0, d1 = 0, bd
rp = c.rowptr.cpu()
s, e = int(rp[d0]), int(rp[d1])
colptr = torch.empty(V + 1, dtype=torch.int64, device=dev)
docidx = torch.empty(c.nnz, dtype=torch.int32, device=dev)
ccounts = torch.empty(c.nnz, dtype=torch.int32, device=dev)
csr2csc = torch.empty(c.nnz, dtype=torch.int32, device=dev)
ws = torch.empty(nv.csr_to_csc_workspace_bytes(e - s, bd, V, 1), dtype=torch.uint8, device=dev)
for _ in range(2):
    nv.csr_to_csc_block(c.rowptr[d0:d1 + 1], c.colidx, c.counts, V, d0, s, e - s, colptr, docidx, ccounts, ws, csr2csc)
torch.cuda.synchronize()
print("nnz block", e - s)
