import torch, time, sys
sys.path.insert(0,'/root/repo')
from tuotuo_b200.synth import synthetic_document
from tuotuo_b200.lda_model import LDASmoothed
dev=torch.device('cuda',0)
doc=synthetic_document('cfg3',device=dev)
lda=LDASmoothed(100,verbose=False,device=dev)
lda.fit(doc,em_num_step=0,update_hyper_parms=False)
c=lda._st.corpus
h=[x.cpu().pin_memory() for x in (c.rowptr,c.colidx,c.counts)]
def t(f,n=3):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
def h2d():
    c.rowptr.copy_(h[0],non_blocking=True); c.colidx.copy_(h[1],non_blocking=True); c.counts.copy_(h[2],non_blocking=True)
b=sum(x.numel()*x.element_size() for x in h)
ms=t(h2d); print("H2D", ms, "ms", b/ms/1e6, "GB/s")
print("csc build", t(lambda: c.build_csc(c.nblocks,c.block_docs)))
print("em_step_from_host chunks=8", t(lambda: lda.em_step_from_host(*h,chunks=8)))
print("em_step_from_host chunks=16", t(lambda: lda.em_step_from_host(*h,chunks=16)))
print("em_step_from_host chunks=2", t(lambda: lda.em_step_from_host(*h,chunks=2)))
