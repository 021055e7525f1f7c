import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np, tol_b200 as T, time
from tol_b200.evaluator import padded_ld
g=np.load("tests/golden/S10_tempest_ts200.npz"); ev=T.Evaluator.from_golden(g)
B=16384
X=torch.zeros(B,padded_ld(ev.n),dtype=torch.float64); T.synth.batch(g["x"][0],1,0,B,out=X.numpy()); X=X.cuda()
F=torch.empty(B,padded_ld(ev.neF),dtype=torch.float64,device="cuda")
Gc=torch.empty(B,padded_ld(ev.compact_len),dtype=torch.float64,device="cuda")
G=torch.empty(B,padded_ld(ev.neG),dtype=torch.float64,device="cuda")
st=torch.cuda.Stream(); torch.cuda.set_stream(st); ev.set_stream(st.cuda_stream)
ev.eval_batch_device(X,F,Gc,compact_rows=True)
G2=torch.empty(B,padded_ld(ev.neG),dtype=torch.float64,device="cuda")
ev.repack_csc_device(G,G2)
for fn,name,by in ((lambda: ev.expand_compact_device(Gc,G,sync=False),"device expansion",8*B*(ev.compact_len+ev.neG)),(lambda: ev.repack_csc_device(G,G2,sync=False),"CSC repack",8*B*2*ev.neG)):
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): fn()
    e1.record(st); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(name, "%.3f ms"%ms, "%.0f GB/s"%(by/ms/1e6))
