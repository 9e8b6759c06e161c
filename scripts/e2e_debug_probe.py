import sys, time; sys.path.insert(0,"."); sys.path.insert(0,"vlm-compression_b200")
import torch, bench
from vlmc import native, schedule
from vlmc.compression.pruners import sparsegpt_pruner as sp
dev=torch.device("cuda",0)
orig=schedule.sparsegpt_block
def timed_block(items, *a, **k):
    torch.cuda.synchronize(); t0=time.perf_counter()
    r=orig(items,*a,**k)
    torch.cuda.synchronize(); print("  sparsegpt_block %.1f ms, distinct H %d" % ((time.perf_counter()-t0)*1e3, len({id(i[1]) for i in items})), flush=True)
    return r
schedule.sparsegpt_block=timed_block
orig_free=sp.SparseGPT.free
def timed_free(self):
    t0=time.perf_counter(); orig_free(self); torch.cuda.synchronize(); print("  free %.1f ms" % ((time.perf_counter()-t0)*1e3), flush=True)
sp.SparseGPT.free=timed_free
inputs=bench.make_inputs(torch, dev, 128, 1000)
class A: pass
args=A(); args.method="sparsegpt"; args.steps=2; args.calib_batch=128
from vlmc import parallel
t0=time.perf_counter()
r=bench.run_e2e(torch, native, parallel, dev, args, 0, 1, inputs)
print(r, time.perf_counter()-t0)
