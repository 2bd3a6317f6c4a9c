"""cfg3 hp apply timing (32^3 elements, p in 1..6), CUDA events on the context stream."""
import sys, numpy as np
sys.path.insert(0,'/root/repo/dune-hpdg_b200'); sys.path.insert(0,'/root/repo')
import torch, hpdg_b200 as hp
rng = np.random.default_rng(1887)
deg = rng.integers(1, 7, 32**3).astype(np.int32)
for variant in (0,):
    ctx = hp.Context((32,32,32), degree=deg)
    nd = ctx.dimension(); dx, dy = ctx.upload(rng.standard_normal(nd)), ctx.vec_alloc()
    op = hp.Operator(ctx)
    st = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3): op.apply_device(dx, dy, sync=False)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(20): op.apply_device(dx, dy, sync=False)
    e1.record(st); ctx.sync(); torch.cuda.synchronize()
    print("variant", variant, "cfg3 apply us", e0.elapsed_time(e1)/20*1e3)
    ctx.close()
