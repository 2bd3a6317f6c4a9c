"""PCIe roofline of the end-to-end (host-pointer) apply: H2D-only, D2H-only and the pipelined hpdg_op_apply on cfg2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
import numpy as np
import hpdg_b200 as hp

ctx = hp.Context((64, 64, 64), degree=3)
nd = ctx.dimension()
hx, px = ctx.host_alloc(nd)
hy, py = ctx.host_alloc(nd)
hx[:] = np.random.default_rng(0).standard_normal(nd)
dx = ctx.vec_alloc()
op = hp.Operator(ctx)
def timeit(f, reps=10):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync()
    return (time.perf_counter() - t0) / reps
th = timeit(lambda: ctx.upload(hx, dx))
td = timeit(lambda: ctx.download(dx, out=hy))
te = timeit(lambda: op.apply(px, py))
gb = nd * 8 / 1e9
print(f"H2D {th*1e3:.2f} ms = {gb/th:.1f} GB/s   D2H {td*1e3:.2f} ms = {gb/td:.1f} GB/s   e2e apply {te*1e3:.2f} ms = {nd/te/1e9:.2f} GDoF/s "
      f"({2*gb/te:.1f} GB/s both directions; max(H2D,D2H) bound {nd/max(th,td)/1e9:.2f} GDoF/s)")
