"""Device-timed numbers for the non-headline BASELINE.json configs (cfg1, cfg3, cfg4) and the block-Jacobi forms.
One JSON line per measurement.  CUDA events on the context stream via torch.cuda.ExternalStream."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
import numpy as np
import torch
import hpdg_b200 as hp

ap = argparse.ArgumentParser()
ap.add_argument("--which", default="cfg1,cfg3,jacobi,cfg4small")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
PEAK = 6536.7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(ctx, fn, reps):
    st = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    ctx.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


which = a.which.split(",")
rng = np.random.default_rng(1887)
if "cfg1" in which:
    ctx = hp.Context((16, 16), degree=2)
    nd = ctx.dimension()
    dx, dy = ctx.upload(rng.standard_normal(nd)), ctx.vec_alloc()
    op = hp.Operator(ctx)
    ms = timed(ctx, lambda: op.apply_device(dx, dy, sync=False), 200)
    emit(config="cfg1 2D 16x16 Q2 apply (generic kernel)", ndof=nd, us=ms * 1e3, gdofs=nd / ms / 1e6)
if "cfg3" in which:
    n = (32, 32, 32)
    deg = rng.integers(1, 7, 32 ** 3).astype(np.int32)
    ctx = hp.Context(n, degree=deg)
    nd = ctx.dimension()
    dx, dy = ctx.upload(rng.standard_normal(nd)), ctx.vec_alloc()
    op = hp.Operator(ctx)
    ms = timed(ctx, lambda: op.apply_device(dx, dy, sync=False), a.reps)
    bytes_alg = 16 * nd + 12 * 32 ** 3
    emit(config="cfg3 3D 32^3 hp p in 1..6 apply (generic kernel, 6 degree buckets)", ndof=nd, us=ms * 1e3, gdofs=nd / ms / 1e6,
         gbs=bytes_alg / ms / 1e6, frac=bytes_alg / ms / 1e6 / PEAK)
    t0 = time.time()
    jf = hp.BlockJacobi(ctx, form=hp.JACOBI_FD)
    tf = time.time() - t0
    ms = timed(ctx, lambda: jf.apply_device(dx, dy, sync=False), a.reps)
    emit(config="cfg3 block-Jacobi fd apply", ndof=nd, us=ms * 1e3, gdofs=nd / ms / 1e6, setup_s=tf, bytes=jf.bytes,
         gbs=bytes_alg / ms / 1e6, frac=bytes_alg / ms / 1e6 / PEAK)
    t0 = time.time()
    jd = hp.BlockJacobi(ctx, form=hp.JACOBI_DENSE)
    td = time.time() - t0
    ms = timed(ctx, lambda: jd.apply_device(dx, dy, sync=False), a.reps)
    off = ctx.block_offsets()
    ne = np.diff(off).astype(np.float64)
    bj = float((8 * ne * ne + 16 * ne).sum())
    emit(config="cfg3 block-Jacobi dense apply", ndof=nd, us=ms * 1e3, gdofs=nd / ms / 1e6, setup_s=td, bytes=jd.bytes,
         gbs=bj / ms / 1e6, frac=bj / ms / 1e6 / PEAK)
    ctx.close()
if "jacobi" in which:
    for n, p in (((64, 64, 64), 3), ((32, 32, 32), 4)):
        ctx = hp.Context(n, degree=p)
        nd = ctx.dimension()
        dx, dy = ctx.upload(rng.standard_normal(nd)), ctx.vec_alloc()
        t0 = time.time()
        jd = hp.BlockJacobi(ctx, form=hp.JACOBI_DENSE)
        td = time.time() - t0
        ms = timed(ctx, lambda: jd.apply_device(dx, dy, sync=False), a.reps)
        ne = (p + 1) ** 3
        bj = float(np.prod(n)) * (8 * ne * ne + 16 * ne)
        emit(config=f"block-Jacobi dense apply {n[0]}^3 Q{p}", ndof=nd, us=ms * 1e3, gdofs=nd / ms / 1e6, setup_s=td, bytes=jd.bytes,
             gbs=bj / ms / 1e6, frac=bj / ms / 1e6 / PEAK)
        jf = hp.BlockJacobi(ctx, form=hp.JACOBI_FD)
        ms = timed(ctx, lambda: jf.apply_device(dx, dy, sync=False), a.reps)
        emit(config=f"block-Jacobi fd apply {n[0]}^3 Q{p}", ndof=nd, us=ms * 1e3, gdofs=nd / ms / 1e6, gbs=16 * nd / ms / 1e6,
             frac=16 * nd / ms / 1e6 / PEAK)
        ctx.close()
for tag, nn in (("cfg4small", 64), ("cfg4", 128)):
    if tag in which:
        ctx = hp.Context((nn,) * 3, degree=4)
        ctx.build_p_hierarchy()
        nd = ctx.dimension()
        b = np.ones(nd)
        dx, db = ctx.upload(np.zeros(nd)), ctx.upload(b)
        mg = hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75)
        t0 = time.time()
        mg.apply_device(dx, db)  # includes setup
        ts = time.time() - t0
        l0 = ctx.launch_count
        ms = timed(ctx, lambda: mg.apply_device(dx, db), 3)
        emit(config=f"{tag} 3D {nn}^3 Q4->Q2->Q1 V-cycle (5+5 damped block-Jacobi fd, 5 coarse its)", ndof=nd, ms=ms,
             gdofs=nd / ms / 1e6, first_call_s=ts, levels=[ctx.dimension(l) for l in range(ctx.num_levels)],
             launches_per_cycle=(ctx.launch_count - l0) // 6)
        ctx.close()
if "blockgs" in which:
    # the reference's default smoother without a matrix: one DynamicBlockGS sweep (hyperplane wavefronts) at cfg2 size
    for n, p in (((64, 64, 64), 3), ((32, 32, 32), 4)):
        ctx = hp.Context(n, degree=p)
        nd = ctx.dimension()
        dx, db = ctx.upload(np.zeros(nd)), ctx.upload(rng.standard_normal(nd))
        gs = hp.MatrixFreeBlockGS(ctx)
        l0 = ctx.launch_count
        gs.iterate_device(dx, db)
        nl = ctx.launch_count - l0
        ms = timed(ctx, lambda: gs.iterate_device(dx, db), 3)
        emit(config=f"matrix-free block-GS sweep {n[0]}^3 Q{p} (hyperplane wavefronts, generic element pass)", ndof=nd, ms=ms,
             gdofs=nd / ms / 1e6, launches_per_sweep=nl)
        ctx.close()
if "nc" in which:
    # non-conforming 2-D mesh: 256 x 256 base grid, every other cell in a checkerboard refined once, Q3
    nb = 256
    ref = ((np.add.outer(np.arange(nb), np.arange(nb)) % 2) == 0).astype(np.uint8).ravel()
    nleaf = int(ref.size + 3 * ref.sum())
    ctx = hp.Context.refined_2d((nb, nb), ref, np.full(nleaf, 3, dtype=np.int32))
    nd = ctx.dimension()
    dx, dy = ctx.upload(rng.standard_normal(nd)), ctx.vec_alloc()
    op = hp.Operator(ctx)
    ms = timed(ctx, lambda: op.apply_device(dx, dy, sync=False), a.reps)
    emit(config="non-conforming 2D 256^2 base grid, checkerboard refined once, Q3: apply (generic kernel, hanging faces)", ndof=nd,
         leaves=nleaf, us=ms * 1e3, gdofs=nd / ms / 1e6)
    ctx.close()
