"""Krylov loop around the hot path on N GPUs (run under torchrun, one rank per GPU): block-Jacobi preconditioned CG on one cfg2
brick per rank (64^3 Q3, Dirichlet), device resident (hpdg_pcg_device).  Per iteration: one operator apply (halo exchange inside), one
fd block-Jacobi application, three dot products that each end in a 1-double ncclAllReduce on the context stream, two fused vector
updates.  Reports wall time per iteration (fixed iteration count, tol = 0 so that no rank stops early; check_every = maxit: the host
never reads a residual inside the loop) next to the time of the iteration's kernels alone (apply + Jacobi) -- the difference is
BLAS-1 + the all-reduce latency.  One JSON line on rank 0."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
import numpy as np
import torch
import torch.distributed as dist
import hpdg_b200 as hp
from hpdg_b200 import partition as part

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
n, p, its = (64, 64, 64), 3, int(os.environ.get("PCG_ITERS", 60))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert hp.lib().hpdg_nccl_unique_id(buf) == 0
        idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    ctx = hp.Context(n, degree=p, dirichlet=True, device=lr, pgrid=part.pgrid_for(world), rank=rank, nranks=world,
                     nccl_id=bytes(idt.cpu().tolist()))
    p2p = part.enable_p2p_halo(ctx, dist, torch, world)
else:
    ctx, p2p = hp.Context(n, degree=p, dirichlet=True, device=lr), False
nd = ctx.dimension()
b = np.random.default_rng(5 + rank).standard_normal(nd)
dx, db, dy = ctx.upload(np.zeros(nd)), ctx.upload(b), ctx.vec_alloc()


def sync():
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


cg = hp.ConjugateGradients(ctx, precond=hp.PRECOND_JACOBI, damping=1.0, tol=0.0, maxit=its, check_every=its)
cg.solve_device(dx, db)   # warm-up (setup of the Jacobi factors, NCCL channels)
ctx.upload(np.zeros(nd), dx)
sync()
t0 = time.perf_counter()
cg.solve_device(dx, db)
sync()
t_iter = (time.perf_counter() - t0) / its
op, jac = hp.Operator(ctx), hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=1.0)
ms_apply = ctx.time_apply_device(db, dy, 20)
ms_jac = jac.time_device(db, dy, 20)
vals = torch.tensor([t_iter, ms_apply, ms_jac], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
if rank == 0:
    t_iter, ms_apply, ms_jac = [float(v) for v in vals.tolist()]
    print(json.dumps({"what": "block-Jacobi PCG, 64^3 Q3 brick per GPU, device resident (hpdg_pcg_device)", "n_gpus": world,
                      "halo": "p2p" if p2p else ("nccl" if world > 1 else "none"), "iterations": its, "relres": cg.relres,
                      "us_per_iteration": t_iter * 1e6, "us_apply": ms_apply * 1e3, "us_jacobi": ms_jac * 1e3,
                      "us_blas1_and_allreduce": t_iter * 1e6 - ms_apply * 1e3 - ms_jac * 1e3,
                      "dot_allreduces_per_iteration": 3 if world > 1 else 0, "dof_per_gpu": nd}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
