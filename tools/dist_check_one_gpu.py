"""Two ranks on ONE GPU (a 1-GPU box has no NCCL pair: NCCL refuses two ranks on one device): both processes create their brick
context on device 0 without an NCCL communicator, exchange the halo arenas' CUDA IPC handles over gloo and run the distributed
operator apply through the peer-memory halo (pack + flags fused into the tile kernel, rank-boundary tiles waiting on the
neighbour's flags).  The two processes time-slice the GPU, so a waiting tile simply spins until the other process has run.
Every rank compares its rows with the CPU oracle applied to the GLOBAL mesh."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
import numpy as np
import torch
import torch.distributed as dist
import hpdg_b200 as hp
from hpdg_b200 import partition as part
from oracle import orc

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
pgrid = part.pgrid_for(world)
ok = True
for p, n in [(3, (8, 8, 8)), (3, (6, 5, 7)), (4, (4, 6, 4)), (2, (5, 5, 5))]:
    N = [n[d] * pgrid[d] for d in range(3)]
    m = orc.Mesh(N, L=[float(pgrid[d]) for d in range(3)], degree=p, sigma=2.0, dirichlet=True)
    xg = orc.fill_random(m.ndof)
    thr = max(1, orc.max_threads() // world)
    ref = m.apply_mf(xg, threads=thr)
    ref2 = m.apply_mf(ref, threads=thr)
    ne = (p + 1) ** 3
    ctx = hp.Context(n, L=[1.0, 1.0, 1.0], degree=p, sigma=2.0, dirichlet=True, device=0, pgrid=pgrid, rank=rank, nranks=world,
                     nccl_id=None)
    ctx.set_option("halo_timeout_ms", 8000)
    if not part.enable_p2p_halo(ctx, dist, torch, world):
        print(f"rank {rank}: peer-memory halo not available between two processes on this GPU", flush=True)
        ok = False
        break
    dx, dy = ctx.upload(part.scatter_global_vector(xg, rank, pgrid, n, ne)), ctx.vec_alloc()
    op = hp.Operator(ctx)
    op.apply_device(dx, dy)
    y1 = ctx.download(dy)
    op.apply_device(dy, dx)            # second step: the other parity of the double-buffered arena
    y2 = ctx.download(dx)
    e1 = np.linalg.norm(y1 - part.scatter_global_vector(ref, rank, pgrid, n, ne)) / np.linalg.norm(ref) * np.sqrt(world)
    e2 = np.linalg.norm(y2 - part.scatter_global_vector(ref2, rank, pgrid, n, ne)) / np.linalg.norm(ref2) * np.sqrt(world)
    good = e1 < 1e-12 and e2 < 1e-12
    ok &= good
    print(f"rank {rank}/{world} p={p} brick={n}: one-GPU p2p halo, apply {e1:.2e} twice {e2:.2e} {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()                     # nobody tears its arena down while the neighbour may still read it
    ctx.close()
t = torch.tensor([1 if ok else 0])
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK_ONE_GPU", "PASS" if t.item() == 1 else "FAIL")
sys.exit(0 if t.item() == 1 else 1)
