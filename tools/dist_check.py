"""Multi-rank parity check (run under torchrun, one rank per GPU): every rank owns a brick, exchanges face traces
over NCCL inside hpdg_op_apply_device, and compares its rows with the CPU oracle applied to the GLOBAL mesh."""
import ctypes
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

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pgrid = part.pgrid_for(world)
THR = max(1, orc.max_threads() // world)   # every rank evaluates the oracle on the global mesh: share the host cores
ok = True
for p, n in [(3, (8, 8, 8)), (3, (6, 5, 7)), (4, (4, 6, 4)), (1, (8, 4, 4)), (2, (5, 5, 5))]:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert hp.lib().hpdg_nccl_unique_id(buf) == 0
        idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    N = [n[d] * pgrid[d] for d in range(3)]
    Lg = [float(pgrid[d]) for d in range(3)]
    m = orc.Mesh(N, L=Lg, degree=p, sigma=2.0, dirichlet=True)
    xg = orc.fill_random(m.ndof)
    ref = m.apply_mf(xg, threads=THR)
    ne = (p + 1) ** 3
    xl = part.scatter_global_vector(xg, rank, pgrid, n, ne)
    rl = part.scatter_global_vector(ref, rank, pgrid, n, ne)
    ctx = hp.Context(n, L=[1.0, 1.0, 1.0], degree=p, sigma=2.0, dirichlet=True, device=lr, pgrid=pgrid, rank=rank,
                     nranks=world, nccl_id=bytes(idt.cpu().tolist()))
    p2p = part.enable_p2p_halo(ctx, dist, torch, world)
    dx, dy = ctx.upload(xl), ctx.vec_alloc()
    op = hp.Operator(ctx)
    op.apply_device(dx, dy)
    y = ctx.download(dy)
    # repeated applies exercise the double-buffered halo arena / step flags: y2 = A (A x) against the oracle
    op.apply_device(dy, dx)
    op.apply_device(dy, dx)
    ctx.sync()
    y2 = ctx.download(dx)
    ref2 = part.scatter_global_vector(m.apply_mf(ref, threads=THR), rank, pgrid, n, ne)
    err2 = np.linalg.norm(y2 - ref2) / np.linalg.norm(ref2)
    ctx.upload(xl, dx)
    err = np.linalg.norm(y - rl) / np.linalg.norm(rl)
    # distributed dot product (Krylov allreduce)
    dd = ctx.dot_device(dx, dx)
    derr = abs(dd - xg @ xg) / (xg @ xg)
    # fast-diagonalisation block Jacobi on a rank-local brick needs no communication
    jref = part.scatter_global_vector(m.blockjacobi_apply(xg, factor=0.75), rank, pgrid, n, ne)
    jac = hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)
    jac.apply_device(dx, dy)
    jerr = np.linalg.norm(ctx.download(dy) - jref) / np.linalg.norm(jref)
    # distributed p-multigrid cycle (fd block-Jacobi smoothing) against the oracle's cycle on the global mesh
    verr = 0.0
    if p >= 2:
        lv = [m]
        nl = int(np.floor(np.log2(p)))
        for idx in range(nl - 1, -1, -1):
            lv.insert(0, lv[0].coarsen(p // ((nl - idx) * 2)))
        bg = orc.fill_random(m.ndof, seed=9)
        xr, rr = orc.vcycle(lv, None, np.zeros(m.ndof), bg, smoother=1, damping=0.75)
        ctx.build_p_hierarchy()
        assert ctx.num_levels == len(lv)
        dxv = ctx.upload(np.zeros(ctx.dimension()))
        dbv = ctx.upload(part.scatter_global_vector(bg, rank, pgrid, n, ne))
        hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75).apply_device(dxv, dbv)
        xv = ctx.download(dxv)
        xrl = part.scatter_global_vector(xr, rank, pgrid, n, ne)
        verr = np.linalg.norm(xv - xrl) / np.linalg.norm(xrl)
    good = err < 1e-12 and err2 < 1e-12 and derr < 1e-12 and jerr < 1e-11 and verr < 1e-10
    ok &= good
    print(f"rank {rank}/{world} p={p} brick={n}: halo={'p2p' if p2p else 'nccl'} apply {err:.2e} twice {err2:.2e} dot {derr:.2e} jacobi {jerr:.2e} vcycle {verr:.2e} {'OK' if good else 'FAIL'}", flush=True)
    ctx.close()
# ---- distributed hp: the hp mesh partitioned element-wise over the ranks (per-element degree map, variable-size halo blocks) ----
for n, pmax, dirichlet in [((4, 3, 5), 4, True), ((6, 6, 4), 6, False), ((3, 4, 4), 3, True)]:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert hp.lib().hpdg_nccl_unique_id(buf) == 0
        idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    N = [n[d] * pgrid[d] for d in range(3)]
    degg = np.random.default_rng(7).integers(1, pmax + 1, int(np.prod(N))).astype(np.int32)   # same on every rank
    m = orc.Mesh(N, L=[float(pgrid[d]) for d in range(3)], degree=degg, sigma=2.0, dirichlet=dirichlet)
    xg = orc.fill_random(m.ndof)
    ref = m.apply_mf(xg, threads=THR)
    ref2 = m.apply_mf(ref, threads=THR)
    jref = m.blockjacobi_apply(xg, factor=0.75)
    gl = part.local_to_global_elements(rank, pgrid, n)
    sc = lambda v: part.scatter_global_blocks(v, m.offsets, gl)
    ctx = hp.Context(n, L=[1.0, 1.0, 1.0], degree=degg[gl], sigma=2.0, dirichlet=dirichlet, device=lr, pgrid=pgrid, rank=rank,
                     nranks=world, nccl_id=bytes(idt.cpu().tolist()))
    dx, dy = ctx.upload(sc(xg)), ctx.vec_alloc()
    op = hp.Operator(ctx)
    op.apply_device(dx, dy)
    e1 = np.linalg.norm(ctx.download(dy) - sc(ref)) / np.linalg.norm(sc(ref))
    op.apply_device(dy, dx)
    e2 = np.linalg.norm(ctx.download(dx) - sc(ref2)) / np.linalg.norm(sc(ref2))
    ctx.upload(sc(xg), dx)
    derr = abs(ctx.dot_device(dx, dx) - xg @ xg) / (xg @ xg)
    jac = hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)
    jac.apply_device(dx, dy)
    jerr = np.linalg.norm(ctx.download(dy) - sc(jref)) / np.linalg.norm(sc(jref))
    good = e1 < 1e-12 and e2 < 1e-12 and derr < 1e-12 and jerr < 1e-11
    ok &= good
    print(f"rank {rank}/{world} hp p=1..{pmax} brick={n} dirichlet={dirichlet}: apply {e1:.2e} twice {e2:.2e} dot {derr:.2e} jacobi {jerr:.2e} {'OK' if good else 'FAIL'}", flush=True)
    ctx.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK", "PASS" if t.item() == 1 else "FAIL")
sys.exit(0 if t.item() == 1 else 1)
