"""Host-side brick partition of a structured mesh over the GPUs of one node.

Models the reference's element-wise owner/overlap decomposition (dune/hpdg/parallel/communicationhpdg.hh:261-263):
every element is owned by exactly one rank; what crosses rank boundaries per operator apply is, per ghost face node,
the pair (der, val) of the neighbour's DoF line normal to the face ("face traces") instead of whole ghost blocks
(communicationhpdg.hh:309-326).  Pure index logic, no arithmetic on DoFs except `face_traces` (test helper layout
contract for the pack kernel k_pack_traces).
"""
import numpy as np

PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def pgrid_for(nranks):
    if nranks not in PGRID:
        raise ValueError(f"unsupported rank count {nranks}; supported: {sorted(PGRID)}")
    return PGRID[nranks]


def coords(rank, pgrid):
    return (rank % pgrid[0], (rank // pgrid[0]) % pgrid[1], rank // (pgrid[0] * pgrid[1]))


def rank_of(c, pgrid):
    return c[0] + pgrid[0] * (c[1] + pgrid[1] * c[2])


def peers(rank, pgrid):
    """face index f = 2*dir + side -> neighbour rank or None (domain boundary)"""
    c = coords(rank, pgrid)
    out = {}
    for d in range(3):
        for s in range(2):
            cc = list(c)
            cc[d] += 1 if s else -1
            out[2 * d + s] = rank_of(cc, pgrid) if 0 <= cc[d] < pgrid[d] else None
    return out


def local_to_global_elements(rank, pgrid, n):
    """global (x-fastest) element index of every local element of the rank's n[0] x n[1] x n[2] brick"""
    c = coords(rank, pgrid)
    N = [n[d] * pgrid[d] for d in range(3)]
    ix = np.arange(n[0]) + c[0] * n[0]
    iy = np.arange(n[1]) + c[1] * n[1]
    iz = np.arange(n[2]) + c[2] * n[2]
    g = ix[None, None, :] + N[0] * (iy[None, :, None] + N[1] * iz[:, None, None])
    return g.reshape(-1)


def scatter_global_vector(xg, rank, pgrid, n, ndof_per_elem):
    """rank-local DynamicBlockVector slice (uniform block size) of a global vector"""
    g = local_to_global_elements(rank, pgrid, n)
    return np.ascontiguousarray(xg.reshape(-1, ndof_per_elem)[g].reshape(-1))


def scatter_global_blocks(xg, offsets, gl):
    """rank-local DynamicBlockVector slice of a global vector with per-element block sizes (hp): `offsets` are the global
    block offsets, `gl` the global indices of the rank's elements in local order"""
    return np.ascontiguousarray(np.concatenate([xg[offsets[g]:offsets[g + 1]] for g in gl]))


def face_traces(x_local, n, p, f, g_end):
    """(der, val) traces the rank SENDS across brick face f = 2*dir+side, in the layout the receiving kernel expects:
    [face element (lower tangential direction fastest)][face node (lower fastest)][2].
    g_end[s] = derivatives l_i'(s) of the 1-D basis at end point s (unit interval)."""
    N = p + 1
    d, s = f // 2, f % 2
    u = x_local.reshape(n[2], n[1], n[0], N, N, N)  # [ez][ey][ex][k][j][i]
    if d == 0:
        blk = u[:, :, n[0] - 1 if s else 0]          # [ez][ey][k][j][i]
        der = np.einsum("zykji,i->zykj", blk, g_end[s])
        val = blk[..., N - 1 if s else 0]
    elif d == 1:
        blk = u[:, n[1] - 1 if s else 0]             # [ez][ex][k][j][i]
        der = np.einsum("zxkji,j->zxki", blk, g_end[s])
        val = blk[:, :, :, N - 1 if s else 0, :]
    else:
        blk = u[n[2] - 1 if s else 0]                # [ey][ex][k][j][i]
        der = np.einsum("yxkji,k->yxji", blk, g_end[s])
        val = blk[:, :, N - 1 if s else 0]
    return np.ascontiguousarray(np.stack([der, val], axis=-1).reshape(-1))


# ---- distributed hp: the layout contract of the variable-size halo (csrc/api.cu: hp_ghost_setup / hp_halo_exchange) ----------------
def face_elements(n, f):
    """local indices of the brick's elements on face f = 2*dir+side, lower tangential direction fastest (the face-element numbering of
    the ghost layer)"""
    d, s = f // 2, f % 2
    idx = np.arange(n[0] * n[1] * n[2]).reshape(n[2], n[1], n[0])
    sl = [slice(None)] * 3
    sl[2 - d] = n[d] - 1 if s else 0
    return np.ascontiguousarray(idx[tuple(sl)].reshape(-1))


def hp_halo_offsets(face_degrees, dim=3):
    """offsets (in (der, val) pairs) of the variable-size trace blocks of a face's elements: block i holds (p_i + 1)^(dim-1) pairs;
    the receiver builds them from the degrees it got in the one-time degree exchange (parallel/updatedegrees.hh:11-46)"""
    return np.concatenate(([0], np.cumsum((np.asarray(face_degrees) + 1) ** (dim - 1))))


def hp_face_traces(x_local, offsets, degrees, n, f, g_end_of):
    """(der, val) traces the rank SENDS across brick face f for a per-element degree map: the face elements' blocks in
    face-element order, each [face node (lower tangential index fastest)][2].  g_end_of(p)[s] = l_i'(s) of degree p."""
    d, s = f // 2, f % 2
    out = []
    for e in face_elements(n, f):
        p = int(degrees[e])
        N = p + 1
        blk = x_local[offsets[e]:offsets[e + 1]].reshape(N, N, N)          # [k][j][i]
        g = g_end_of(p)[s]
        if d == 0:
            der, val = np.einsum("kji,i->kj", blk, g), blk[:, :, N - 1 if s else 0]
        elif d == 1:
            der, val = np.einsum("kji,j->ki", blk, g), blk[:, N - 1 if s else 0, :]
        else:
            der, val = np.einsum("kji,k->ji", blk, g), blk[N - 1 if s else 0]
        out.append(np.stack([der, val], axis=-1).reshape(-1))
    return np.ascontiguousarray(np.concatenate(out))


def enable_p2p_halo(ctx, dist, torch, world):
    """all-gather the ranks' halo-arena IPC handles and attach (NVLink peer-memory halo).  HPDG_HALO=nccl keeps NCCL send/recv.
    If any rank cannot map its neighbours' arenas (no peer access / IPC not permitted) every rank falls back to NCCL.
    Works over any torch.distributed backend (NCCL: CUDA tensors, gloo: host tensors)."""
    import os
    if os.environ.get("HPDG_HALO", "p2p") != "p2p" or world == 1:
        return False
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    ok = 1
    try:
        mine = torch.tensor(list(ctx.halo_ipc_handle()), dtype=torch.uint8, device=dev)
    except Exception:
        mine, ok = torch.zeros(64, dtype=torch.uint8, device=dev), 0
    allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine)
    if ok:
        try:
            ctx.halo_ipc_attach([bytes(t.cpu().tolist()) for t in allh])
        except Exception:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    use = bool(flag.item())
    if ok:
        ctx.set_option("halo_p2p", 1 if use else 0)
    dist.barrier()
    return use
