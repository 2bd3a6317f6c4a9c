"""N > 1 path.  CPU: world_size-2 gloo test of the host-side partition / halo layout logic.  GPU: 2-rank NCCL parity
(skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_distributed.py -m gpu`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, n, p, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
    from hpdg_b200 import partition as part
    from oracle import orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pgrid = part.pgrid_for(world)
        N1 = p + 1
        ne = N1 ** 3
        Ng = [n[d] * pgrid[d] for d in range(3)]
        g_end = np.array([[orc.lagrange_prime(p, i, s) for i in range(N1)] for s in range(2)])
        xg = orc.fill_random(int(np.prod(Ng)) * ne)          # same global vector on every rank
        xl = part.scatter_global_vector(xg, rank, pgrid, n, ne)
        peers = part.peers(rank, pgrid)
        # exchange: send my face-f traces to the peer across f, receive the peer's (its face f^1) traces
        recv = {}
        for f, peer in peers.items():
            if peer is None:
                continue
            send = torch.from_numpy(part.face_traces(xl, n, p, f, g_end))
            buf = torch.empty_like(send)
            if rank < peer:
                dist.send(send, peer); dist.recv(buf, peer)
            else:
                dist.recv(buf, peer); dist.send(send, peer)
            recv[f] = buf.numpy()
        # check against traces computed from the peer's slice of the global vector
        ok = True
        for f, peer in peers.items():
            if peer is None:
                continue
            xp = part.scatter_global_vector(xg, peer, pgrid, n, ne)
            expect = part.face_traces(xp, n, p, f ^ 1, g_end)
            ok &= np.array_equal(recv[f], expect)
            # and, element by element, against the definition: (g_{1-s} . line, line[end]) of the neighbour element
            d, s = f // 2, f % 2
            gl = part.local_to_global_elements(rank, pgrid, n)
            stride = [1, Ng[0], Ng[0] * Ng[1]][d]
            e_loc = {0: (n[0] - 1 if s else 0), 1: (n[1] - 1 if s else 0) * n[0], 2: (n[2] - 1 if s else 0) * n[0] * n[1]}[d]
            eg = gl[e_loc] + (stride if s else -stride)        # neighbour of the first face element
            blk = xg.reshape(-1, N1, N1, N1)[eg]                # [k][j][i]
            line = {0: blk[0, 0, :], 1: blk[0, :, 0], 2: blk[:, 0, 0]}[d]
            ok &= abs(recv[f][0] - g_end[1 - s] @ line) < 1e-12 and recv[f][1] == line[0 if s else N1 - 1]
        q.put((rank, bool(ok), len(recv)))
    except Exception as exc:  # surface worker failures instead of timing out
        q.put((rank, False, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_partition_and_halo_layout_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29541
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, (3, 2, 4), 2, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1], res
    assert all(r[1] for r in res) and all(r[2] == 1 for r in res), res


def _gloo_hp_worker(rank, world, port, n, pmax, q):
    """distributed hp on 2 ranks over gloo: one-time degree exchange, receive offsets, variable-size trace blocks"""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
    from hpdg_b200 import partition as part
    from oracle import orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pgrid = part.pgrid_for(world)
        Ng = [n[d] * pgrid[d] for d in range(3)]
        degg = np.random.default_rng(7).integers(1, pmax + 1, int(np.prod(Ng))).astype(np.int64)   # same on every rank
        offg = np.concatenate(([0], np.cumsum((degg + 1) ** 3)))
        xg = orc.fill_random(int(offg[-1]))
        g_end_of = lambda p: np.array([[orc.lagrange_prime(p, i, s) for i in range(p + 1)] for s in range(2)])
        gl = part.local_to_global_elements(rank, pgrid, n)
        degl = degg[gl]
        offl = np.concatenate(([0], np.cumsum((degl + 1) ** 3)))
        xl = part.scatter_global_blocks(xg, offg, gl)
        ok, nrecv = True, 0
        for f, peer in part.peers(rank, pgrid).items():
            if peer is None:
                continue
            fel = part.face_elements(n, f)
            # 1. degree exchange (once per level)
            sdeg = torch.from_numpy(np.ascontiguousarray(degl[fel]))
            rdeg = torch.empty_like(sdeg)
            if rank < peer:
                dist.send(sdeg, peer); dist.recv(rdeg, peer)
            else:
                dist.recv(rdeg, peer); dist.send(sdeg, peer)
            roff = part.hp_halo_offsets(rdeg.numpy())
            # 2. per apply: variable-size blocks, sizes known to both sides from the degrees
            send = torch.from_numpy(part.hp_face_traces(xl, offl, degl, n, f, g_end_of))
            assert send.numel() == 2 * part.hp_halo_offsets(degl[fel])[-1]
            buf = torch.empty(2 * int(roff[-1]), dtype=torch.float64)
            if rank < peer:
                dist.send(send, peer); dist.recv(buf, peer)
            else:
                dist.recv(buf, peer); dist.send(send, peer)
            nrecv += 1
            # 3. the received degrees and blocks are those of the elements across the face in the GLOBAL mesh
            d, s = f // 2, f % 2
            stride = [1, Ng[0], Ng[0] * Ng[1]][d]
            for i, e in enumerate(fel):
                eg = gl[e] + (stride if s else -stride)
                po = int(degg[eg])
                ok &= int(rdeg[i]) == po
                N = po + 1
                blk = xg[offg[eg]:offg[eg + 1]].reshape(N, N, N)
                g = g_end_of(po)[1 - s]
                end = 0 if s else N - 1
                if d == 0:
                    der, val = np.einsum("kji,i->kj", blk, g), blk[:, :, end]
                elif d == 1:
                    der, val = np.einsum("kji,j->ki", blk, g), blk[:, end, :]
                else:
                    der, val = np.einsum("kji,k->ji", blk, g), blk[end]
                got = buf.numpy()[2 * roff[i]:2 * roff[i + 1]].reshape(-1, 2)
                ok &= np.allclose(got[:, 0], der.reshape(-1), rtol=0, atol=1e-12) and np.array_equal(got[:, 1], val.reshape(-1))
        q.put((rank, bool(ok), nrecv))
    except Exception as exc:
        q.put((rank, False, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_hp_halo_layout_gloo():
    # distributed hp (hpdg_create_distributed_hp): the host-side protocol of csrc/api.cu (hp_ghost_setup: degree exchange + receive
    # offsets; hp_halo_exchange: variable-size blocks) restated in hpdg_b200.partition and run on 2 ranks over gloo
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_hp_worker, args=(r, 2, 29543, (3, 2, 4), 4, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1], res
    assert all(r[1] for r in res) and all(r[2] == 1 for r in res), res


def test_partition_index_logic():
    sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
    from hpdg_b200 import partition as part
    for world in (1, 2, 4, 8):
        pg = part.pgrid_for(world)
        n = (3, 2, 2)
        seen = np.concatenate([part.local_to_global_elements(r, pg, n) for r in range(world)])
        assert sorted(seen) == list(range(world * 12))          # every element owned exactly once
        for r in range(world):
            for f, peer in part.peers(r, pg).items():
                if peer is not None:
                    assert part.peers(peer, pg)[f ^ 1] == r       # symmetric neighbour relation
    with pytest.raises(ValueError):
        part.pgrid_for(3)


@pytest.mark.gpu
def test_two_rank_nccl_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29551", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
def test_two_ranks_one_gpu_peer_memory_halo():
    # runs on a 1-GPU box: two processes share device 0, no NCCL communicator, halo through CUDA IPC peer memory
    # (tools/dist_check_one_gpu.py); covers pack + flags + waiting rank-boundary tiles of the distributed operator apply
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tools", "dist_check_one_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "DIST_CHECK_ONE_GPU PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
