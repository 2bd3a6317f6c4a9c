#!/usr/bin/env python
"""bench.py -- IPDG operator-apply throughput (DoF/s, FP64) on N B200s of one node.

A "step" is one pass of the hot path, y = A x, over one synthetic vector on the BASELINE.json
configs[1] workload: 3-D Poisson SIPG on a 64^3 structured mesh, uniform DG Q3 (16 777 216 DoF per GPU).
N > 1 is weak scaling: every rank owns a 64^3 brick of a (px,py,pz) mesh and exchanges face traces with
NCCL each step.  One JSON line on rank 0; see DESIGN.md section 6 for every field.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg5]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "dune-hpdg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (elements per direction per GPU, degree, description)
    "cfg2": (64, 3, "3D Poisson SIPG, 64^3 YaspGrid-like mesh per GPU, uniform DG Q3, operator apply"),
    "cfg5": (128, 4, "3D Poisson SIPG, 128^3 elements per GPU, uniform DG Q4, operator apply"),
}
PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
METRIC = "IPDG matvec DoF/s (3D Q3, FP64)"
NBUF = 3  # rotating (x, y) buffer pairs
BYTES_PER_DOF = 16  # algorithmic: read x once + write y once


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic(workload):
    try:
        s = json.load(open(os.path.join(ROOT, "profiles", "apply_uniform_ncu_summary.json")))
        return s.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def host_threads():
    """threads the CPU arm uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which is ignored here:
    the thread count is passed to the oracle explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(n, degree, budget_s, threads=None, min_reps=1):
    """Time the CPU restatement of the reference's matrix-free apply (Operator::apply over IPDGOperator,
    /root/reference/dune/hpdg/matrix-free/operator.hh:41, localoperators/ipdgoperator.hh:80) on an n^3 sample."""
    from oracle import orc
    threads = threads or host_threads()
    m = orc.Mesh((n, n, n), degree=degree, sigma=2.0, dirichlet=True)
    x = orc.fill_random(m.ndof)
    m.apply_mf(x, threads=threads)  # warm caches
    t0 = time.perf_counter()
    reps = 0
    while True:
        m.apply_mf(x, threads=threads)
        reps += 1
        el = time.perf_counter() - t0
        if reps >= min_reps and el >= budget_s:
            break
        if reps >= 50:
            break
    return m.ndof * reps / el, threads, m.ndof, reps, el


def workload_config(workload, world, p2p=True):
    nelem_dir, degree, desc = WORKLOADS[workload]
    ndof = nelem_dir ** 3 * (degree + 1) ** 3
    return {"workload": desc, "elements_per_gpu": [nelem_dir] * 3, "degree": degree, "dof_per_gpu": ndof,
            "pgrid": list(PGRID[world]), "sigma": 2.0, "dirichlet": True,
            "l2": f"{NBUF} rotating (x,y) pairs of {ndof * 8 / 1e6:.0f} MB each: inputs larger than the 126 MB L2",
            "halo": ("none" if world == 1 else "NVLink peer-memory stores of the face traces + step flags, issued by the tile kernel itself "
                     "before its first tile; its rank-boundary tiles (scheduled last) wait on the neighbours' flags" if p2p else
                     "NCCL send/recv of face traces overlapped with interior tiles")}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (here: its plain-C restatement, the
    reference being uncompilable in this image) on the host cores, all threads, a bounded sample of the workload per step.
    The sample and the thread count do not depend on the number of ranks, so the driver's ratios are comparable across N."""
    if rank != 0:
        return
    from oracle import orc
    nelem_dir, degree, desc = WORKLOADS[args.workload]
    threads = host_threads()
    rate, _, _, _, _ = cpu_reference_rate(8, degree, 0.5, threads=threads)
    # one brick of the workload per step if the whole run then stays within ~2.5 minutes, otherwise a 32^3 (16^3) sample
    n = 16
    for cand in (nelem_dir, 32):
        if cand <= nelem_dir and cand ** 3 * (degree + 1) ** 3 / rate * (args.steps + args.warmup) <= 150.0:
            n = cand
            break
    m = orc.Mesh((n, n, n), degree=degree, sigma=2.0, dirichlet=True)
    x = orc.fill_random(m.ndof)
    for _ in range(args.warmup):
        m.apply_mf(x, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.apply_mf(x, threads=threads)
    el = time.perf_counter() - t0
    val = m.ndof * args.steps / el
    sample = (f"{n}^3 elements Q{degree} ({m.ndof} DoF) per step" + (" = one GPU's brick of the workload" if n == nelem_dir else "") +
              f", matrix-free quadrature-loop apply (CPU restatement of ipdgoperator.hh), {threads} OpenMP threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "DoF/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world if world in PGRID else 1),
        "cpu_baseline": {"value": val, "unit": "DoF/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def make_context(hp, torch, dist, n, degree, world, rank, local_rank):
    """one rank's brick context (distributed when world > 1: NCCL id from rank 0, NVLink peer-memory halo when possible)"""
    from hpdg_b200 import partition as part
    if world == 1:
        return hp.Context(n, L=[1.0, 1.0, 1.0], degree=degree, sigma=2.0, dirichlet=True, device=local_rank), False
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        import ctypes
        buf = ctypes.create_string_buffer(128)
        assert hp.lib().hpdg_nccl_unique_id(buf) == 0
        idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    # the global domain is [0,px]x[0,py]x[0,pz] so that every brick is a unit cube
    ctx = hp.Context(n, L=[1.0, 1.0, 1.0], degree=degree, sigma=2.0, dirichlet=True, device=local_rank, pgrid=PGRID[world],
                     rank=rank, nranks=world, nccl_id=bytes(idt.cpu().tolist()))
    return ctx, part.enable_p2p_halo(ctx, dist, torch, world)


def parity_check(hp, torch, dist, degree, world, rank, local_rank):
    """Before any number is reported: the same code path (same kernels, same halo transport) on an 8^3-per-rank brick of the
    global (8 px, 8 py, 8 pz) mesh against the CPU oracle applied to the GLOBAL mesh; two chained applies so that both parities
    of the double-buffered halo arena are exercised.  Returns the worst relative L2 error over all ranks."""
    import numpy as np
    from hpdg_b200 import partition as part
    from oracle import orc
    nb, pgrid = (8, 8, 8), PGRID[world]
    m = orc.Mesh([nb[d] * pgrid[d] for d in range(3)], L=[float(pgrid[d]) for d in range(3)], degree=degree, sigma=2.0, dirichlet=True)
    xg = orc.fill_random(m.ndof)
    thr = max(1, host_threads() // world)
    r1 = m.apply_mf(xg, threads=thr)
    r2 = m.apply_mf(r1, threads=thr)
    ne = (degree + 1) ** 3
    ctx, _ = make_context(hp, torch, dist, nb, degree, world, rank, local_rank)
    dx, dy = ctx.upload(part.scatter_global_vector(xg, rank, pgrid, nb, ne)), ctx.vec_alloc()
    op = hp.Operator(ctx)
    op.apply_device(dx, dy)
    y1 = ctx.download(dy)
    op.apply_device(dy, dx)
    y2 = ctx.download(dx)
    ctx.vec_free(dx)
    ctx.vec_free(dy)
    ctx.close()
    e = 0.0
    for y, r in ((y1, r1), (y2, r2)):
        rl = part.scatter_global_vector(r, rank, pgrid, nb, ne)
        e = max(e, float(np.linalg.norm(y - rl) / np.linalg.norm(rl)))
    if world > 1:
        t = torch.tensor([e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e = float(t.item())
    return {"rel_l2": e, "tolerance": 1e-12, "n": int(m.ndof), "mesh": [nb[d] * pgrid[d] for d in range(3)], "degree": degree,
            "against": "CPU oracle (restatement of ipdgoperator.hh) on the global mesh, two chained applies", "ranks": world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-pointer leg (default min(steps, 50))")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:  # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import hpdg_b200 as hp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nelem_dir, degree, desc = WORKLOADS[args.workload]
    n = (nelem_dir,) * 3
    if world not in PGRID:
        raise SystemExit(f"unsupported GPU count {world}")
    pgrid = PGRID[world]
    parity = parity_check(hp, torch, dist, degree, world, rank, local_rank)
    if not parity["rel_l2"] < parity["tolerance"]:
        raise SystemExit(f"bench.py: parity check against the CPU oracle FAILED ({parity}); no number is reported")
    ctx, p2p = make_context(hp, torch, dist, n, degree, world, rank, local_rank)
    ndof = ctx.dimension()
    op = hp.Operator(ctx)
    assert ctx.uses_uniform_kernel()

    # synthetic input resident in HBM: NBUF (x, y) pairs rotated so no step re-reads what the previous one left in L2
    rng = np.random.default_rng(1887 + rank)
    hx, hx_ptr = ctx.host_alloc(ndof)
    hy, hy_ptr = ctx.host_alloc(ndof)
    hx[:] = rng.standard_normal(ndof)
    dxs = [ctx.upload(hx) for _ in range(NBUF)]
    dys = [ctx.vec_alloc() for _ in range(NBUF)]

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def aligned_start():
        """After the barrier the ranks leave it tens of microseconds apart, which a 20-step timed region of a 90 us step would book
        as step time (the ranks are coupled through the halo flags).  All ranks of the node share CLOCK_MONOTONIC: rank 0 names
        an instant 0.5 ms ahead, everyone spins until then and only then records its start event and launches."""
        if world > 1:
            t = torch.tensor([time.monotonic() + 0.0005], dtype=torch.float64, device="cuda")
            dist.broadcast(t, 0)
            target = float(t.item())
            torch.cuda.synchronize()
            while time.monotonic() < target:
                pass

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    # the clock sampler is started BEFORE the warm-up so that nothing but the barrier separates the warm-up steps from the timed
    # ones: a GPU that idles for the 0.3 s of the sampler's start-up needs ~100 us to get back to speed, which a 20-step region of an
    # 83 us step books as 4-5 us per step (measured: 87.6 against 82.9 us per step at 200 steps)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(args.warmup):
        op.apply_device(dxs[i % NBUF], dys[i % NBUF], sync=False)
    barrier()
    l0 = ctx.launch_count
    aligned_start()
    e0.record(stream)
    for i in range(args.steps):
        op.apply_device(dxs[i % NBUF], dys[i % NBUF], sync=False)
    e1.record(stream)
    ctx.sync()
    barrier()
    launches = ctx.launch_count - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = ndof * world / (ms_per_step * 1e-3)

    # dominant kernel alone (the tile kernel; at N=1 it is the whole step), timed on its launching stream
    kernel_ms = ctx.time_apply_device(dxs[0], dys[0], max(10, min(args.steps, 100)))

    # the smoother sweep of the same level (block Jacobi, fast-diagonalisation form: c = damping * D_e^-1 r per element), N = 1 only:
    # it needs no exchange, so every rank would repeat the same number
    smoother = None
    if world == 1:
        jac = hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)
        for i in range(args.warmup):
            jac.time_device(dxs[i % NBUF], dys[i % NBUF], 1)
        jsteps = max(10, min(args.steps, 100))
        j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lj = ctx.launch_count
        barrier()
        j0.record(stream)
        for i in range(jsteps):
            jac.apply_device(dxs[i % NBUF], dys[i % NBUF], sync=False)
        j1.record(stream)
        ctx.sync()
        jms = j0.elapsed_time(j1) / jsteps
        smoother = {"what": "block-Jacobi sweep, fast-diagonalisation form (exact D_e^-1), same level, rotating buffers",
                    "kernel": "hpdg_k_jacobi_fd_q3_persist" if degree == 3 else "k_jacobi_fd_uniform",
                    "value": ndof / (jms * 1e-3), "unit": "DoF/s", "ms_per_sweep": jms, "steps": jsteps,
                    "gpu_launches": ctx.launch_count - lj, "algorithmic_bytes_per_launch": BYTES_PER_DOF * ndof}

    # end-to-end leg: the drop-in call with HOST buffers, copies inside the timed region
    e2e_steps = args.e2e_steps or min(args.steps, 50)
    for _ in range(3):
        op.apply(hx_ptr, hy_ptr)
    barrier()
    aligned_start()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        op.apply(hx_ptr, hy_ptr)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_val = ndof * world * e2e_steps / t_e2e
    checksum = float(np.abs(hy[:1024]).sum())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_peak()
        # roofline of the dominant kernel over the SAME rotating buffers as the timed region (at N = 1 the step is that one launch;
        # at N > 1 the launch also packs / waits for the halo); kernel_ms_same_buffers (one fixed buffer pair, partly L2-warm) is
        # reported beside it
        achieved = BYTES_PER_DOF * ndof / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "DoF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args.workload, world, p2p),
            "parity": parity,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "DoF/s", "h2d_bytes_per_step": ndof * 8, "d2h_bytes_per_step": ndof * 8,
                    "steps": e2e_steps, "api": "hpdg_op_apply (host pointers, pinned)", "checksum": checksum},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload), "peak_source": peak_src,
                         "kernel": "hpdg_k_apply_q3_persist" if degree == 3 else "k_apply_uniform", "kernel_ms": ms_per_step,
                         "kernel_ms_same_buffers": kernel_ms, "algorithmic_bytes_per_launch": BYTES_PER_DOF * ndof},
        }
        if smoother:
            smoother["roofline_frac"] = smoother["algorithmic_bytes_per_launch"] / (smoother["ms_per_sweep"] * 1e-3) / 1e9 / peak
            out["smoother"] = smoother
        if world == 1 and not args.no_cpu_baseline:
            rate, thr, sdof, reps, el = cpu_reference_rate(32, degree, 10.0, threads=host_threads())
            out["cpu_baseline"] = {"value": rate, "unit": "DoF/s", "cores": thr, "kind": "port",
                                   "sample": f"32^3 elements Q{degree} ({sdof} DoF) x {reps} applies in {el:.1f} s, matrix-free "
                                             f"quadrature-loop apply (CPU restatement of ipdgoperator.hh), {thr} OpenMP threads"}
        print(json.dumps(out), flush=True)
    for d in dxs + dys:
        ctx.vec_free(d)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
