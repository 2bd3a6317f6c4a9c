"""Generates tests/golden/oracle_golden.npz from the CPU oracle (oracle/hpdg_oracle.c, oracle/sf2d.py).

The reference itself cannot be built in this image and ships no golden vectors, so these fixtures freeze the
ORACLE's outputs (which are pinned by the reference's differential tests, tests/test_oracle_pins.py) on small seeded cases:
they guard the oracle against regressions (CPU test) and give the CUDA path a fixed target that travels to the GPU box
(GPU test).  Inputs follow the reference's fixtures: x = interp(|x|^2) (matrix-free/test/testdg.cc:97) or
mt19937(1887) + normal_distribution (test/randomvector.hh:11-21).  Run:  python tests/golden/gen_oracle_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc

out = {}
# cfg1: 2-D 16x16 Q2, sigma 2, Dirichlet: apply on |x|^2, then 10 block-GS sweeps with b = 1, x0 = 1
m = orc.Mesh((16, 16), degree=2, sigma=2.0, dirichlet=True)
A = m.assemble()
x = m.interpolate_normsq()
out["cfg1_x"] = x
out["cfg1_Ax"] = m.apply_mf(x)
b, xg = np.ones(m.ndof), np.ones(m.ndof)
for _ in range(10):
    A.blockgs_iterate(b, xg)
out["cfg1_gs10"] = xg
# 3-D uniform Q3 on a ragged anisotropic brick, both boundary types
for tag, dirichlet in (("d", True), ("n", False)):
    m = orc.Mesh((5, 6, 3), L=[1.0, 1.5, 0.5], degree=3, sigma=2.0, dirichlet=dirichlet)
    x = orc.fill_random(m.ndof)
    out[f"q3_{tag}_Ax"] = m.apply_mf(x)
    out[f"q3_{tag}_jac"] = m.blockjacobi_apply(x, factor=0.75)
# 3-D hp, degrees 1..6 from numpy's default_rng(1887)
rng = np.random.default_rng(1887)
deg = rng.integers(1, 7, 4 * 3 * 5).astype(np.int32)
m = orc.Mesh((4, 3, 5), degree=deg, sigma=2.0, dirichlet=True)
x = orc.fill_random(m.ndof)
out["hp_deg"] = deg
out["hp_Ax"] = m.apply_mf(x)
out["hp_jac"] = m.blockjacobi_apply(x, factor=1.0)
# p-multigrid pieces on 4^3 Q4: restrict / prolong / one V-cycle with damped exact block Jacobi
fine = orc.Mesh((4, 4, 4), degree=4)
l1 = fine.coarsen(2)
l0 = l1.coarsen(1)
xf = orc.fill_random(fine.ndof)
out["mg_restrict"] = fine.restrict(l1, xf)
out["mg_prolong"] = fine.prolong(l1, out["mg_restrict"])
bvec = orc.fill_random(fine.ndof, seed=5)
xv, rv = orc.vcycle([l0, l1, fine], None, np.zeros(fine.ndof), bvec, smoother=1, damping=0.75)
out["mg_vcycle_x"] = xv
out["mg_vcycle_r"] = rv
# the third formulation (oracle/sf2d.py: sum-factorised Gauss-Lobatto operator of sfipdg.hh) on a 2-D hp mesh, and its non-conforming
# branch on a once-refined mesh with hanging nodes (leaf numbering of hpdg_create_refined_2d)
from oracle import sf2d
deg2 = np.random.default_rng(5).integers(1, 6, 42).astype(np.int32)
s2 = sf2d.SumFactIPDG2D((7, 6), (1.0, 2.0), deg2, 2.0, True)
x2 = orc.fill_random(s2.ndof)
out["sf2d_deg"] = deg2
out["sf2d_Ax"] = s2.apply(x2)
ref = np.zeros(20, dtype=np.uint8)
ref[[1, 6, 7, 12, 18]] = 1
degn = np.random.default_rng(17).integers(1, 6, 20 + 3 * 5).astype(np.int32)
nc = sf2d.RefinedSumFactIPDG2D((5, 4), ref, degn, (1.0, 1.5), 2.0, True)
out["nc_refine"] = ref
out["nc_deg"] = degn
out["nc_Ax"] = nc.apply(orc.fill_random(nc.ndof))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
