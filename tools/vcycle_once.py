import sys, os
sys.path.insert(0, '/root/repo/dune-hpdg_b200')
import numpy as np, hpdg_b200 as hp
ctx = hp.Context((64,)*3, degree=4); ctx.build_p_hierarchy()
nd = ctx.dimension(); dx, db = ctx.upload(np.zeros(nd)), ctx.upload(np.ones(nd))
mg = hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75)
mg.apply_device(dx, db); mg.apply_device(dx, db)
