import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/dune-hpdg_b200')
import hpdg_b200 as hp
from oracle import orc
variant = int(sys.argv[1])
for p in (3, 1, 2, 5):
    for n, L, dirichlet in [((5,6,3),[1.0,1.5,0.5],True), ((8,8,8),None,False), ((9,4,13),None,True), ((1,1,1),None,True), ((16,12,8),None,True)]:
        m = orc.Mesh(n, L=L, degree=p, dirichlet=dirichlet)
        x = orc.fill_random(m.ndof)
        ref = m.apply_mf(x, threads=8)
        ctx = hp.Context(n, L=L, degree=p, dirichlet=dirichlet)
        ctx.set_option("variant", variant)
        y = hp.Operator(ctx).apply(x)
        err = np.linalg.norm(y-ref)/np.linalg.norm(ref)
        print(p, n, dirichlet, '%.2e'%err, 'OK' if err<1e-12 else 'FAIL')
