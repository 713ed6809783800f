import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
import proto_admm as P
for (m, n) in [(500, 600), (500, 1500)]:
    X, idx, K = P.make_problem(m, n)
    ref = P.admm(X, idx, 0.5)
    print('n', n, 'ref iters', ref['iters'], flush=True)
    def skip_sat(it, mu, A, J):
        return A.copy() if mu >= 1e6 else J          # C := 0 in the saturated phase
    r = P.admm(X, idx, 0.5, perturb=skip_sat)
    print('  C:=0 in saturated phase: iters', r['iters'], 'relF(Z)', np.linalg.norm(r['Z']-ref['Z'])/np.linalg.norm(ref['Z']), flush=True)
    def half_sat(it, mu, A, J):
        return A - 0.5*(A - J) if mu >= 1e6 else J   # C := C/2
    r = P.admm(X, idx, 0.5, perturb=half_sat)
    print('  C:=C/2 in saturated phase: iters', r['iters'], 'relF(Z)', np.linalg.norm(r['Z']-ref['Z'])/np.linalg.norm(ref['Z']), flush=True)
