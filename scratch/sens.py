import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
import proto_admm as P
X, idx, K = P.make_problem(500, 600)
ref = P.admm(X, idx, 0.5)
print('ref iters', ref['iters'])
rng = np.random.default_rng(0)
for phase, lo, hi in [('saturated (mu=1e6)', 1e6, 1e7), ('ramp (mu<1e6)', 0, 9.9e5)]:
    for eps in (1e-1, 1e-2, 1e-3, 1e-4):
        def pert(it, mu, A, J, eps=eps, lo=lo, hi=hi):
            if lo <= mu < hi:
                C = A - J
                return A - C * (1 + eps * rng.standard_normal(C.shape))
            return J
        r = P.admm(X, idx, 0.5, perturb=pert)
        print(phase, 'rel noise on C', eps, 'iters', r['iters'], 'relF(Z)', np.linalg.norm(r['Z']-ref['Z'])/np.linalg.norm(ref['Z']))
