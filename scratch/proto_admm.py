"""numpy prototype of the GPU-side ADMM formulation (sparse Z, support-only gradient,
stop check hoisted before the SVT, pluggable SVT).  Scratch: used to choose the SVT
algorithm/precision by measuring final-Z error against the literal oracle."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.linalg as sla
from oracle import rsoptsc_oracle as O

def svt_exact(A, tau):
    U, s, Vt = sla.svd(A, full_matrices=False, lapack_driver='gesdd')
    return (U * np.maximum(s - tau, 0)) @ Vt

def svt_gram(A, tau, dtype=np.float64, floor=None):
    A_ = A.astype(dtype)
    G = (A_.T @ A_).astype(dtype)
    lam, V = np.linalg.eigh(G)
    lam = lam.astype(np.float64); V = V.astype(np.float64)
    if floor is not None:
        lam = np.maximum(lam, floor * lam.max())
    sig = np.sqrt(np.maximum(lam, 0))
    with np.errstate(divide='ignore'):
        h = np.where(sig > tau, 1 - tau / sig, 0.0)
    keep = h > 0
    return (A @ (V[:, keep] * h[keep])) @ V[:, keep].T

def admm(X, idx, lam, svt=svt_exact, maxiter=100, verbose=False, perturb=None):
    m, n = X.shape
    K = idx.shape[1]
    rows = np.repeat(np.arange(n), K); cols = idx.ravel()
    Zs = np.zeros(n * K)
    Y3 = np.zeros((n, n)); Y1 = np.zeros((m, n)); Y2 = np.zeros(n); E = np.zeros((m, n))
    XXZ = X.copy()
    Err1 = np.zeros(maxiter)
    mu = 1e-6; eps = 1e-5
    eta = O.norm2(X, False) ** 2 + n + 1.0
    it = 0
    nsvt = 0
    while True:
        it += 1
        if it >= maxiter: stop = 'max iter'; break
        mu = min(5 * mu, 1e6); tau = 1.0 / mu
        if it >= 3 and (Err1[it - 2] >= Err1[it - 3] or O.norm2(XXZ, False) <= eps):
            stop = 'error min'; break
        A = -Y3 / mu
        A[rows, cols] += Zs
        if np.linalg.norm(A, 'fro') <= tau:
            J = np.zeros((n, n))
        else:
            J = svt(A, tau); nsvt += 1
            if perturb is not None:
                J = perturb(it, mu, A, J)
        QQ = XXZ + Y1 / mu
        nq = np.sqrt((QQ * QQ).sum(0))
        with np.errstate(invalid='ignore', divide='ignore'):
            sc = np.where(nq > lam / mu, (nq - lam / mu) / nq, 0.0)
        E = QQ * sc
        R = QQ - E
        colsum = np.bincount(cols, weights=Zs, minlength=n)
        r = 1.0 - colsum + Y2 / mu
        XtR_s = np.einsum('ij,ij->j', X[:, rows], R[:, cols]) if n * K * m < 3e8 else np.array([X[:, a] @ R[:, b] for a, b in zip(rows, cols)])
        Hs = -XtR_s - r[cols] + (Zs - J[rows, cols] + Y3[rows, cols] / mu)
        Zs = Zs - Hs / eta
        Zd = np.zeros((n, n)); Zd[rows, cols] = Zs
        XXZ = X - X @ Zd
        Y1 += mu * (XXZ - E)
        Y2 += mu * (1.0 - np.bincount(cols, weights=Zs, minlength=n))
        Y3 += mu * (Zd - J)
        Err1[it - 1] = O.norm2(XXZ - E, False)
        if verbose: print(it, mu, Err1[it - 1])
        # Err2 test omitted in the prototype (never fires: Err2 grows)
    Zd = np.zeros((n, n)); Zd[rows, cols] = Zs
    return dict(Z=Zd, E=np.linalg.norm(XXZ, 'fro'), iters=it, stop=stop, nsvt=nsvt)

def make_problem(m, n, seed=0, K=None):
    from rsoptsc_b200.synthetic import synthetic_lognorm
    data, lab = synthetic_lognorm(m, n, seed)
    X = O.normalize_columns(data)
    pe = O.prcomp_sdev2(X.T)
    K = K or O.choose_K(pe)
    emb = O.pca_embedding(X)
    idx = O.knn_indices(emb, K)
    return X, idx, K

if __name__ == '__main__':
    m, n = int(sys.argv[1]), int(sys.argv[2])
    X, idx, K = make_problem(m, n)
    print('K', K)
    t = time.time(); ref = O.computeM(O.mask_from_idx(idx, n), X, 0.5, literal=False); print('oracle', ref['iters'], ref['stop'], ref['E'], time.time() - t)
    np.save('/tmp/ref_Z_%d_%d.npy' % (m, n), ref['Z'])
    for name, f in [('exact', svt_exact), ('gram64', svt_gram), ('gram32', lambda A, t: svt_gram(A, t, np.float32)), ('gram32floor', lambda A, t: svt_gram(A, t, np.float32, 1e-7))]:
        t = time.time(); r = admm(X, idx, 0.5, svt=f)
        print(name, r['iters'], r['stop'], r['E'], 'nsvt', r['nsvt'], 'relF', np.linalg.norm(r['Z'] - ref['Z']) / np.linalg.norm(ref['Z']), time.time() - t)
