"""Synthetic CyTOF data following the recipe of reference
cytofpy/util/simdata.py:11-75: W_i ~ Dir(a_W) with permuted columns, Z
block-structured when K divides J (else Bernoulli(v_k), v ~ Beta(alpha/K, 1)),
eta_z,ij ~ Dir(1/L_z), mu0 = -(0..L0-1) - 3 sig_max, mu1 = (0..L1-1) + 3 sig_max,
lambda_n ~ Cat(W_i) (sorted), gamma ~ Cat(eta), y_nj ~ N(mu_{Z[j,lambda_n]}[gamma], sig_i).

Vectorised (numpy) so that millions of cells are generated in seconds; the
random streams are NOT the reference's (parity tests use recorded fixtures)."""
import numpy as np
import torch

F64 = torch.float64


def _cat_rows(p, rng):
    """one categorical draw per row of p (n, L) -> (n,) int64"""
    c = np.cumsum(p, axis=-1)
    u = rng.random(p.shape[:-1] + (1,)) * c[..., -1:]
    return np.minimum((u > c).sum(-1), p.shape[-1] - 1)


def synth_cytof(N, J, a_W, L0, L1, sig=None, alpha=None, seed=0, dtype=np.float32):
    """Returns dict(y=list of (N_i,J) arrays, params=...) — numpy, `dtype` storage."""
    rng = np.random.default_rng(seed)
    I, K = len(N), len(a_W)
    a_W = np.asarray(a_W, dtype=np.float64)
    alpha = K if alpha is None else alpha
    v = rng.beta(alpha / K, 1.0, size=K)
    if J % K == 0 and J > K:
        Z = np.kron(np.eye(K), np.ones(J // K)).T
    else:
        Z = (rng.random((J, K)) < v[None, :]).astype(np.float64)
    W = rng.dirichlet(a_W, size=I)
    for i in range(I):
        W[i] = W[i, rng.permutation(K)]
    eta0 = rng.dirichlet(np.ones(L0) / L0, size=(I, J))
    eta1 = rng.dirichlet(np.ones(L1) / L1, size=(I, J))
    sig = np.ones(I) if sig is None else np.asarray(sig, dtype=np.float64)
    mu0 = -(np.arange(L0) + 3 * sig.max())
    mu1 = np.arange(L1) + 3 * sig.max()
    ys, lams = [], []
    for i in range(I):
        n = int(N[i])
        lam = np.sort(rng.choice(K, size=n, p=W[i] / W[i].sum()))
        Zi = Z[:, lam].T                                  # (n, J)
        u = rng.random((n, J, 1))
        g0 = np.minimum((u > np.cumsum(eta0[i], -1)[None]).sum(-1), L0 - 1)
        g1 = np.minimum((u > np.cumsum(eta1[i], -1)[None]).sum(-1), L1 - 1)
        mu = np.where(Zi > 0.5, mu1[g1], mu0[g0])
        ys.append((mu + sig[i] * rng.standard_normal((n, J))).astype(dtype))
        lams.append(lam)
    return {'y': ys, 'params': {'W': W, 'v': v, 'eta0': eta0, 'eta1': eta1, 'mu0': mu0, 'mu1': mu1,
                                'sig': sig, 'Z': Z, 'lam': lams}}


def inject_missing(y, beta, rate=0.3, seed=0):
    """Sets entries to NaN with probability c * sigmoid(b0 + b1 y + b2 y^2), c chosen so the
    overall NaN rate is `rate` (the reference's generator is a stub, simdata.py:79-86; the
    mechanism is its missingness model, Model.py:36-38).  In place; returns y."""
    rng = np.random.default_rng(seed)
    b0, b1, b2 = [float(b) for b in beta]
    ps = []
    for yi in y:
        a = np.asarray(yi, dtype=np.float64)
        ps.append(1.0 / (1.0 + np.exp(-(b0 + b1 * a + b2 * a * a))))
    mean_p = sum(p.sum() for p in ps) / sum(p.size for p in ps)
    c = min(rate / max(mean_p, 1e-12), 1.0 / max(max(p.max() for p in ps), 1e-12))
    for yi, p in zip(y, ps):
        yi[rng.random(p.shape) < c * p] = np.nan
    return y


def simdata(N=[300, 100, 200], J=25, a_W=[200., 500., 200., 100.], L0=5, L1=3, sig=None, alpha=None,
            seed=None):
    """API of reference cytofpy.util.simdata (simdata.py:11-75): fp64 torch tensors in
    {'data': {'y', 'm'}, 'params': {...}}.  seed=None derives one from numpy's global RNG."""
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    sig_np = None if sig is None else np.asarray(sig, dtype=np.float64)
    s = synth_cytof(N, J, a_W, L0, L1, sig=sig_np, alpha=alpha, seed=seed, dtype=np.float64)
    ys = [torch.tensor(a, dtype=F64) for a in s['y']]
    P = s['params']
    params = {k: torch.tensor(P[k], dtype=F64) for k in ('W', 'v', 'eta0', 'eta1', 'mu0', 'mu1', 'sig', 'Z')}
    params['lam'] = [torch.tensor(l) for l in P['lam']]
    return {'data': {'y': ys, 'm': [t.clone() for t in ys]}, 'params': params}


def simdata_with_missing_values(N=[300, 100, 200], J=25, a_W=[200., 500., 200., 100.], L0=5, L1=3,
                                sig=None, alpha=None, beta=(-16.98, -9.01, -1.29), rate=0.3, seed=None):
    """simdata + NaN injection (default beta = solve_beta([-5,-3.5,-2],[.05,.8,.05]))."""
    out = simdata(N=N, J=J, a_W=a_W, L0=L0, L1=L1, sig=sig, alpha=alpha, seed=seed)
    arrs = [t.numpy() for t in out['data']['y']]
    inject_missing(arrs, beta, rate=rate, seed=0 if seed is None else seed + 1)
    out['data']['m'] = [torch.isnan(t) for t in out['data']['y']]
    return out
