"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

fp64 numpy restatement of the O(cells) part of CytofPy's ELBO — imputation
(reference cytofpy/model/vae.py:17-41) + per-cell likelihood
(cytofpy/model/Model.py:210-278) — *and of its analytic gradient* with respect
to the small global inputs ("gradient block").  It takes exactly the inputs the
CUDA entry point `cytof_elbo_fwdbwd_f32` takes and returns exactly its outputs,
so the GPU parity tests compare array against array.

Pinning: tests/test_oracle_golden.py checks (a) ll / log_qy of this module
against the golden fixtures produced by the live reference
(oracle/gen_golden.py) and (b) the gradient block against torch autograd of
oracle/elbo_port.loglike_port (itself pinned to the reference fixtures).

Closed form (per sample i, cell n):
  y    = miss ? vm_ij + exp(vls_ij) * e_nj : ydata_nj
  a_zl = log eta_zijl - log sig_i - .5 log 2pi - (y - mu_zl)^2 / (2 sig_i^2)
  A_z  = LSE_l a_zl ; p_zl = exp(a_zl - A_z) ; D_j = A1_j - A0_j
  f_k  = sum_j A0_j + sum_j Z_jk D_j + log W_ik ; lse = LSE_k f ; r_k = exp(f_k - lse)
  q    = lse + log1p(-eps_i) ; z = sum_j logN(y; 0, s) + log eps_i
  L    = logaddexp(q, z) ; rho = exp(q - L)             (model_noisy=False: L = lse, rho = 1)
  ll   = sum_i fac_i sum_n (L + sum_{missing j} log sigmoid(b0 + b1 y + b2 y^2))
  lqy  = sum_i fac_i scale sum_{missing} (-.5 e^2 - vls_ij - .5 log 2pi)
"""
import numpy as np

LOG2PI = np.log(2.0 * np.pi)


def block_layout(I, J, K, L0, L1):
    """Offsets of the gradient-block sections (floats).  Mirrors
    include/cytof_b200.h `cytof_grad_block_layout`."""
    names = [('dmu0', L0), ('dmu1', L1), ('dsig', I), ('dlogeta0', I * J * L0),
             ('dlogeta1', I * J * L1), ('dlogW', I * K), ('dZ', J * K), ('deps', I),
             ('dvm', I * J), ('dvls', I * J), ('dlqy_dvls', I * J)]
    off, o = {}, 0
    for n, s in names:
        off[n] = (o, s)
        o += s
    off['_total'] = o
    return off


def _lse(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True))).squeeze(axis)


def _log_sigmoid(u):
    return -(np.maximum(-u, 0.0) + np.log1p(np.exp(-np.abs(u))))


def cells_fwdbwd(y_rows, eps, g, fac, scale, noisy_sd, model_noisy):
    """y_rows: list of (B_i,J) fp64 gathered rows (NaN = missing)
    eps    : list of (B_i,J) standard-normal draws (only missing entries matter)
    g      : dict mu0 (L0), mu1 (L1), sig (I), logeta0 (I,J,L0), logeta1 (I,J,L1),
             logW (I,K), Z (J,K) in {0,1}, eps (I) [if model_noisy], b (I,3),
             vm (I,J), vls (I,J)
    fac    : (I,) N_i / B_i
    returns ll, lqy (floats), block (1-d fp64 array, block_layout order), y_imp list."""
    I = len(y_rows)
    J, K = g['Z'].shape
    L0, L1 = g['mu0'].shape[0], g['mu1'].shape[0]
    lay = block_layout(I, J, K, L0, L1)
    out = np.zeros(lay['_total'])
    sec = lambda n: out[lay[n][0]:lay[n][0] + lay[n][1]]
    dmu = [sec('dmu0'), sec('dmu1')]
    dsig = sec('dsig')
    dle = [sec('dlogeta0').reshape(I, J, L0), sec('dlogeta1').reshape(I, J, L1)]
    dlw = sec('dlogW').reshape(I, K)
    dZ = sec('dZ').reshape(J, K)
    deps = sec('deps')
    dvm = sec('dvm').reshape(I, J)
    dvls = sec('dvls').reshape(I, J)
    dlq = sec('dlqy_dvls').reshape(I, J)
    Z = g['Z'].astype(np.float64)
    mus = [g['mu0'], g['mu1']]
    les = [g['logeta0'], g['logeta1']]
    ll, lqy, y_imp = 0.0, 0.0, []
    for i in range(I):
        yd = np.asarray(y_rows[i], dtype=np.float64)
        B = yd.shape[0]
        if B == 0:
            y_imp.append(yd)
            continue
        miss = np.isnan(yd)
        e = np.asarray(eps[i], dtype=np.float64)
        sd = np.exp(g['vls'][i])[None, :]
        y = np.where(miss, g['vm'][i][None, :] + sd * e, yd)
        y_imp.append(y)
        s = g['sig'][i]
        A, P, Dv = [], [], []
        for z in (0, 1):
            d = y[:, :, None] - mus[z][None, None, :]
            a = les[z][i][None] - np.log(s) - .5 * LOG2PI - d * d / (2 * s * s)
            Az = _lse(a, 2)
            A.append(Az)
            P.append(np.exp(a - Az[:, :, None]))
            Dv.append(d)
        D = A[1] - A[0]
        f = A[0].sum(1, keepdims=True) + D @ Z + g['logW'][i][None, :]
        lse = _lse(f, 1)
        r = np.exp(f - lse[:, None])
        if model_noisy:
            ei = g['eps'][i]
            q = lse + np.log1p(-ei)
            zz = (-.5 * LOG2PI - np.log(noisy_sd) - y * y / (2 * noisy_sd ** 2)).sum(1) + np.log(ei)
            L = np.logaddexp(q, zz)
            rho = np.exp(q - L)
            one_m_rho = np.exp(zz - L)
        else:
            L, rho, one_m_rho = lse, np.ones(B), np.zeros(B)
        b0, b1, b2 = g['b'][i]
        u = b0 + b1 * y + b2 * y * y
        ls_u = np.where(miss, _log_sigmoid(u), 0.0)
        ll += fac[i] * (L.sum() + ls_u.sum())
        lqy += fac[i] * scale * np.where(miss, -.5 * e * e - g['vls'][i][None, :] - .5 * LOG2PI, 0.0).sum()
        # ---- gradient ----
        gk = fac[i] * rho[:, None] * r                       # (B,K)
        dlw[i] += gk.sum(0)
        dZ += D.T @ gk
        c1 = gk @ Z.T                                        # (B,J)
        c = [fac[i] * rho[:, None] - c1, c1]
        dy = np.zeros_like(y)
        for z in (0, 1):
            w = c[z][:, :, None] * P[z]                      # (B,J,L)
            t = Dv[z] / (s * s)
            dle[z][i] += w.sum(0)
            dmu[z] += (w * t).sum((0, 1))
            dsig[i] += (w * (t * t * s - 1.0 / s)).sum()
            dy -= (w * t).sum(2)
        if model_noisy:
            deps[i] += fac[i] * (-rho / (1 - ei) + one_m_rho / ei).sum()
            dy += fac[i] * one_m_rho[:, None] * (-y / noisy_sd ** 2)
        sig_neg_u = np.exp(_log_sigmoid(-u))                 # 1 - sigmoid(u)
        dy += np.where(miss, fac[i] * sig_neg_u * (b1 + 2 * b2 * y), 0.0)
        dym = np.where(miss, dy, 0.0)
        dvm[i] += dym.sum(0)
        dvls[i] += (dym * sd * e).sum(0)
        dlq[i] += -fac[i] * scale * miss.sum(0)
    return ll, lqy, out, y_imp
