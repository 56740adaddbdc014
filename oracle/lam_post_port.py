"""TEST INFRASTRUCTURE — fp64 numpy restatement of the class probabilities of reference
cytofpy/model/post_proc/lam_post.py:20-51 (log_p0 = log eps + sum_j logN(y;0,sd_eps);
log_pk = log1p(-eps) + log W_k + sum_j [Z_jk logmix1_j + (1-Z_jk) logmix0_j]; softmax over K+1).
Pinned by tests/test_oracle_golden.py::test_lam_post_port_matches_reference (live reference,
CPU box only) and used by the GPU test of cytof_label_sample_f32."""
import numpy as np

LOG2PI = np.log(2 * np.pi)


def _lse(a, axis):
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def label_probs(theta):
    """theta: dict of numpy arrays mu0, mu1, eta0 (I,J,L0), eta1, sig (I), W (I,K), Z (J,K),
    eps (I), noisy_sd, y (list of (N_i,J)).  Returns list of (N_i, K+1) probabilities."""
    out = []
    for i, y in enumerate(theta['y']):
        s = theta['sig'][i]
        lm = []
        for mu, eta in ((theta['mu0'], theta['eta0']), (theta['mu1'], theta['eta1'])):
            d = y[:, :, None] - mu[None, None, :]
            lm.append(_lse(-0.5 * LOG2PI - np.log(s) - d * d / (2 * s * s) + np.log(eta[i])[None], 2))
        Z = theta['Z']
        f = lm[1] @ Z + lm[0] @ (1 - Z)
        sd = theta['noisy_sd']
        lp0 = np.log(theta['eps'][i]) + (-0.5 * LOG2PI - np.log(sd) - y * y / (2 * sd * sd)).sum(1, keepdims=True)
        lpk = np.log1p(-theta['eps'][i]) + np.log(theta['W'][i])[None, :] + f
        lp = np.concatenate([lp0, lpk], 1)
        out.append(np.exp(lp - _lse(lp, 1)[:, None]))
    return out
