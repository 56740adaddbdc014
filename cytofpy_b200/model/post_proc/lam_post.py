"""Posterior draw of the cluster label of every cell (0 = noisy class, 1..K = features);
API mirror of reference cytofpy/model/post_proc/lam_post.py:6-55.  The (N_i, K+1)
probabilities and the categorical draw are computed by cytof_label_sample_f32 (forward half
of the fused likelihood kernel, no gradient)."""
import ctypes as C

import torch

from ... import _lib
from ..engine import Batch, _ptr

_draw_counter = [0]


def sample(theta, mod, return_probs=False, seed=None):
    """theta, mod: output of sample_params.theta(mod, priors, y).  Returns a list of I int64
    tensors (N_i,) [and, with return_probs, the list of (N_i, K+1) probability matrices]."""
    eng, dev = mod.engine, mod.device
    I, J, K = mod.I, mod.J, mod.K
    f32 = lambda t: torch.as_tensor(t).to(device=dev, dtype=torch.float32).reshape(-1)
    eps = theta['eps'] if 'eps' in theta else torch.zeros(I)
    zero_ij = torch.zeros(I * J)
    g = torch.cat([f32(theta['mu0']), f32(theta['mu1']), f32(theta['sig']), f32(torch.log(theta['eta0'])),
                   f32(torch.log(theta['eta1'])), f32(torch.log(theta['W'])), f32(eps),
                   eng.b_dev.reshape(-1).float(), f32(zero_ij), f32(zero_ij)]).contiguous()
    bits = (torch.as_tensor(theta['Z']).to(dev) > 0.5).to(torch.int64)
    zmask = (bits << torch.arange(K, device=dev, dtype=torch.int64)[None, :]).sum(1).contiguous()
    ys = [torch.as_tensor(t).to(device=dev, dtype=torch.float32) for t in theta['y']]
    counts = [int(t.shape[0]) for t in ys]
    y = torch.cat(ys, 0).contiguous()
    _draw_counter[0] += 1
    batch = Batch(counts, None, None, step=_draw_counter[0])
    d = eng._desc(batch, noise_mode=_lib.NOISE_PHILOX)
    if seed is not None:
        d.seed = int(seed) & (2 ** 64 - 1)
    lam = torch.empty(sum(counts), dtype=torch.int32, device=dev)
    probs = torch.empty((sum(counts), K + 1), dtype=torch.float32, device=dev) if return_probs else None
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(eng.lib.cytof_label_sample_f32(C.byref(d), _ptr(y), _ptr(g), _ptr(zmask), _ptr(lam),
                                                 _ptr(probs), stream))
    lam_list = list(torch.split(lam.to(torch.int64), counts))
    if return_probs:
        return lam_list, list(torch.split(probs, counts))
    return lam_list
