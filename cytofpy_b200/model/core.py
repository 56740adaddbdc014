"""`Model`: drop-in for reference cytofpy.model.Model (cytofpy/model/Model.py:114-346).

Same constructor arguments, attributes (`mp`, `y_vae`, `N`, `Nsum`, `I`, `J`, `K`,
`L`, `m`, `y_data`, `b0/b1/b2`, `noisy_sd`, `tau`, `scale`, `use_stick_break`,
`model_noisy`, `log_qy`, `priors`) and methods (`forward`, `loglike`, `log_prior`,
`log_q`, `sample_reals`, `transform`, `sample_params`).  What differs is where the
O(cells) work runs: imputation + likelihood + their gradient are ONE fused CUDA
kernel call (cytof_elbo_fwdbwd_f32) wrapped in a torch.autograd.Function; only the
O(3.5k-scalar) global part (sampling, transforms, prior, entropy) stays in torch
(fp64, on the device) so that `v > H` is thresholded in fp64 exactly as the
reference does.
"""
import math

import numpy as np
import torch
from torch.distributions import Beta
from torch.distributions import normal as _tdn
from torch.nn import ParameterList

from .engine import Batch, CellEngine
from .priors import compute_Z, dist_to
from .variational import F64, ModelParam, VAE


class _CellELBO(torch.autograd.Function):
    """(ll, log_qy) = fused per-cell ELBO terms; backward multiplies the
    kernel's gradient block by the upstream gradients (a handful of tiny ops)."""

    @staticmethod
    def forward(ctx, model, batch, mu0, mu1, sig, le0, le1, lw, Z, eps_n, vm, vls):
        eng = model.engine
        g = torch.cat([mu0.reshape(-1), mu1.reshape(-1), sig.reshape(-1), le0.reshape(-1),
                       le1.reshape(-1), lw.reshape(-1), eps_n.reshape(-1), eng.b_dev.reshape(-1),
                       vm.reshape(-1), vls.reshape(-1)]).to(torch.float32)
        bits = (Z.detach() > 0.5).to(torch.int64)
        zmask = (bits << torch.arange(eng.K, device=bits.device, dtype=torch.int64)[None, :]).sum(1)
        blk32, ll_lqy = eng.fwdbwd(batch, g, zmask.contiguous(), model._counts_global(batch))
        out = torch.cat([blk32.to(F64), ll_lqy])
        if model.dist_group is not None:
            torch.distributed.all_reduce(out, group=model.dist_group)
        ctx.blk = out[:-2]
        ctx.off = eng.b_off
        ctx.shapes = [t.shape for t in (mu0, mu1, sig, le0, le1, lw, Z, eps_n, vm, vls)]
        return out[-2].clone(), out[-1].clone()

    @staticmethod
    def backward(ctx, g_ll, g_lq):
        o, blk, sh = ctx.off, ctx.blk, ctx.shapes
        sec = lambda n: blk[o[n]:o[n + 1]]
        grads = [g_ll * sec(n).view(sh[n]) for n in range(9)]
        grads.append((g_ll * sec(9) + g_lq * sec(10)).view(sh[9]))
        return (None, None) + tuple(grads)


class Model(torch.nn.Module):
    def __init__(self, y, priors, m=None, y_mean_init=-3.0, y_sd_init=0.1, tau=0.1, verbose=1,
                 use_stick_break=True, model_noisy=True, scale=1, device=None, noise='philox',
                 seed=0, N_global=None, row_global=None, dist_group=None):
        """Arguments up to `scale` are the reference's (Model.py:115-116).  Extra, all optional:
        device (default cuda:current), noise ('philox' in-kernel draws | 'external' torch
        draws, used by the parity tests), seed (Philox key), and for cell-sharded multi-GPU
        runs N_global / row_global / dist_group (y then holds only this rank's rows)."""
        super().__init__()
        self.model_noisy = model_noisy
        self.verbose = verbose
        if self.verbose >= 0:
            print('use_stick_break: {}'.format(use_stick_break))
            print('tau: {}'.format(tau))
            print('model_noisy: {}'.format(model_noisy))
        self.use_stick_break = use_stick_break
        self.I, self.J = priors['I'], priors['J']
        self.N = list(N_global) if N_global is not None else priors['N']
        self.Nsum = sum(self.N)
        self.scale = scale
        self.tau = tau
        self.b0, self.b1, self.b2 = priors['b0'], priors['b1'], priors['b2']
        self.L, self.K = priors['L'], priors['K']
        self.noisy_sd = torch.sqrt(torch.tensor(priors['noisy_var'], dtype=F64))
        self.priors = priors
        self.m = [torch.isnan(yi) for yi in y] if m is None else m
        self.y_data = y
        self.log_qy = None
        self.msum = [mi.sum() for mi in self.m]
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device('cuda')
        self.device = torch.device(device)
        self.noise = noise
        self.dist_group = dist_group
        self._step = 0
        self._prior_cache = {}
        self._counts_cache = {}

        dev, I, J, K, L = self.device, self.I, self.J, self.K, self.L
        ones = lambda n: torch.ones(n, dtype=F64)
        # construction order = reference Model.py:169-192 (it fixes the order of the N(0,1) init draws)
        mp = {}
        mp['delta0'] = ModelParam(L[0], 'positive', m=ones(L[0]), s=ones(L[0]), device=dev)
        mp['delta1'] = ModelParam(L[1], 'positive', m=ones(L[1]), s=ones(L[1]), device=dev)
        if model_noisy:
            mp['eps'] = ModelParam(I, 'unit_interval', m=ones(I) * priors['eps'].mean,
                                   s=ones(I) * .001, device=dev)
        mp['sig2'] = ModelParam(I, 'positive', m=ones(I) * -1.0, s=ones(I) * .1, device=dev)
        mp['eta0'] = ModelParam((I, J, L[0] - 1), 'simplex', device=dev)
        mp['eta1'] = ModelParam((I, J, L[1] - 1), 'simplex', device=dev)
        mp['W'] = ModelParam((I, K - 1), 'simplex', device=dev)
        mp['alpha'] = ModelParam(1, 'positive', device=dev)
        mp['v'] = ModelParam(K, 'unit_interval', ones(K) * .99 if use_stick_break else None, device=dev)
        mp['H'] = ModelParam((J, K), 'unit_interval', device=dev)
        self._mp = mp
        for key, p in mp.items():
            setattr(self, key + '_vp', p.vp)
        self._y_vae = [VAE(J, mean_init=y_mean_init, sd_init=y_sd_init, device=dev) for _ in range(I)]
        self.y_vae_vp = ParameterList([q for vae in self._y_vae for q in vae.parameters()])

        # device-resident data + C-ABI engine (masks given separately are folded into NaN)
        y_eff = []
        for yi, mi in zip(y, self.m):
            t = yi.detach().to(torch.float32).clone()
            t[mi.to(torch.bool)] = float('nan')
            y_eff.append(t)
        b = torch.stack([torch.as_tensor(self.b0, dtype=F64), torch.as_tensor(self.b1, dtype=F64),
                         torch.as_tensor(self.b2, dtype=F64)], 1)
        self.engine = self._make_engine(y_eff, dev, K, L[0], L[1], model_noisy, float(self.noisy_sd),
                                        float(scale), b, N_global=N_global, row_global=row_global, seed=seed)
        self.N_local = self.engine.N_local

    # `mod.mp = out['mp']; mod.y_vae = out['vae']` is how the reference's callers restore a fitted model
    # (sims/cb/cb.py:173-179).  Here the registered parameters (and the fused step's flat buffer they
    # may be views of) stay in place on the device and the assigned values are copied into them.
    @property
    def mp(self):
        return self._mp

    @mp.setter
    def mp(self, new):
        with torch.no_grad():
            for key, p in self._mp.items():
                p.vp.copy_(new[key].vp.to(device=p.vp.device, dtype=p.vp.dtype))

    @property
    def y_vae(self):
        return self._y_vae

    @y_vae.setter
    def y_vae(self, new):
        if len(new) != len(self._y_vae):
            raise ValueError('y_vae: expected %d samples, got %d' % (len(self._y_vae), len(new)))
        with torch.no_grad():
            for mine, theirs in zip(self._y_vae, new):
                mine.mean.copy_(theirs.mean.to(device=mine.mean.device, dtype=mine.mean.dtype))
                mine.log_sd.copy_(theirs.log_sd.to(device=mine.log_sd.device, dtype=mine.log_sd.dtype))

    def _make_engine(self, *args, **kwargs):
        """The per-cell engine: always the CUDA one (no CPU fallback; CellEngine raises without the
        library or a CUDA device).  The CPU-only host-logic tests subclass Model under tests/."""
        return CellEngine(*args, **kwargs)

    # ------------------------------------------------------------------ batches
    def _counts_global(self, batch):
        """Per-sample cell counts of the step summed over ranks (for fac_i = N_i / B_i)."""
        if self.dist_group is None:
            return None
        key = tuple(batch.counts)
        hit = self._counts_cache.get(key)
        if hit is None:      # counts are static per (minibatch size, shard): one collective, once
            on_cpu = torch.distributed.get_backend(self.dist_group) == 'gloo'
            t = torch.tensor(batch.counts, dtype=torch.int64, device='cpu' if on_cpu else self.device)
            torch.distributed.all_reduce(t, group=self.dist_group)
            hit = self._counts_cache[key] = [int(v) for v in t.tolist()]
        return hit

    def _make_batch(self, idx):
        counts, idx_dev = self.engine.make_idx(idx)
        eps = None
        if self.noise == 'external':
            # one (B_i, J) standard-normal draw per sample, in sample order (reference
            # Model.py:314-317 / vae.py:33) through the same torch hook the reference uses
            draws = [_tdn._standard_normal((c, self.J), dtype=F64, device=self.device) for c in counts]
            eps = torch.cat(draws, 0).to(torch.float32).contiguous()
        self._step += 1
        return Batch(counts, idx_dev, eps, step=self._step)

    # ------------------------------------------------------------------ global part (torch, fp64)
    def _sample_globals(self):
        return {key: self.mp[key].real_sample() for key in self.mp}

    def transform(self, reals):
        params = {key: self.mp[key].transform(reals[key]) for key in self.mp}
        if 'y' in reals:
            params['y'] = reals['y']
        return params

    def _kernel_inputs(self, params):
        sig = torch.sqrt(params['sig2'])
        mu0 = -torch.cumsum(params['delta0'], 0)
        mu1 = torch.cumsum(params['delta1'], 0)
        v = torch.cumprod(params['v'], 0) if self.use_stick_break else params['v']
        Z = compute_Z(v, params['H'], self.tau)
        eps_n = params['eps'] if self.model_noisy else torch.zeros(self.I, dtype=F64, device=self.device)
        vm = torch.cat([vae.mean for vae in self.y_vae], 0)
        vls = torch.cat([vae.log_sd for vae in self.y_vae], 0)
        return (mu0, mu1, sig, torch.log(params['eta0']), torch.log(params['eta1']),
                torch.log(params['W']), Z, eps_n, vm, vls)

    def _cells(self, params, batch):
        return _CellELBO.apply(self, batch, *self._kernel_inputs(params))

    def _prior(self, key):
        d = self.priors[key]
        hit = self._prior_cache.get(key)
        if hit is None or hit[0] is not d:
            hit = (d, dist_to(d, self.device))
            self._prior_cache[key] = hit
        return hit[1]

    def log_prior(self, reals, params, idx=None):
        """reference Model.py:289-304"""
        lp = 0.0
        for key in self.mp:
            if key == 'v':
                conc = params['alpha'] if self.use_stick_break else params['alpha'] / self.K
                one = torch.ones((), dtype=F64, device=self.device)
                t = Beta(conc, one, validate_args=False).log_prob(params['v'])
            else:
                t = self._prior(key).log_prob(params[key])
            lp = lp + (t + self.mp[key].logabsdetJ(reals[key], params[key])).sum()
        return lp

    def log_q(self, reals, idx=None):
        """reference Model.py:280-287 (entropy term incl. the cached imputation part)"""
        lq = self.log_qy if self.log_qy is not None else 0.0
        for key in self.mp:
            lq = lq + self.mp[key].log_q(reals[key])
        return lq

    # ------------------------------------------------------------------ ELBO
    def forward(self, idx):
        """reference Model.py:339-346: returns (elbo tensor with grad, ll, lp, lq floats)."""
        reals = self._sample_globals()
        batch = self._make_batch(idx)
        params = self.transform(reals)
        ll, self.log_qy = self._cells(params, batch)
        lp = self.log_prior(reals, params, idx)
        lq = self.log_q(reals, idx)
        elbo = ll + lp - lq
        vals = torch.stack([ll.detach(), lp.detach(), lq.detach()]).tolist()   # one sync, not three
        return elbo, vals[0], vals[1], vals[2]

    def loglike(self, params, idx):
        """reference Model.py:210-278.  `params['y']` (list of imputed (B_i, J) minibatches, as
        returned by sample_params) fixes the imputed values: the kernel is driven with the noise
        that reproduces them."""
        counts, idx_dev = self.engine.make_idx(idx)
        vm = torch.cat([vae.mean for vae in self.y_vae], 0).detach()
        sd = torch.exp(torch.cat([vae.log_sd for vae in self.y_vae], 0)).detach()
        eps = torch.cat([(params['y'][i].to(self.device, F64) - vm[i:i + 1]) / sd[i:i + 1]
                         for i in range(self.I)], 0).to(torch.float32).contiguous()
        self._step += 1
        ll, _ = self._cells(params, Batch(counts, idx_dev, eps, step=self._step))
        return ll

    def sample_reals(self, idx):
        """reference Model.py:306-322; 'y' = list of imputed minibatches (emit kernel)."""
        reals = self._sample_globals()
        batch = self._make_batch(idx)
        with torch.no_grad():
            g = torch.zeros(self.engine.g_off[-1], dtype=torch.float32, device=self.device)
            o = self.engine.g_off
            g[o[8]:o[9]] = torch.cat([vae.mean for vae in self.y_vae], 0).reshape(-1).float()
            g[o[9]:o[10]] = torch.cat([vae.log_sd for vae in self.y_vae], 0).reshape(-1).float()
            y_imp = self.engine.impute(batch, g).to(F64)
            noise = batch.eps if batch.eps is not None else self.engine.philox_noise(batch)
            miss = torch.isnan(self._gather_rows(batch))
            vls = torch.cat([vae.log_sd for vae in self.y_vae], 0)
            lq, start = 0.0, 0
            for i, c in enumerate(batch.counts):
                sl = slice(start, start + c)
                term = -0.5 * noise[sl].to(F64) ** 2 - vls[i:i + 1] - 0.5 * math.log(2 * math.pi)
                fac = self.N[i] / c if c else 0.0
                lq = lq + torch.where(miss[sl], term, torch.zeros_like(term)).sum() * fac * self.scale
                start += c
            self.log_qy = lq
        reals['y'] = list(torch.split(y_imp, batch.counts, 0))
        return reals

    def _gather_rows(self, batch):
        eng = self.engine
        if batch.idx is None:
            return eng.y_dev
        base = torch.repeat_interleave(
            torch.as_tensor(eng.row_base[:-1], device=self.device),
            torch.as_tensor(batch.counts, device=self.device))
        return eng.y_dev[base + batch.idx.long()]

    def sample_params(self, idx):
        """used for post processing (reference Model.py:332-337)"""
        return self.transform(self.sample_reals(idx))
