"""Step loop: drop-in for reference cytofpy/model/fit.py (`update` :32-48,
`fit` :50-147, `check_nan_in_grad` :12-30)."""
import copy
import datetime
import sys

import numpy as np
import torch

from .. import _lib
from .core import Model
from .fused import FusedStep, recognise_priors
from .priors import default_priors

import ctypes as C


def check_nan_in_grad(mp, key, fixed_grad, i=None):
    """reference fit.py:12-30 for the ModelParam blocks: NaN gradient entries -> 0, flag set."""
    vp = mp[key].vp
    if vp.grad is None:
        return
    bad = torch.isnan(vp.grad)
    if bool(bad.any()):
        print("WARNING: Setting a nan gradient to zero in {}!".format(key))
        vp.grad[bad] = 0.0
        fixed_grad[0] = True


def update(opt, mod, idx):
    """One optimisation step (reference fit.py:32-48).  Works with any torch optimizer over
    mod.parameters(); the NaN scrub covers the global blocks only, as in the reference."""
    elbo, ll, lp, lq = mod(idx)
    loss = -elbo / mod.Nsum
    opt.zero_grad()
    loss.backward()
    fixed_grad = [False]
    with torch.no_grad():
        grads = [mod.mp[key].vp.grad for key in mod.mp if mod.mp[key].vp.grad is not None]
        # one device sync for all blocks instead of one per block
        if grads and bool(torch.stack([torch.isnan(g).any() for g in grads]).any()):
            for key in mod.mp:
                check_nan_in_grad(mod.mp, key, fixed_grad)
    opt.step()
    return elbo, fixed_grad[0], ll, lp, lq


def device_minibatch(mod, minibatch_size, step, seed):
    """Without-replacement minibatch drawn ON the device (cytof_sample_minibatch): replaces the
    host-side np.random.choice(N_i, mb, replace=False) of reference fit.py:96-102.  Returns a
    list of int32 device tensors / `range` objects for full batches."""
    lib = _lib.load()
    out = []
    stream = torch.cuda.current_stream(mod.device).cuda_stream
    for i, n in enumerate(mod.N_local):
        if minibatch_size >= n:
            out.append(range(n))
            continue
        t = torch.empty(minibatch_size, dtype=torch.int32, device=mod.device)
        with torch.cuda.device(mod.device):
            _lib.check(lib.cytof_sample_minibatch(n, minibatch_size, seed, step, i,
                                                  C.c_void_p(t.data_ptr()), C.c_void_p(stream)))
        out.append(t)
    return out


def is_snapshot_step(t, backup_every, trace_every):
    """steps at which fit copies the parameters (reference fit.py:117-127)"""
    return (backup_every > 0 and t % backup_every == 0) or (trace_every > 0 and t % trace_every == 0)


def chunk_length(backup_every, trace_every, cap=32):
    """steps per multi-step graph: the smaller snapshot period (a chunk must END on a snapshot step)"""
    every = [v for v in (backup_every, trace_every) if v > 0]
    return max(1, min(min(every) if every else 16, cap))


def run_length(t, max_iter, backup_every, trace_every):
    """number of steps from t up to and including the next snapshot step (or the last step)"""
    s = t
    while s < max_iter and not is_snapshot_step(s, backup_every, trace_every):
        s += 1
    return s - t + 1


def step_schedule(max_iter, backup_every, trace_every, chunk):
    """[(first step, number of steps)] as the pipelined fit loop issues them: step 0 alone (host-driven),
    then a chunk of `chunk` steps wherever no snapshot step lies strictly inside it, single steps elsewhere"""
    out, t = [], 0
    while t <= max_iter:
        n = chunk if (t >= 1 and chunk > 1 and run_length(t, max_iter, backup_every, trace_every) >= chunk) else 1
        out.append((t, n))
        t += n
    return out


def check_exchange(scrubbed):
    """Code 2 in the scrub slot of a fused step = the peer-memory exchange of the packed gradient block gave up on
    this or another rank (a rank died or stalled for ~2 s).  Nothing was updated; every rank sees the same code from
    that step on, so all of them stop here instead of training on as diverged replicas."""
    if scrubbed == 2.0:
        raise RuntimeError('cytofpy_b200: the multi-GPU exchange of the gradient block timed out (a rank stalled or '
                           'died); no update was applied on any rank from that step on')


def fit(y, minibatch_size=500, priors=None, max_iter=1000, lr=1e-1, print_freq=10, seed=1,
        y_mean_init=-3.0, y_sd_init=0.1, trace_every=None, eps_conv=1e-6, tau=0.1,
        backup_every=10, y_quantiles=[0, 35, 70], p_bounds=[.05, .8, .05], scale=1,
        use_stick_break=True, model_noisy=True, init_mp=None, verbose=1, flush=True,
        K=30, L=None, device=None, sampler='device', noise='philox', fused=True, graph=True,
        pipeline=True):
    """reference fit.py:50-147, same arguments and the same returned dict keys
    ('elbo', 'trace', 'mp', 'tau', 'vae', 'use_stick_break', 'y', 'priors', 'model_noisy').

    Extras: K / L (used when priors is None — the reference's `priors=None` branch raises
    NameError, fit.py:68-70), device, sampler ('device' = on-GPU keyed permutation,
    'numpy' = the reference's np.random.choice stream), noise ('philox' | 'external'), fused
    (True: the whole step runs as 4-5 kernel launches, cytofpy_b200/model/fused.py, when the
    priors are of the default families; False / other priors: torch autograd + torch Adam
    around the fused per-cell kernel), graph (True: with the fused path and the device sampler,
    every step after the first is ONE CUDA-graph replay, FusedStep.capture), pipeline (True: on the
    graph path the steps between two backup / trace snapshots run as ONE graph of min(backup_every,
    trace_every) consecutive steps with one read-back for all of them, so the GPU never idles on the
    host round trip; histories and snapshots are identical to pipeline=False -- only when the
    convergence rule fires inside such a chunk have its remaining steps already been applied to the
    live imputation parameters returned in 'vae')."""
    print('scale: {}'.format(scale))
    torch.manual_seed(seed)
    np.random.seed(seed)
    if trace_every is None:
        trace_every = int(max_iter / 50) if max_iter >= 50 else 1
    m = [torch.isnan(yi) for yi in y]
    if priors is None:
        priors = default_priors(y, K=K, L=L, y_quantiles=y_quantiles, p_bounds=p_bounds)
    model = Model(y=y, m=m, priors=priors, tau=tau, verbose=verbose, use_stick_break=use_stick_break,
                  model_noisy=model_noisy, y_mean_init=y_mean_init, y_sd_init=y_sd_init, scale=scale,
                  device=device, noise=noise, seed=seed)
    if init_mp is not None:
        # warm start that actually trains the given state (the reference swaps `model.mp` after
        # the optimizer captured the old tensors, fit.py:76-83, so its resume is a no-op)
        with torch.no_grad():
            for key in model.mp:
                model.mp[key].vp.copy_(init_mp[key].vp.to(model.device))
    if verbose >= 1.1:
        print('state_dict:')
        print(model.state_dict())
    stepper = None
    if fused and noise == 'philox' and recognise_priors(model.priors, model_noisy) is not None:
        stepper = FusedStep(model, lr=lr, seed=seed)
    else:
        optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    elbo_hist, trace = [], []
    state = {'best_mp': copy.deepcopy(model.mp)}

    def finish(t, elbo_t, ll, lp, lq, fixed_grad):
        """bookkeeping of step t (reference fit.py:109-141); True = stop"""
        elbo_hist.append(elbo_t)
        elbo_good = not fixed_grad
        if t % print_freq == 0:
            print('{} | {}/{} | elbo: {:.3f} | ll: {:.3f} | lp: {:.3f} | lq: {:.3f}'.format(
                datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"), t, max_iter,
                elbo_hist[-1] / model.Nsum, ll / model.Nsum, lp / model.Nsum, lq / model.Nsum))
        if backup_every > 0 and t % backup_every == 0 and elbo_good:
            state['best_mp'] = copy.deepcopy(model.mp)
        if trace_every > 0 and t % trace_every == 0 and elbo_good:
            trace.append({key: copy.deepcopy(model.mp[key]) for key in model.mp})
        if t > 10 and abs(elbo_hist[-1] / elbo_hist[-2] - 1) < eps_conv:
            print('Convergence suspected! Ending optimizer early.')
            return True
        if flush:
            sys.stdout.flush()
        return False

    # `pipeline`: between two snapshot steps nothing on the host depends on the parameters, so those steps
    # run as ONE graph of `chunk` consecutive steps (FusedStep.graph_steps) and their results are read back
    # together; a chunk always ENDS on the snapshot step, so every backup / trace copy holds exactly the
    # parameters the reference's loop would have copied
    replayable = stepper is not None and graph and sampler == 'device'
    chunk = chunk_length(backup_every, trace_every) if (pipeline and replayable) else 1
    t = 0
    while t <= max_iter:
        replay = replayable and t >= 1
        if replay and stepper.graph is None:
            stepper.capture(minibatch_size, repeat=chunk)
        if replay and chunk > 1 and run_length(t, max_iter, backup_every, trace_every) >= chunk:
            rows = stepper.graph_steps().tolist()                 # one launch, one read-back per chunk
            stop = False
            for j, (elbo_t, ll, lp, lq, scrubbed) in enumerate(rows):
                check_exchange(scrubbed)
                if scrubbed != 0.0:
                    print("WARNING: Setting a nan gradient to zero!")
                if finish(t + j, elbo_t, ll, lp, lq, scrubbed != 0.0):
                    stop = True         # the remaining steps of the chunk have run; their results are dropped
                    break
            t += chunk
            if stop:
                break
            continue
        if replay:
            idx = None      # the captured graph draws its own minibatch on the device
        elif sampler == 'numpy':
            idx = [np.arange(n) if minibatch_size >= n else
                   np.random.choice(n, minibatch_size, replace=False) for n in model.N_local]
        else:
            idx = device_minibatch(model, minibatch_size, t, seed)
        if stepper is not None:
            out_t = stepper.graph_step() if replay else stepper.step(idx)
            elbo_t, ll, lp, lq, scrubbed = out_t.tolist()     # one 40-byte read-back
            check_exchange(scrubbed)
            fixed_grad = scrubbed != 0.0
            if fixed_grad:
                print("WARNING: Setting a nan gradient to zero!")
        else:
            elbo, fixed_grad, ll, lp, lq = update(optimizer, model, idx)
            elbo_t = elbo.item()
        if finish(t, elbo_t, ll, lp, lq, fixed_grad):
            break
        t += 1
    # like the reference, the result lives on the CPU and pickles without a GPU: 'mp' = the last good backup,
    # 'vae' = the imputation parameters at return (reference fit.py:145-147 returns its live model.y_vae)
    best_mp = {key: blk.to('cpu') for key, blk in state['best_mp'].items()}
    best_vae = [v.to_cpu_copy() for v in model.y_vae]
    trace = [{key: blk.to('cpu') for key, blk in snap.items()} for snap in trace]
    return {'elbo': elbo_hist, 'trace': trace, 'mp': best_mp, 'tau': tau, 'vae': best_vae,
            'use_stick_break': use_stick_break, 'y': y, 'priors': str(model.priors),
            'model_noisy': model_noisy}
