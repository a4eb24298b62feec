"""APG sweep for the GMM task: drop-in for the hot-path functions of apgs/gmm/objectives.py
(apg_objective :9-53, oneshot :74-110, apg_update_joint :112-158, apg_update_eta :161-195,
apg_update_z :197-227, resample_variables :229-234).

Differences a caller can observe: none in names, argument meaning or returned shapes.
Differences in how the work is done:
  * each block update evaluates its encoder ONCE and scores both the new and the old values
    against the same proposal parameters (the reference re-runs the encoder with identical
    inputs for the backward kernel, objectives.py:174,210,131,135);
  * every log-density, the N-sums and the f - b difference are one fused kernel (csrc/gmm.cu);
  * softmax / ESS over S is one kernel (W1); resampling is ancestors + one gather launch for
    (tau, mu, z) with no int64 index expansion;
  * `x` may be passed as (B,N,D); the reference's S-replicated (S,B,N,D) is accepted and only
    its first slice is read;
  * inside `apg_objective` (decomposed block, evaluation flags) the assignments travel as uint8 labels (S,B,N) instead
    of one-hot (S,B,N,K) floats -- the objective returns only the trace, so the one-hot form is never needed between
    the one-shot and the last sweep; traces are identical bit for bit (`compact_labels` switches it off).  Called on
    their own, the step functions take and return the reference's one-hot tensors.
Gradients: the kernels are forward-only.  The APG loss is  sum_s w_s.detach() * (-log q_s), so the
only graph autograd needs is that of the FORWARD proposal log-density; when `loss_required` is set and the encoder
parameters require grad, that term is rebuilt with PyTorch ops (models.posterior_eta_torch / eta_log_q_torch / the
un-fused Enc_apg_z MLP) and multiplied by the detached kernel weights -- the reference's train() works unchanged and
gets the reference's gradients, while every forward quantity still comes from the kernels.
"""
import math

import torch
from torch.distributions.gamma import Gamma
from torch.distributions.normal import Normal
from torch.distributions.one_hot_categorical import OneHotCategorical as cat

from .. import ops
from ..rng import normal_sample
from . import kernels
from .models import eta_log_q_torch, needs_grad


def _sample_eta(post):
    q_alpha, q_beta, q_mu, q_nu = post
    # validate_args=False: the default argument validation calls .item() on the device (a host sync per
    # distribution); it carries no arithmetic
    tau = Gamma(q_alpha, q_beta, validate_args=False).sample()
    mu = normal_sample(q_mu, 1. / (q_nu * tau).sqrt())
    return tau, mu


def _book_weights(trace, result_flags, log_q_f, stats):
    """loss / ess / log-marginal from the normalised weights (objectives.py:93-98,181-186).  stats = (w, ess, lse) of the
    step's log-weights.  log Z-hat_b = logsumexp_s log_w - log S (the reference never materialises it; it
    is the `lse` the softmax kernel computes anyway) is appended to trace['log_z'] next to the ESS."""
    w, ess, lse = stats
    if result_flags['loss_required']:
        trace['loss'].append((w * (-log_q_f)).sum(0).mean().unsqueeze(0))
    if result_flags['ess_required']:
        trace['ess'].append(ess.unsqueeze(0))
        S = log_q_f.shape[0]
        trace.setdefault('log_z', []).append((lse - math.log(S)).unsqueeze(0))


def _book(trace, result_flags, log_w, log_q_f, density, defer=False):
    """loss / ess / density bookkeeping shared by the steps (objectives.py:93-109,181-194).  defer: the caller takes the
    weight statistics from the resampling launch that follows (Resampler.resample(want_stats=True): the ancestor kernel
    normalises the same log-weights anyway) and calls _book_weights itself."""
    if not defer and (result_flags['loss_required'] or result_flags['ess_required']):
        _book_weights(trace, result_flags, log_q_f, ops.weight_stats(log_w, want_w=result_flags['loss_required']))
    if result_flags['density_required']:
        trace['density'].append(density.unsqueeze(0))


def apg_objective(models, x, result_flags, num_sweeps, block, resampler):
    """Amortized Population Gibbs objective in the GMM problem (objectives.py:9-53).
    Besides the reference's trace keys, `log_z` (num_sweeps, B) carries the log-marginal estimate
    logsumexp_s log_w - log S of every step that appends an ESS (north_star: "log-marginal/ESS metric")."""
    trace = {'loss': [], 'ess': [], 'E_tau': [], 'E_mu': [], 'E_z': [], 'density': [], 'log_z': []}
    (enc_rws_eta, enc_apg_z, enc_apg_eta, generative) = models
    if x.dim() == 3:
        x = x.unsqueeze(0).expand(resampler.S, -1, -1, -1)       # stride-0 S axis; never materialised
    need_stats = bool(result_flags['loss_required'] or result_flags['ess_required'])
    # the part of Enc_apg_z that depends only on (x, weights) is the same for every particle and every sweep of this call:
    # tabulated once per instance, fetched by the z-step CTAs with one bulk copy (kernels.z_tables)
    ztab = None
    if (z_tables_per_objective and rng_in_kernels and getattr(enc_apg_z, 'fused', False) and hasattr(enc_apg_z, 'mlp')
            and enc_apg_z.can_fuse(False)):
        with torch.no_grad():                    # weights are constant within one objective call (training mode included)
            ztab = kernels.z_tables(x, enc_apg_z.mlp(), generative)
    evaluating = not result_flags['loss_required']
    etab = None                                  # the same for Enc_apg_eta's head factors over (point, label): kernels.eta_tables
    if z_tables_per_objective and evaluating and block == 'decomposed' and _eta_fused(enc_apg_eta) and num_sweeps > 1:
        with torch.no_grad():
            etab = kernels.eta_tables(x, enc_apg_eta.head_weights())

    def resample_and_book(tau, mu, z, log_w, log_q_f):
        """resample_variables + the W1 bookkeeping of the step from the SAME launch: the ancestor kernel's softmax over S
        is the normalisation the loss / ESS need (the reference computes it twice, objectives.py:93 and resampler.py:30)."""
        if not need_stats:
            return resample_variables(resampler, tau, mu, z, log_weights=log_w)
        (tau, mu, z), _, stats = resampler.resample([tau, mu, z], log_w, want_stats=True,
                                                    want_w=bool(result_flags['loss_required']))
        _book_weights(trace, result_flags, log_q_f, stats)
        return tau, mu, z

    if _compact_labels(block, result_flags, enc_apg_z, enc_apg_eta, generative):
        # Inside the sweeps the assignments never need their one-hot (S,B,N,K) fp32 form (the objective returns only
        # the trace): carry them as one byte per point.  The two per-sweep kernels take / return labels, and the z row
        # every resampling gather moves shrinks from 4*K*N to ~N bytes (720 -> 64 B for the GMM shapes; rows padded to
        # 16 bytes so that the gather moves 16-byte vectors).  The one-shot's sampler emits the labels with its draw.
        N = x.shape[-2]
        log_w, tau, mu, _, trace, z, log_q = _oneshot(enc_rws_eta, enc_apg_z, generative, x, trace, result_flags,
                                                      label_stride=(N + 15) // 16 * 16, defer=True, tables=ztab)
    else:
        log_w, tau, mu, z, trace, _, log_q = _oneshot(enc_rws_eta, enc_apg_z, generative, x, trace, result_flags, defer=True)
    compact = z.dtype == torch.uint8
    # gatherless sweeps (evaluation flags, decomposed block, label mode, in-kernel draws): see the loop below
    gatherless_now = bool(gatherless and evaluating and compact and block == 'decomposed' and ztab is not None and not fuse_eta_stats
                          and rng_in_kernels and _eta_fused(enc_apg_eta) and num_sweeps > 1)
    z_map = eta_map = None
    if gatherless_now:
        if need_stats:
            z_map, stats = resampler.ancestors32(log_w, want_stats=True, want_w=bool(result_flags['loss_required']))
            _book_weights(trace, result_flags, log_q, stats)
        else:
            z_map = resampler.ancestors32(log_w)
        eta_map = z_map                                  # the one-shot made (tau, mu, z) in the same order
    else:
        tau, mu, z = resample_and_book(tau, mu, z, log_w, log_q)
    pre = None         # (posterior parameters, centred statistics) of the current z, emitted by the previous z step
    for m in range(num_sweeps - 1):
        if block == 'decomposed':
            log_w_eta, tau, mu, trace, log_q_f = _update_eta(enc_apg_eta, generative, x, z, tau, mu, trace, result_flags,
                                                             defer=True, pre=pre, z_map=z_map, eta_map=eta_map,
                                                             tables=etab if compact else None)
            if gatherless_now:
                # Gatherless sweep: no latent is moved.  Every block update reads its inputs through ancestor maps --
                # (tau, mu) made by the eta step through A (this step's resampling), the labels made by the previous z step
                # through its resampling C composed with A -- and writes its outputs in place order; a latent is at most two
                # resampling steps old when it is read (ops.compose_ancestors).
                if need_stats:
                    A, stats = resampler.ancestors32(log_w_eta, want_stats=True, want_w=bool(result_flags['loss_required']))
                    _book_weights(trace, result_flags, log_q_f, stats)
                else:
                    A = resampler.ancestors32(log_w_eta)
                zA = ops.compose_ancestors(z_map, A) if z_map is not None else A
                log_w_z, z, trace, _ = _update_z(enc_apg_z, generative, x, tau, mu, z, trace, result_flags, tables=ztab,
                                                 src_ancestors=A, src_ancestors_z=zA)
                z_map = resampler.ancestors32(log_w_z)                  # C: the new labels sit in the z step's order
                eta_map = ops.compose_ancestors(A, z_map)               # (tau, mu) sit in the eta step's order: A then C
                continue
            if compact and defer_gather and evaluating and ztab is not None and not fuse_eta_stats:
                # Deferred gather: the z step only READS (tau, mu, z) of the resampled particles, so it takes them through
                # this resampling step's ancestors, and the gather after the z step composes the two index maps for the
                # latents that are still in the old order (tau, mu) -- one gather launch per sweep instead of two.
                if need_stats:
                    anc1, stats = resampler.ancestors32(log_w_eta, want_stats=True, want_w=bool(result_flags['loss_required']))
                    _book_weights(trace, result_flags, log_q_f, stats)
                else:
                    anc1 = resampler.ancestors32(log_w_eta)
                log_w_z, z_new, trace, _ = _update_z(enc_apg_z, generative, x, tau, mu, z, trace, result_flags, tables=ztab,
                                                     src_ancestors=anc1)
                anc2 = resampler.ancestors32(log_w_z)
                tau, mu, z = ops.gather_ancestors([tau, mu, z_new], anc2, first=[anc1, anc1, None])
                continue
            tau, mu, z = resample_and_book(tau, mu, z, log_w_eta, log_q_f)
            # the z step of every sweep but the last also emits the NEXT eta step's posterior parameters and statistics
            # (functions of the new labels and this particle's mu); they ride through the resampling gather with the
            # particle, so apg_update_eta needs no pass of its own over the labels
            emit = compact and fuse_eta_stats and m < num_sweeps - 2 and _can_emit_eta(enc_apg_eta, generative)
            log_w_z, z, trace, pre = _update_z(enc_apg_z, generative, x, tau, mu, z, trace, result_flags,
                                               eta_heads=enc_apg_eta.head_weights() if emit else None, tables=ztab)
            if pre is not None:
                outs = resampler.resample([tau, mu, z] + list(pre[0]) + [pre[1]], log_w_z)
                tau, mu, z, pre = outs[0], outs[1], outs[2], (tuple(outs[3:7]), outs[7])
            else:
                tau, mu, z = resample_variables(resampler, tau, mu, z, log_weights=log_w_z)
        elif block == 'joint':
            log_w, tau, mu, z, trace, log_q_f = _update_joint(enc_apg_z, enc_apg_eta, generative, x, z, tau, mu, trace,
                                                              result_flags, defer=True)
            tau, mu, z = resample_and_book(tau, mu, z, log_w, log_q_f)
        else:
            raise ValueError
    if result_flags['loss_required']:
        trace['loss'] = torch.cat(trace['loss'], 0)
    if result_flags['ess_required']:
        trace['ess'] = torch.cat(trace['ess'], 0)
        trace['log_z'] = torch.cat(trace['log_z'], 0)
    if result_flags['mode_required']:
        trace['E_tau'] = torch.cat(trace['E_tau'], 0)
        trace['E_mu'] = torch.cat(trace['E_mu'], 0)
        trace['E_z'] = torch.cat(trace['E_z'], 0)
    if result_flags['density_required']:
        trace['density'] = torch.cat(trace['density'], 0)
    return trace


rng_in_kernels = True          # eta step: Gamma / Normal variates drawn inside the weight kernel (else torch._standard_gamma / randn);
                               # z step: the uniforms of the categorical draw made inside the fused kernel (else torch.rand)
_SIDE_STREAMS = {}


def apg_objective_streams(models, x, result_flags, num_sweeps, block, resampler, parts=2):
    """apg_objective with the instance batch cut into `parts` contiguous slices that run on separate CUDA streams.
    Instances never interact (every op normalises / scans / gathers over S within one instance), so the slices
    are independent sweeps; on separate streams the latency-bound launches of one slice (ancestors, gather, eta weights,
    the small PyTorch glue) overlap the compute-bound z step of the other.  Capturable: inside a CUDA graph the slices
    become parallel branches.  Same trace as apg_objective (slices concatenated along the instance axis)."""
    B = x.shape[-3]
    parts = max(1, min(int(parts), B))
    if parts == 1:
        return apg_objective(models, x, result_flags, num_sweeps, block, resampler)
    main = torch.cuda.current_stream(x.device)
    key = (x.device, parts)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=x.device) for _ in range(parts)]
    bounds = [(i * B // parts, (i + 1) * B // parts) for i in range(parts)]
    outs = []
    for (lo, hi), st in zip(bounds, _SIDE_STREAMS[key]):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            xs = x.narrow(x.dim() - 3, lo, hi - lo)
            outs.append(apg_objective(models, xs, result_flags, num_sweeps, block, resampler))
    for st in _SIDE_STREAMS[key]:
        main.wait_stream(st)
    trace = {}
    for k, v0 in outs[0].items():
        if not torch.is_tensor(v0):
            trace[k] = v0
            continue
        vs = [o[k] for o in outs]
        for v in vs:
            v.record_stream(main)
        if k.startswith('loss'):                                   # means over the instances of a slice -> of the batch
            w = torch.tensor([hi - lo for lo, hi in bounds], dtype=v0.dtype, device=v0.device) / float(B)
            trace[k] = sum(wi * v for wi, v in zip(w, vs))
        elif k == 'density':
            trace[k] = torch.cat(vs, 2)                            # (sweeps, S, B)
        else:
            trace[k] = torch.cat(vs, 1)                            # ess / log_z (sweeps, B); E_* (sweeps, B, ...)
    return trace


train_on_kernels = True        # loss_required: graphs of log q through the kernels (train.py) instead of PyTorch encoders
compact_labels = True          # module switches (tests compare the representations / the two one-shot paths)
lean_oneshot = True
# z step emits the next eta step's posterior parameters + statistics (apgs_gmm_z_step_fused_labels_eta_f32; one launch
# less per sweep, bit-identical traces).  Measured on B200 at S=1000, B=1024 (profiles/r2_fused_eta_stats.md): the
# statistics are ~1000 instructions per thread wherever they run, so the z step grows by 0.25 ms against 0.32 ms for the
# standalone kernel, and the 156 extra bytes per particle cost the gather 0.07 ms -- no net gain, hence off by default.
fuse_eta_stats = False
gatherless = True              # evaluation sweeps: no gather at all -- every step reads through (composed) ancestor maps
defer_gather = True            # evaluation sweeps: the z step reads through the eta step's ancestors (one gather per sweep)
z_tables_per_objective = True  # evaluation sweeps: per-instance tables of the z step built once per apg_objective call


def _can_emit_eta(enc_apg_eta, generative):
    K, D = generative.prior_pi.numel(), generative_dims(enc_apg_eta)
    return K * D <= 8


def _compact_labels(block, result_flags, enc_apg_z, enc_apg_eta, generative):
    K, D = generative.prior_pi.numel(), generative_dims(enc_apg_eta)
    if result_flags['loss_required']:
        # training: the graphs of log q must take labels too -- FusedZStep (z) and LabelTableSum (eta) do
        if not (_train_z(enc_apg_z) and train_on_kernels and hasattr(enc_apg_eta, 'params_graph')):
            return False
    return bool(compact_labels and block == 'decomposed'
                and enc_apg_z.can_fuse(False) and _eta_fused(enc_apg_eta) and K <= 255
                and kernels.supported(K, D) & kernels.LABELS)


def _train_z(enc_apg_z):
    """training mode of Enc_apg_z on the kernels: parameters take part in autograd and the fused step is compiled for (K, D)"""
    return (train_on_kernels and needs_grad(enc_apg_z) and getattr(enc_apg_z, 'fused', False) and hasattr(enc_apg_z, 'mlp')
            and enc_apg_z.can_fuse(False))


def _sample_z_train(enc_z, generative, x, tau, mu, label_stride=0, tables=None):
    """sample-only fused z step with a graph: z (one-hot), q_probs, log q(z) summed over the points (S,B), labels or None"""
    from .train import FusedZStep
    z, q_probs, _, _, _, log_q_local, z_hot = FusedZStep.apply(*enc_z.mlp(), kernels._x3(x), tau, mu, None, generative,
                                                               int(label_stride), _zseed(tau.device), tables)
    if label_stride:
        return z_hot, q_probs, log_q_local.sum(-1), z
    return z, q_probs, log_q_local.sum(-1), None


def _eta_fused(enc_apg_eta):
    return enc_apg_eta.can_fuse() if hasattr(enc_apg_eta, 'can_fuse') else bool(getattr(enc_apg_eta, 'fused', False))


def generative_dims(enc_apg_eta):
    """data dimension D from the Enc_apg_eta heads (ob: Linear(K + D, D))."""
    return enc_apg_eta.ob[0].weight.shape[0]


def oneshot(enc_rws_eta, enc_rws_z, generative, x, trace, result_flags):
    """One-shot for eta and z, like a normal RWS (objectives.py:74-110)."""
    return _oneshot(enc_rws_eta, enc_rws_z, generative, x, trace, result_flags)[:5]


def _zseed(dev):
    """Philox key of an in-kernel z draw (None: the kernel reads torch.rand uniforms)"""
    return kernels.philox_seed(dev) if rng_in_kernels else None


def _oneshot(enc_rws_eta, enc_rws_z, generative, x, trace, result_flags, label_stride=0, defer=False, tables=None):
    """oneshot; with label_stride (fused sampler only) the draw of z also comes back as uint8 labels (6th value);
    the 7th value is log q of the draw (what a deferred _book_weights needs)."""
    post = enc_rws_eta.params(x, generative.prior_ng)
    tau, mu = _sample_eta(post)
    z_lab = None
    if label_stride and lean_oneshot and not result_flags['mode_required'] and not result_flags['loss_required']:
        # nothing downstream needs q_probs or the one-hot draw: the sampler returns labels plus log p(x, z | eta) and
        # log q(z) of its draw, and the eta terms come from the statistics weight kernel on EMPTY statistics (no point
        # assigned: its log-likelihoods vanish, its density / log_q_f are log p(eta) and log q(eta)).  Same terms as
        # kernels.logw_oneshot, summed as (ll + log p(z)) + log p(eta) instead of (ll + log p(eta)) + log p(z).
        (_, z_lab), _, wz = kernels.z_step_fused(x, tau, mu, enc_rws_z.mlp(), generative, want_q=False,
                                                 also_labels=label_stride, want_onehot=False, seed=_zseed(tau.device),
                                                 tables=tables if rng_in_kernels else None)
        S, B, K, D = tau.shape
        we = kernels.logw_eta_stats(torch.zeros(S, B, K, 1 + 2 * D, device=tau.device), mu, post, generative,
                                    tau, mu, tau, mu)
        log_p = wz.density + we.density
        log_q = we.log_q_f + wz.log_q_f
        log_w = log_p - log_q
        _book(trace, result_flags, log_w, log_q, log_p, defer)
        return log_w, tau, mu, None, trace, z_lab, log_q
    log_q_z = None
    if result_flags['loss_required'] and _train_z(enc_rws_z):
        z, q_probs, log_q_z, z_lab = _sample_z_train(enc_rws_z, generative, x, tau, mu, label_stride, tables=tables)
    elif label_stride:
        z, q_probs, z_lab = enc_rws_z.sample_fused(x, tau, mu, generative, label_stride=label_stride, seed=_zseed(tau.device))
    elif enc_rws_z.can_fuse(result_flags['loss_required']):
        z, q_probs = enc_rws_z.sample_fused(x, tau, mu, generative, seed=_zseed(tau.device))
    else:
        q_probs = enc_rws_z.probs(x, tau, mu)
        z = cat(q_probs, validate_args=False).sample()
    w = kernels.logw_oneshot(x, post, generative, tau, mu, q_probs, z)
    log_q = w.log_q_f
    if result_flags['loss_required'] and needs_grad(enc_rws_eta, enc_rws_z):
        if log_q_z is None:
            log_q_z = cat(q_probs, validate_args=False).log_prob(z).sum(-1)
        log_q = eta_log_q_torch(enc_rws_eta.params_torch(x, generative.prior_ng), tau, mu) + log_q_z
    _book(trace, result_flags, w.log_w, log_q, w.density, defer)
    if result_flags['mode_required']:
        trace['E_tau'].append((post[0] / post[1]).mean(0).unsqueeze(0))
        trace['E_mu'].append(post[2].mean(0).unsqueeze(0))
        trace['E_z'].append((q_probs / q_probs.sum(-1, keepdim=True)).mean(0).unsqueeze(0).detach())
    return w.log_w, tau, mu, z, trace, z_lab, log_q


def apg_update_joint(enc_apg_z, enc_apg_eta, generative, x, z_old, tau_old, mu_old, trace, result_flags):
    """Jointly update all the variables (objectives.py:112-158)."""
    return _update_joint(enc_apg_z, enc_apg_eta, generative, x, z_old, tau_old, mu_old, trace, result_flags)[:5]


def _update_joint(enc_apg_z, enc_apg_eta, generative, x, z_old, tau_old, mu_old, trace, result_flags, defer=False):
    post = enc_apg_eta.params(x, z_old, generative.prior_ng)
    tau, mu = _sample_eta(post)
    log_q_z = None
    if result_flags['loss_required'] and _train_z(enc_apg_z):
        z, q_probs, log_q_z, _ = _sample_z_train(enc_apg_z, generative, x, tau, mu)
    elif enc_apg_z.can_fuse(result_flags['loss_required']):
        z, q_probs = enc_apg_z.sample_fused(x, tau, mu, generative, seed=_zseed(tau.device))
    else:
        q_probs = enc_apg_z.probs(x, tau, mu)
        z = cat(q_probs, validate_args=False).sample()
    w = kernels.logw_joint(x, post, generative, tau, mu, tau_old, mu_old, q_probs, z, z_old)
    log_q_f = w.log_q_f
    if result_flags['loss_required'] and needs_grad(enc_apg_eta, enc_apg_z):
        if log_q_z is None:
            log_q_z = cat(q_probs, validate_args=False).log_prob(z).sum(-1)
        if train_on_kernels and hasattr(enc_apg_eta, 'log_q_graph'):
            log_q_f = enc_apg_eta.log_q_graph(x, z_old, generative.prior_ng, tau, mu) + log_q_z
        else:
            log_q_f = eta_log_q_torch(enc_apg_eta.params_torch(x, z_old, generative.prior_ng), tau, mu) + log_q_z
    _book(trace, result_flags, w.log_w, log_q_f, w.density, defer)
    if result_flags['mode_required']:
        # the reference's branch (objectives.py:149-155) reads an undefined name; emit what it intends
        trace['E_tau'].append((post[0] / post[1]).mean(0).unsqueeze(0))
        trace['E_mu'].append(post[2].mean(0).unsqueeze(0))
        trace['E_z'].append(z.unsqueeze(0))
    return w.log_w, tau, mu, z, trace, log_q_f


def apg_update_eta(enc_apg_eta, generative, x, z, tau_old, mu_old, trace, result_flags):
    """Given local variable z, update global variables eta := {mu, tau} (objectives.py:161-195)."""
    return _update_eta(enc_apg_eta, generative, x, z, tau_old, mu_old, trace, result_flags)[:4]


def _update_eta(enc_apg_eta, generative, x, z, tau_old, mu_old, trace, result_flags, defer=False, pre=None, z_map=None,
                eta_map=None, tables=None):
    """z_map / eta_map (S,B) int32 (gatherless sweeps): z and (tau_old, mu_old) are still in the order of the steps that made
    them; particle (s,b) reads them from rows (z_map[s,b], b) / (eta_map[s,b], b)."""
    if _eta_fused(enc_apg_eta):
        if pre is not None:
            post, stats = pre            # emitted by the z step that drew these labels, centred at its mu == mu_old
        elif z_map is not None or eta_map is not None or tables is not None:
            *post, stats = enc_apg_eta.params(x, z, generative.prior_ng, center=mu_old, z_ancestors=z_map, center_ancestors=eta_map,
                                              tables=tables)
            post = tuple(post)
        else:
            # one pass over z: posterior parameters + the centred one-hot statistics both log-likelihoods are functions of
            *post, stats = enc_apg_eta.params(x, z, generative.prior_ng, center=mu_old)
            post = tuple(post)
        # the draw of _sample_eta with its elementwise tail fused into the weight kernel: same variates, same IEEE
        # operations (tau = max(g / beta, tiny), mu = e / sqrt(nu tau) + q_mu), seven launches less per step
        if rng_in_kernels:
            # the variates themselves made in the kernel (Philox keyed by two words from torch's generator): no
            # _standard_gamma / randn launches, no variate tensors -- distribution-equal to the torch draws below
            w, tau, mu = kernels.draw_logw_eta_stats_rng(stats, mu_old, post, generative, kernels.philox_seed(mu_old.device),
                                                         tau_old, mu_old, old_ancestors=eta_map)
        else:
            g_std = torch._standard_gamma(post[0])
            e_std = torch.randn_like(post[2])
            w, tau, mu = kernels.draw_logw_eta_stats(stats, mu_old, post, generative, g_std, e_std, tau_old, mu_old)
    else:
        post = enc_apg_eta.params(x, z, generative.prior_ng)
        tau, mu = _sample_eta(post)
        w = kernels.logw_eta(x, z, post, generative, tau, mu, tau_old, mu_old)
    log_q_f = w.log_q_f
    if result_flags['loss_required'] and needs_grad(enc_apg_eta):
        if train_on_kernels and hasattr(enc_apg_eta, 'log_q_graph'):
            log_q_f = enc_apg_eta.log_q_graph(x, z, generative.prior_ng, tau, mu)
        else:
            log_q_f = eta_log_q_torch(enc_apg_eta.params_torch(x, z, generative.prior_ng), tau, mu)
    _book(trace, result_flags, w.log_w, log_q_f, w.density, defer)
    if result_flags['mode_required']:
        trace['E_tau'].append((post[0] / post[1]).mean(0).unsqueeze(0))
        trace['E_mu'].append(post[2].mean(0).unsqueeze(0))
    return w.log_w, tau, mu, trace, log_q_f


def apg_update_z(enc_apg_z, generative, x, tau, mu, z_old, trace, result_flags):
    """Given the current samples of eta, update the assignments z (objectives.py:197-227)."""
    return _update_z(enc_apg_z, generative, x, tau, mu, z_old, trace, result_flags)[:3]


def _update_z(enc_apg_z, generative, x, tau, mu, z_old, trace, result_flags, eta_heads=None, tables=None, src_ancestors=None,
              src_ancestors_z=None):
    """apg_update_z; with eta_heads (label mode, evaluation flags) the fused kernel also emits the next eta step's
    (posterior parameters, statistics) as the 4th value (else None)."""
    need_local = bool(result_flags['loss_required'])
    pre = None
    if need_local and _train_z(enc_apg_z):
        # N2: the fused forward launch + a backward kernel for the MLP gradients (train.FusedZStep)
        from .train import FusedZStep
        z, q_probs, log_w, density, log_w_local, log_q_local, _ = FusedZStep.apply(*enc_apg_z.mlp(), kernels._x3(x), tau, mu,
                                                                                  z_old, generative, 0, _zseed(tau.device), tables)
        w_local, _ = ops.local_weight_stats(log_w_local, want_w=True, want_sum=False)
        loss = (w_local * (-log_q_local)).sum(0).sum(-1).mean()
        trace['loss'][-1] = trace['loss'][-1] + loss.unsqueeze(0)
        if result_flags['mode_required']:
            trace['E_z'].append((q_probs / q_probs.sum(-1, keepdim=True)).mean(0).unsqueeze(0))
        if result_flags['density_required']:
            trace['density'][-1] = trace['density'][-1] + density.unsqueeze(0)
        return log_w, z, trace, None
    if eta_heads is not None and not (result_flags['mode_required'] or need_local):
        z, q_probs, w, post, stats = kernels.z_step_fused(x, tau, mu, enc_apg_z.mlp(), generative, z_old=z_old,
                                                          eta_heads=eta_heads, seed=_zseed(tau.device),
                                                          tables=tables if rng_in_kernels else None)
        pre = (post, stats)
    elif enc_apg_z.can_fuse(result_flags['loss_required']):
        # N1: MLP + softmax + draw + weights in one kernel; q_probs only materialised when asked for
        z, q_probs, w = kernels.z_step_fused(x, tau, mu, enc_apg_z.mlp(), generative, z_old=z_old,
                                             want_q=bool(result_flags['mode_required']), local=need_local,
                                             seed=_zseed(tau.device), tables=tables if rng_in_kernels else None,
                                             src_ancestors=src_ancestors, src_ancestors_z=src_ancestors_z)
    else:
        q_probs = enc_apg_z.probs(x, tau, mu)
        z = cat(q_probs, validate_args=False).sample()
        w = kernels.logw_z(x, tau, mu, q_probs, generative, z, z_old, local=need_local)
    if result_flags['loss_required']:
        w_local, _ = ops.local_weight_stats(w.log_w_local, want_w=True, want_sum=False)
        log_q_local = w.log_q_local
        if needs_grad(enc_apg_z):
            log_q_local = cat(q_probs, validate_args=False).log_prob(z)       # q_probs from the PyTorch MLP: has a graph
        loss = (w_local * (-log_q_local)).sum(0).sum(-1).mean()
        trace['loss'][-1] = trace['loss'][-1] + loss.unsqueeze(0)
    if result_flags['mode_required']:
        trace['E_z'].append((q_probs / q_probs.sum(-1, keepdim=True)).mean(0).unsqueeze(0))
    if result_flags['density_required']:
        trace['density'][-1] = trace['density'][-1] + w.density.unsqueeze(0)
    return w.log_w, z, trace, pre


def resample_variables(resampler, tau, mu, z, log_weights):
    """objectives.py:229-234: one ancestor draw, one gather launch for the three latents."""
    tau, mu, z = resampler.resample([tau, mu, z], log_weights)
    return tau, mu, z


# ------------------------------------------------------------------ baselines around the path
def _prior_as_post(generative, S, B):
    return tuple(p.unsqueeze(0).unsqueeze(0).expand(S, B, -1, -1).contiguous() for p in generative.prior_ng)


def posterior_z(x, tau, mu, prior_pi):
    """log posterior responsibilities (S,B,N,K) given eta (apgs/gmm/kls_gmm.py:121-133)."""
    sigma = 1. / tau.sqrt()
    xe = x.unsqueeze(2)                                                  # S,B,1,N,D (broadcast over K)
    log_gammas = Normal(mu.unsqueeze(-2), sigma.unsqueeze(-2), validate_args=False).log_prob(xe).sum(-1).transpose(-1, -2)
    return torch.log_softmax(log_gammas + prior_pi.log(), dim=-1)


def gibbs_objective(models, x, result_flags, num_sweeps):
    """The Gibbs sampler baseline (objectives.py:236-247)."""
    trace = {'density': []}
    (enc_rws_eta, enc_rws_z, _, generative) = models
    if x.dim() == 3:
        raise ValueError("gibbs_objective takes the S-replicated (S,B,N,D) data, as the reference")
    _, tau, mu, z, trace = oneshot(enc_rws_eta, enc_rws_z, generative, x, trace, result_flags)
    for m in range(num_sweeps - 1):
        tau, mu, z, trace = gibbs_sweep(generative, x, z, trace)
    if result_flags['density_required']:
        trace['density'] = torch.cat(trace['density'], 0)
    return trace


def gibbs_sweep(generative, x, z, trace):
    """Exact conjugate updates (objectives.py:249-274): eta | z through the G1 kernel on the plain statistics,
    z | eta in PyTorch, the log joint through the G4 density kernel."""
    post = kernels.nss_posterior(x.contiguous(), z, generative)          # posterior_eta(x, z): gamma = z, x' = x
    tau, mu = _sample_eta(post)
    logits = posterior_z(x, tau, mu, generative.prior_pi)
    q_probs = logits.exp()
    z = cat(logits=logits, validate_args=False).sample()
    w = kernels.logw_oneshot(x, post, generative, tau, mu, q_probs, z)   # density = ll + log p(eta) + log p(z)
    trace['density'].append(w.density.unsqueeze(0))
    return tau, mu, z, trace


def bpg_objective(models, x, result_flags, num_sweeps, resampler):
    """APG whose eta block proposes from the prior (objectives.py:291-306); the reference runs it with S * bpg_factor
    particles (evaluation.py:27,41-42), which is what the split-S ancestor pipeline is for."""
    trace = {'density': []}
    (enc_rws_eta, enc_apg_z, enc_apg_eta, generative) = models
    if x.dim() == 3:
        x = x.unsqueeze(0).expand(resampler.S, -1, -1, -1)
    log_w, tau, mu, z, trace = oneshot(enc_rws_eta, enc_apg_z, generative, x, trace, result_flags)
    tau, mu, z = resample_variables(resampler, tau, mu, z, log_weights=log_w)
    for m in range(num_sweeps - 1):
        log_w_eta, tau, mu, trace = bpg_update_eta(generative, x, z, tau, mu, trace)
        tau, mu, z = resample_variables(resampler, tau, mu, z, log_weights=log_w_eta)
        log_w_z, z, trace = apg_update_z(enc_apg_z, generative, x, tau, mu, z, trace, result_flags)
        tau, mu, z = resample_variables(resampler, tau, mu, z, log_weights=log_w_z)
    trace['density'] = torch.cat(trace['density'], 0)
    return trace


def bpg_update_eta(generative, x, z, tau_old, mu_old, trace):
    """eta proposed from the prior: weight = likelihood ratio (objectives.py:308-320).  One G2 launch with the prior as
    the proposal: log p(eta) - log q(eta) cancels bit for bit, leaving ll_f - ll_b; density = log p(eta_new)."""
    S, B = z.shape[0], z.shape[1]
    post = _prior_as_post(generative, S, B)
    tau, mu = _sample_eta(post)
    w = kernels.logw_eta(x, z, post, generative, tau, mu, tau_old, mu_old)
    trace['density'].append(w.density.unsqueeze(0))
    return w.log_w, tau, mu, trace
