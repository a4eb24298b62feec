"""APG sweep for bouncing MNIST: drop-in for the hot-path functions of apgs/bmnist/objectives.py
(resample_variables :5-9, apg_objective :11-58, propose_one_movement :60-107, oneshot :109-165,
apg_where :167-225, apg_what :228-257).

The MLPs stay in PyTorch; the template matching (grouped conv2d + softmax), crop, paste,
paste+sum+clamp+log-likelihood, every Normal log-density, the step weights, softmax/ESS and the per-frame
resampling run in the sm_100a kernels.  Reference behaviours that change numbers are kept:
  * the one-shot counts the t = 0 terms of log p / log q(z_where) twice (objectives.py:124-139);
  * apg_where resamples a local z_what at every frame but returns only z_where, so apg_what scores the new
    z_what against the caller's un-resampled z_what (objectives.py:212,225,45);
  * frame t+1 is proposed from the un-resampled z_where_t (objectives.py:189).
Gradients: the kernels are forward-only.  The losses are  sum_s w_s.detach() (-log q_s)  and
sum_s w_s.detach() (-ll_s);  when `loss_required` is set and encoder / decoder parameters require grad, the FORWARD
log q (from the proposal parameters, which are PyTorch MLP outputs anyway) and ll (paste + Bernoulli likelihood as
gathers, affine_transformer.reconstruct_ll_torch) are rebuilt with PyTorch ops on the kernels' draws and multiplied by
the detached kernel weights.  Everything else (weights, ESS, densities, resampling) stays in the
kernels.
"""
import math

import torch
from torch.distributions.normal import Normal

from .. import ops
from ..rng import normal_sample
from . import kernels
from .models import needs_grad


def resample_variables(resampler, z_where, z_what, log_weights):
    z_where, z_what = resampler.resample([z_where, z_what], log_weights)
    return z_where, z_what


def _frames5(frames, S):
    return frames if frames.dim() == 5 else frames.unsqueeze(0).expand(S, -1, -1, -1, -1)


def apg_objective(models, AT, frames, K, result_flags, num_sweeps, resampler, mnist_mean):
    """Amortized Population Gibbs objective in the bouncing MNIST problem (objectives.py:11-58)."""
    trace = {'loss_phi': [], 'loss_theta': [], 'ess': [], 'E_where': [], 'E_what': [], 'E_recon': [], 'density': []}
    frames = _frames5(frames, resampler.S)
    (enc_coor, dec_coor, enc_digit, dec_digit) = models
    log_w, z_where, z_what, trace = oneshot(enc_coor, dec_coor, enc_digit, dec_digit, AT, frames, mnist_mean, trace, result_flags)
    z_where, z_what = resample_variables(resampler, z_where, z_what, log_weights=log_w)
    for m in range(num_sweeps - 1):
        z_where, trace = apg_where(enc_coor, dec_coor, dec_digit, AT, resampler, frames, z_what, z_where, trace, result_flags)
        log_w, z_what, trace = apg_what(enc_digit, dec_digit, AT, frames, z_where, z_what, trace, result_flags)
        z_where, z_what = resample_variables(resampler, z_where, z_what, log_weights=log_w)
    if result_flags['loss_required']:
        trace['loss_phi'] = torch.cat(trace['loss_phi'], 0)
        trace['loss_theta'] = torch.cat(trace['loss_theta'], 0)
    if result_flags['ess_required']:
        trace['ess'] = torch.cat(trace['ess'], 0)
        trace['log_z'] = torch.cat(trace['log_z'], 0)
    if result_flags['mode_required']:
        trace['E_where'] = torch.cat(trace['E_where'], 0)
        trace['E_what'] = torch.cat(trace['E_what'], 0)
        trace['E_recon'] = torch.cat(trace['E_recon'], 0).cpu()
    if result_flags['density_required']:
        trace['density'] = torch.cat(trace['density'], 0)
    return trace


def propose_one_movement(enc_coor, dec_coor, AT, frame, template, z_where_t_1, z_where_old_t, z_where_old_t_1):
    """objectives.py:60-107: K sequential template matches; returns the reference's tuple."""
    return _propose(enc_coor, dec_coor, AT, frame, template, z_where_t_1, z_where_old_t, z_where_old_t_1)[0]


def _propose(enc_coor, dec_coor, AT, frame, template, z_where_t_1, z_where_old_t, z_where_old_t_1, train=False):
    """-> (the reference's tuple, (q_mean, q_std)) -- the proposal parameters feed the caller's weight kernel.
    train: log_q_f is rebuilt from (q_mean, q_std) with PyTorch ops so that it carries the graph of Enc_coor (N2).  The
    templates reach this function detached (models.py:137-138, `digit_mean.detach()`; the one-shot's are constants), so the
    matched score map is a constant of the graph in the reference too and stays in the kernel."""
    FP = frame.shape[-1]
    S, B, K, DP, _ = template.shape
    zs, means, stds = [], [], []
    frame_left = frame
    packed = enc_coor.wants_packed(FP, DP)
    # the residual frame_left - digit_to_frame(template_k, z_where_k) (:87-89) is formed inside the next match (two digits deep;
    # nearest-mode transformer at the compiled shape), otherwise by the paste-subtract kernel
    in_match = getattr(AT, "mode", "nearest") == "nearest" and kernels.template_match_supports_residual(FP, DP)
    prev = []
    for k in range(K):
        template_k = template[:, :, k, :, :]
        # conv2d(groups=S*B) + softmax (:72-74); written as the packed operand of Enc_coor's tensor-core layer when it takes one
        conved_k = kernels.template_match(frame_left, template_k, packed=packed, prev=prev, lin_fp=AT.lin_fp if prev else None)
        q_mean, q_std = enc_coor.params(conved_k)
        z_where_k = normal_sample(q_mean.detach(), q_std.detach())
        zs.append(z_where_k.unsqueeze(2)); means.append(q_mean.unsqueeze(2)); stds.append(q_std.unsqueeze(2))
        if k + 1 < K:                                        # frame_left - digit_to_frame(template_k, z_where_k) (:87-89)
            if in_match and len(prev) < 2:
                prev = prev + [(template_k, z_where_k)]
            else:
                for d, z in prev + [(template_k, z_where_k)]:
                    frame_left = AT.subtract_digit(frame_left, d, z)
                prev = []
    z_where = torch.cat(zs, 2)
    q_mean, q_std = torch.cat(means, 2), torch.cat(stds, 2)
    log_q_graph = Normal(q_mean, q_std, validate_args=False).log_prob(z_where).sum(-1).sum(-1) if train else None
    q_mean, q_std = q_mean.detach(), q_std.detach()
    # every Normal log-density of this frame (all K digits, forward and backward) in one kernel (B4a)
    E_where = q_mean
    if z_where_old_t is not None:
        _, log_q_f, log_p_f, log_q_b, log_p_b = kernels.logw_where(q_mean, q_std, z_where, z_where_t_1, z_where_old_t,
                                                                   z_where_old_t_1, dec_coor, want_w=False, want_b=True)
        return (log_p_f, log_q_f if log_q_graph is None else log_q_graph, log_p_b, log_q_b, z_where, E_where), (q_mean, q_std)
    _, log_q_f, log_p_f = kernels.logw_where(q_mean, q_std, z_where, z_where_t_1, None, None, dec_coor, want_w=False)
    return (log_p_f, log_q_f if log_q_graph is None else log_q_graph, z_where, E_where), (q_mean, q_std)


def oneshot(enc_coor, dec_coor, enc_digit, dec_digit, AT, frames, digit, trace, result_flags):
    """objectives.py:109-165."""
    T = frames.shape[2]
    S, B, K, DP, DP = digit.shape
    train = bool(result_flags['loss_required']) and needs_grad(enc_coor, enc_digit, dec_digit)
    train_coor = bool(result_flags['loss_required']) and needs_grad(enc_coor)
    z_where, E_where = [], []
    z_where_t = None
    for t in range(T):
        (log_p_where_t, log_q_where_t, z_where_t, E_where_t), _ = _propose(
            enc_coor, dec_coor, AT, frames[:, :, t, :, :], digit, z_where_t, None, None, train=train_coor)
        if t == 0:
            log_p_where, log_q_where = log_p_where_t, log_q_where_t
        log_q_where = log_q_where + log_q_where_t          # t = 0 is counted twice, as in the reference
        log_p_where = log_p_where + log_p_where_t
        z_where.append(z_where_t.unsqueeze(2))
        E_where.append(E_where_t.detach().unsqueeze(2))
    z_where = torch.cat(z_where, 2)
    E_where = torch.cat(E_where, 2)
    cropped = AT.frame_to_digit(frames=frames, z_where=z_where).view(S, B, T, K, DP * DP)
    q_mu, q_std = enc_digit.params(cropped)
    z_what = normal_sample(q_mu, q_std)
    ll, recon = AT.reconstruct_ll(dec_digit.templates(z_what), z_where, frames, want_recon=bool(result_flags['mode_required']))
    log_w, log_q, log_p, ll_sum = kernels.logw_what(q_mu, q_std, z_what, None, ll, None, extra_log_p=log_p_where,
                                                    extra_log_q=log_q_where)
    if train:                                                # objectives.py:142-146 with a graph
        log_q = log_q_where + Normal(q_mu, q_std, validate_args=False).log_prob(z_what).sum(-1).sum(-1)
        ll_sum = AT.reconstruct_ll_torch(dec_digit.templates(z_what), z_where, frames).sum(-1)
    if result_flags['loss_required'] or result_flags['ess_required']:
        w, ess, lse = ops.weight_stats(log_w, want_w=result_flags['loss_required'])
        if result_flags['loss_required']:
            trace['loss_phi'].append((w * (-log_q)).sum(0).mean().unsqueeze(0))
            trace['loss_theta'].append((w * (-ll_sum)).sum(0).mean().unsqueeze(0))
        if result_flags['ess_required']:
            trace['ess'].append(ess.unsqueeze(0))
            trace.setdefault('log_z', []).append((lse - math.log(log_w.shape[0])).unsqueeze(0))     # log-marginal estimate (W1)
    if result_flags['mode_required']:
        trace['E_where'].append(E_where.mean(0).unsqueeze(0))
        trace['E_what'].append(q_mu.detach().mean(0).unsqueeze(0))
        trace['E_recon'].append(recon.mean(0).unsqueeze(0))
    if result_flags['density_required']:
        trace['density'].append(log_p.unsqueeze(0))
    return log_w, z_where, z_what, trace


def apg_where(enc_coor, dec_coor, dec_digit, AT, resampler, frames, z_what, z_where_old, trace, result_flags):
    """objectives.py:167-225: per-frame propose -> reweight -> resample of the growing z_where[:, :, :t+1]."""
    T = frames.shape[2]
    train_phi = bool(result_flags['loss_required']) and needs_grad(enc_coor)
    train_theta = bool(result_flags['loss_required']) and needs_grad(dec_digit)
    template = dec_digit.templates(z_what).detach()          # models.py:137-138
    E_where, LOSS_phi, LOSS_theta = [], [], []
    log_prior = 0.0
    z_where_t = None
    z_where = None
    for t in range(T):
        frame_t = frames[:, :, t, :, :]
        z_old_t = z_where_old[:, :, t, :, :]
        z_old_t_1 = None if t == 0 else z_where_old[:, :, t - 1, :, :]
        z_prev = z_where_t
        (log_p_f, log_q_f, log_p_b, log_q_b, z_where_t, E_where_t), (q_mean, q_std) = _propose(
            enc_coor, dec_coor, AT, frame_t, template, z_where_t, z_old_t, z_old_t_1, train=train_phi)
        if result_flags['density_required']:
            log_prior = log_prior + log_p_f
        if result_flags['mode_required']:
            E_where.append(E_where_t.detach().unsqueeze(2))
        digit = dec_digit.templates(z_what)
        ll_f, _ = AT.reconstruct_ll(digit, z_where_t.unsqueeze(2), frames[:, :, t:t + 1])
        ll_f_loss = AT.reconstruct_ll_torch(digit, z_where_t.unsqueeze(2), frames[:, :, t:t + 1]) if train_theta else ll_f
        ll_b, _ = AT.reconstruct_ll(digit, z_old_t.unsqueeze(2), frames[:, :, t:t + 1])
        # objectives.py:198-206 in one kernel: (log_p_f - log_q_f) - (log_p_b - log_q_b) + ll_f - ll_b
        log_w, _, _ = kernels.logw_where(q_mean, q_std, z_where_t, z_prev, z_old_t, z_old_t_1, dec_coor,
                                         ll_f=ll_f.squeeze(-1), ll_b=ll_b.squeeze(-1))
        z_where = z_where_t.unsqueeze(2) if t == 0 else torch.cat((z_where, z_where_t.unsqueeze(2)), 2)
        if result_flags['loss_required']:
            w, _, _ = ops.weight_stats(log_w)
            LOSS_phi.append((w * (-log_q_f)).sum(0).mean().unsqueeze(-1))
            LOSS_theta.append((w * (-ll_f_loss.squeeze(-1))).sum(0).mean().unsqueeze(-1))
        z_where, z_what = resample_variables(resampler, z_where, z_what, log_weights=log_w)
    if result_flags['mode_required']:
        trace['E_where'].append(torch.cat(E_where, 2).mean(0).unsqueeze(0))
    if result_flags['loss_required']:
        trace['loss_phi'].append(torch.cat(LOSS_phi, -1).sum(-1).unsqueeze(0))
        trace['loss_theta'].append(torch.cat(LOSS_theta, -1).sum(-1).unsqueeze(0))
    if result_flags['density_required']:
        trace['density'].append(log_prior.unsqueeze(0))
    return z_where, trace


def apg_what(enc_digit, dec_digit, AT, frames, z_where, z_what_old, trace, result_flags):
    """objectives.py:228-257."""
    S, B, T, K, _ = z_where.shape
    cropped = AT.frame_to_digit(frames=frames, z_where=z_where)
    DP = cropped.shape[-1]
    q_mu, q_std = enc_digit.params(cropped.view(S, B, T, K, int(DP * DP)))
    z_what = normal_sample(q_mu, q_std)
    want_recon = bool(result_flags['mode_required'])
    ll_f, recon = AT.reconstruct_ll(dec_digit.templates(z_what), z_where, frames, want_recon=want_recon)
    ll_b, _ = AT.reconstruct_ll(dec_digit.templates(z_what_old), z_where, frames)
    log_w, log_q_f, dens, ll_sum = kernels.logw_what(q_mu, q_std, z_what, z_what_old, ll_f, ll_b)
    if result_flags['loss_required']:
        if needs_grad(enc_digit):                            # objectives.py:236 with a graph
            log_q_f = Normal(q_mu, q_std, validate_args=False).log_prob(z_what).sum(-1).sum(-1)
        if needs_grad(dec_digit):
            ll_sum = AT.reconstruct_ll_torch(dec_digit.templates(z_what), z_where, frames).sum(-1)
    if result_flags['loss_required'] or result_flags['ess_required']:
        w, ess, lse = ops.weight_stats(log_w, want_w=result_flags['loss_required'])
        if result_flags['loss_required']:
            trace['loss_phi'][-1] = trace['loss_phi'][-1] + (w * (-log_q_f)).sum(0).mean().unsqueeze(0)
            trace['loss_theta'][-1] = trace['loss_theta'][-1] + (w * (-ll_sum)).sum(0).mean().unsqueeze(0)
        if result_flags['ess_required']:
            trace['ess'].append(ess.unsqueeze(0))
            trace.setdefault('log_z', []).append((lse - math.log(log_w.shape[0])).unsqueeze(0))     # log-marginal estimate (W1)
    if result_flags['mode_required']:
        trace['E_what'].append(q_mu.detach().mean(0).unsqueeze(0))
        trace['E_recon'].append(recon.mean(0).unsqueeze(0))
    if result_flags['density_required']:
        trace['density'][-1] = trace['density'][-1] + dens.unsqueeze(0)
    return log_w, z_what, trace


# ------------------------------------------------------------------ baseline around the path
def bpg_objective(models, AT, frames, result_flags, num_sweeps, resampler, mnist_mean):
    """APG whose z_what block proposes from the prior (apgs/bmnist/objectives.py:272-288)."""
    trace = {'density': []}
    frames = _frames5(frames, resampler.S)
    (enc_coor, dec_coor, enc_digit, dec_digit) = models
    log_w, z_where, z_what, trace = oneshot(enc_coor, dec_coor, enc_digit, dec_digit, AT, frames, mnist_mean, trace, result_flags)
    z_where, z_what = resample_variables(resampler, z_where, z_what, log_weights=log_w)
    for m in range(num_sweeps - 1):
        z_where, trace = apg_where(enc_coor, dec_coor, dec_digit, AT, resampler, frames, z_what, z_where, trace, result_flags)
        log_w, z_what, trace = bpg_what(dec_digit, AT, frames, z_where, z_what, trace)
        z_where, z_what = resample_variables(resampler, z_where, z_what, log_weights=log_w)
    trace['density'] = torch.cat(trace['density'], 0)
    return trace


def bpg_what(dec_digit, AT, frames, z_where, z_what_old, trace):
    """z_what proposed from its prior: weight = likelihood ratio (apgs/bmnist/objectives.py:290-302).  The B4 `what`
    kernel with the prior as the proposal (log p - log q cancels term by term); density += ll_f + log p(z_what_new)."""
    S, B, T, K, _ = z_where.shape
    shape = (S, B, K, z_what_old.shape[-1])
    q_mu = dec_digit.prior_mu.expand(shape).contiguous()
    q_std = dec_digit.prior_std.expand(shape).contiguous()
    z_what = normal_sample(q_mu, q_std)
    ll_f, _ = AT.reconstruct_ll(dec_digit.templates(z_what), z_where, frames)
    ll_b, _ = AT.reconstruct_ll(dec_digit.templates(z_what_old), z_where, frames)
    log_w, _, dens, _ = kernels.logw_what(q_mu, q_std, z_what, z_what_old, ll_f, ll_b)
    trace['density'][-1] = trace['density'][-1] + dens.unsqueeze(0)
    return log_w, z_what, trace
