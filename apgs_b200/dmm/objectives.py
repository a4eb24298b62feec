"""APG sweep for the DMM ring-mixture task: drop-in for the hot-path functions of apgs/dmm/objectives.py
(apg_objective :9-56, oneshot :80-114, apg_update_mu :116-151, apg_update_local :153-188,
resample_variables :190-195).

Each block update evaluates its encoder once (the reference re-runs it with identical inputs for the backward
kernel, and the decoder MLP twice although only mu moves in the mu step); log-densities, likelihood, the N-sums
and the f - b difference are one kernel (csrc/dmm.cu).
Gradients: the kernels are forward-only.  The APG losses are  loss_phi = sum_s w_s.detach() (-log q_s)
and  loss_theta = sum_s w_s.detach() (-ll_s), so autograd only needs the graphs of the FORWARD proposal log-density and of
the forward log-likelihood.  When `loss_required` is set and the encoder (decoder) parameters require grad, those two
terms are rebuilt with PyTorch ops on the kernels' draws and multiplied by the detached kernel weights -- the reference's
train() works unchanged and gets the reference's gradients, while weights, ESS, densities and resampling still come
from the kernels.
"""
import math

import torch
from torch.distributions.beta import Beta
from torch.distributions.normal import Normal
from torch.distributions.one_hot_categorical import OneHotCategorical as cat

from .. import ops
from ..rng import normal_sample
from . import kernels
from .models import needs_grad


def _book(trace, result_flags, log_w, log_q, ll, density):
    if result_flags['loss_required'] or result_flags['ess_required']:
        w, ess, lse = ops.weight_stats(log_w, want_w=result_flags['loss_required'])
        if result_flags['loss_required']:
            trace['loss_phi'].append((w * (-log_q)).sum(0).mean().unsqueeze(0))
            trace['loss_theta'].append((w * (-ll)).sum(0).mean().unsqueeze(0))
        if result_flags['ess_required']:
            trace['ess'].append(ess.unsqueeze(0))
            # log-marginal estimate of the step, logsumexp_s log_w - log S: the kernel's lse output
            trace.setdefault('log_z', []).append((lse - math.log(log_w.shape[0])).unsqueeze(0))
    if result_flags['density_required']:
        trace['density'].append(density.unsqueeze(0))


def apg_objective(models, x, K, result_flags, num_sweeps, resampler):
    """Amortized Population Gibbs objective in the DMM problem (objectives.py:9-56)."""
    trace = {'loss_phi': [], 'loss_theta': [], 'ess': [], 'E_mu': [], 'E_z': [], 'E_recon': [], 'density': []}
    (enc_rws_mu, enc_apg_local, enc_apg_mu, dec) = models
    if x.dim() == 3:
        x = x.unsqueeze(0).expand(resampler.S, -1, -1, -1)
    # the decoder output h2 is a function of beta alone, so it travels with beta through the resampling steps
    # (one more 8-byte-per-point latent in the same gather launch) instead of being re-evaluated from the gathered
    # beta: one decoder evaluation per sweep instead of three
    tables = None
    if getattr(enc_apg_mu, 'fused', False) and num_sweeps > 1:
        with torch.no_grad():
            tables = enc_apg_mu.label_tables(x[0] if x.dim() == 4 else x)      # data + weights only: once per objective
    log_w, mu, z, beta, h2, trace = _oneshot(enc_rws_mu, enc_apg_local, dec, x, K, trace, result_flags)
    mu, z, beta, h2 = resampler.resample([mu, z, beta, h2], log_w)
    for m in range(num_sweeps - 1):
        log_w_mu, mu, trace = apg_update_mu(enc_apg_mu, dec, x, z, beta, mu, K, trace, result_flags, h2=h2, tables=tables)
        mu, z, beta, h2 = resampler.resample([mu, z, beta, h2], log_w_mu)
        log_w_z, z, beta, h2, trace = _update_local(enc_apg_local, dec, x, mu, z, beta, K, trace, result_flags, h2_old=h2)
        mu, z, beta, h2 = resampler.resample([mu, z, beta, h2], log_w_z)
    if result_flags['loss_required']:
        trace['loss_phi'] = torch.cat(trace['loss_phi'], 0)
        trace['loss_theta'] = torch.cat(trace['loss_theta'], 0)
    if result_flags['ess_required']:
        trace['ess'] = torch.cat(trace['ess'], 0)
        trace['log_z'] = torch.cat(trace['log_z'], 0)
    if result_flags['mode_required']:
        trace['E_mu'] = torch.cat(trace['E_mu'], 0)
        trace['E_z'] = torch.cat(trace['E_z'], 0)
        trace['E_recon'] = torch.cat(trace['E_recon'], 0)
    if result_flags['density_required']:
        trace['density'] = torch.cat(trace['density'], 0)
    return trace


rng_in_kernels = True          # evaluation sweeps: the categorical draw's uniforms and beta ~ Beta(con1, con0) are made inside the
                               # fused local-encoder kernel (else torch.rand / torch._standard_gamma), which then also evaluates
                               # the decoder's hidden2 of the drawn beta


def _sample_local(enc_local, x, mu, K, z_old=None, dec=None):
    """q_probs, z, con1, con0, beta (+ the concentrations at z_old for the backward kernel when z_old is given).
    dec (evaluation, fused encoder, in-kernel draws): the last value returned is the decoder's hidden2 of beta (else None)."""
    if enc_local.can_fuse():
        h2 = None
        if rng_in_kernels and dec is not None and getattr(dec, 'fused', False) and not torch.is_grad_enabled():
            from ..gmm.kernels import philox_seed
            q_probs, z, con1, con0, con1_b, con0_b, beta, h2 = enc_local.sample_fused(
                x, mu, z_old=z_old, seed=philox_seed(mu.device), decoder=dec.recon_mu)
        else:
            q_probs, z, con1, con0, con1_b, con0_b = enc_local.sample_fused(x, mu, z_old=z_old)
            beta = kernels.beta_sample(con1, con0)
        if dec is not None:
            return ((q_probs, z, con1, con0, beta) if z_old is None else (q_probs, z, con1, con0, beta, con1_b, con0_b)) + (h2,)
        return (q_probs, z, con1, con0, beta) if z_old is None else (q_probs, z, con1, con0, beta, con1_b, con0_b)
    if z_old is not None:
        out = _sample_local(enc_local, x, mu, K)
        return out + tuple(enc_local.cons(x, mu, z_old))
    q_probs = enc_local.probs(x, mu, K)
    z = cat(q_probs, validate_args=False).sample()
    con1, con0 = enc_local.cons(x, mu, z)
    beta = Beta(con1, con0, validate_args=False).sample()
    return q_probs, z, con1, con0, beta


def _E_recon(dec, mu, z, h2):
    mu_expand = torch.gather(mu, -2, z.argmax(-1).unsqueeze(-1).expand(-1, -1, -1, mu.shape[-1]))
    return (h2 * dec.radi + mu_expand).mean(0).unsqueeze(0)


def oneshot(enc_rws_mu, enc_rws_local, dec, x, K, trace, result_flags):
    """One-shot predicts mu, z and beta like a normal RWS (objectives.py:80-114)."""
    log_w, mu, z, beta, _, trace = _oneshot(enc_rws_mu, enc_rws_local, dec, x, K, trace, result_flags)
    return log_w, mu, z, beta, trace


def _oneshot(enc_rws_mu, enc_rws_local, dec, x, K, trace, result_flags):
    """oneshot that also returns the decoder output h2 of the new beta."""
    q_mu, q_sigma = enc_rws_mu.params(x, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
    mu = normal_sample(q_mu, q_sigma)
    if enc_rws_local.can_fuse():
        q_probs, z, con1, con0, beta, h2 = _sample_local(enc_rws_local, x, mu, K, dec=dec)
    else:
        (q_probs, z, con1, con0, beta), h2 = _sample_local(enc_rws_local, x, mu, K), None
    if h2 is None:
        h2 = dec.hidden2(beta)
    w = kernels.logw_oneshot(x, q_mu, q_sigma, mu, q_probs, z, beta, con1, con0, h2, dec)
    log_q, ll = w.log_q_f, w.ll_f
    if result_flags['loss_required']:
        if needs_grad(enc_rws_mu, enc_rws_local):            # objectives.py:88-91 with a graph
            q_mu_t, q_sigma_t = enc_rws_mu.params_torch(x, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
            c1_t, c0_t = enc_rws_local.cons(x, mu, z)
            log_q = (Normal(q_mu_t, q_sigma_t, validate_args=False).log_prob(mu).sum(-1).sum(-1)
                     + cat(probs=enc_rws_local.probs(x, mu, K), validate_args=False).log_prob(z).sum(-1)
                     + Beta(c1_t, c0_t, validate_args=False).log_prob(beta).sum(-1).sum(-1))
        if needs_grad(dec):
            ll = dec.log_likelihood_torch(x, mu, z, beta).sum(-1)
    _book(trace, result_flags, w.log_w, log_q, ll, w.density)
    if result_flags['mode_required']:
        trace['E_mu'].append(q_mu.mean(0).unsqueeze(0))
        trace['E_z'].append((q_probs / q_probs.sum(-1, keepdim=True)).mean(0).unsqueeze(0))
        trace['E_recon'].append(_E_recon(dec, mu, z, h2))
    return w.log_w, mu, z, beta, h2, trace


def apg_update_mu(enc_apg_mu, dec, x, z, beta, mu_old, K, trace, result_flags, h2=None, tables=None):
    """Given local variables {z, beta}, update the centres mu (objectives.py:116-151).  h2: the decoder output of
    beta when the caller already holds it; tables: Enc_apg_mu.label_tables(x) when the caller already holds them."""
    if tables is not None:
        q_mu, q_sigma = enc_apg_mu.params(x, z, beta, K, (dec.prior_mu_mu, dec.prior_mu_sigma), tables=tables)
    else:
        q_mu, q_sigma = enc_apg_mu.params(x, z, beta, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
    mu = normal_sample(q_mu, q_sigma)
    if h2 is None:
        h2 = dec.hidden2(beta)             # the decoder sees beta only: one evaluation serves f and b
    w = kernels.logw_mu(x, z, h2, q_mu, q_sigma, mu, mu_old, dec)
    log_q, ll = w.log_q_f, w.ll_f
    if result_flags['loss_required']:
        if needs_grad(enc_apg_mu):                           # objectives.py:123 with a graph
            q_mu_t, q_sigma_t = enc_apg_mu.params_torch(x, z, beta, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
            log_q = Normal(q_mu_t, q_sigma_t, validate_args=False).log_prob(mu).sum(-1).sum(-1)
        if needs_grad(dec):
            ll = dec.log_likelihood_torch(x, mu, z, beta).sum(-1)
    _book(trace, result_flags, w.log_w, log_q, ll, w.density)
    if result_flags['mode_required']:
        trace['E_mu'].append(q_mu.mean(0).unsqueeze(0))
    return w.log_w, mu, trace


def apg_update_local(enc_apg_local, dec, x, mu, z_old, beta_old, K, trace, result_flags):
    """Given mu, update the local variables {z, beta} (objectives.py:153-188)."""
    log_w, z, beta, _, trace = _update_local(enc_apg_local, dec, x, mu, z_old, beta_old, K, trace, result_flags)
    return log_w, z, beta, trace


def _update_local(enc_apg_local, dec, x, mu, z_old, beta_old, K, trace, result_flags, h2_old=None):
    """apg_update_local that takes the decoder output of beta_old when the caller holds it and returns the new one."""
    if enc_apg_local.can_fuse():
        q_probs, z, con1, con0, beta, con1_b, con0_b, h2 = _sample_local(enc_apg_local, x, mu, K, z_old=z_old, dec=dec)
    else:
        (q_probs, z, con1, con0, beta, con1_b, con0_b), h2 = _sample_local(enc_apg_local, x, mu, K, z_old=z_old), None
    if h2 is None:
        h2 = dec.hidden2(beta)
    if h2_old is None:
        h2_old = dec.hidden2(beta_old)
    need_local = bool(result_flags['loss_required'])
    w = kernels.logw_local(x, mu, q_probs, z, beta, con1, con0, h2, z_old, beta_old, con1_b, con0_b, h2_old, dec,
                           local=need_local)
    if result_flags['loss_required']:
        w_local, _ = ops.local_weight_stats(w.log_w_local, want_w=True, want_sum=False)
        log_q_local, ll_local = w.log_q_local, w.ll_local
        if needs_grad(enc_apg_local):                        # objectives.py:162 with a graph
            c1_t, c0_t = enc_apg_local.cons(x, mu, z)
            log_q_local = (cat(probs=enc_apg_local.probs(x, mu, K), validate_args=False).log_prob(z)
                           + Beta(c1_t, c0_t, validate_args=False).log_prob(beta).sum(-1))
        if needs_grad(dec):
            ll_local = dec.log_likelihood_torch(x, mu, z, beta)
        trace['loss_phi'][-1] = trace['loss_phi'][-1] + (w_local * (-log_q_local)).sum(0).sum(-1).mean().unsqueeze(0)
        trace['loss_theta'][-1] = trace['loss_theta'][-1] + (w_local * (-ll_local)).sum(0).sum(-1).mean().unsqueeze(0)
    if result_flags['mode_required']:
        trace['E_z'].append((q_probs / q_probs.sum(-1, keepdim=True)).mean(0).unsqueeze(0))
        trace['E_recon'].append(_E_recon(dec, mu, z, h2))
    if result_flags['density_required']:
        trace['density'][-1] = trace['density'][-1] + w.density.unsqueeze(0)
    return w.log_w, z, beta, h2, trace


def resample_variables(resampler, mu, z, beta, log_weights):
    mu, z, beta = resampler.resample([mu, z, beta], log_weights)
    return mu, z, beta


# ------------------------------------------------------------------ baseline around the path
def bpg_objective(models, x, K, result_flags, num_sweeps, resampler):
    """APG whose mu block proposes from the prior (apgs/dmm/objectives.py:209-224); the reference runs it with
    S * bpg_factor particles, which is what the split-S ancestor pipeline is for."""
    trace = {'density': []}
    (enc_rws_mu, enc_apg_local, enc_apg_mu, dec) = models
    if x.dim() == 3:
        x = x.unsqueeze(0).expand(resampler.S, -1, -1, -1)
    log_w, mu, z, beta, h2, trace = _oneshot(enc_rws_mu, enc_apg_local, dec, x, K, trace, result_flags)
    mu, z, beta, h2 = resampler.resample([mu, z, beta, h2], log_w)
    for m in range(num_sweeps - 1):
        log_w_mu, mu, trace = bpg_update_mu(enc_apg_mu, dec, x, z, beta, mu, K, trace, h2=h2)
        mu, z, beta, h2 = resampler.resample([mu, z, beta, h2], log_w_mu)
        log_w_z, z, beta, h2, trace = _update_local(enc_apg_local, dec, x, mu, z, beta, K, trace, result_flags, h2_old=h2)
        mu, z, beta, h2 = resampler.resample([mu, z, beta, h2], log_w_z)
    trace['density'] = torch.cat(trace['density'], 0)
    return trace


def bpg_update_mu(enc_apg_mu, dec, x, z, beta, mu_old, K, trace, h2=None):
    """mu proposed from its prior: weight = likelihood ratio (apgs/dmm/objectives.py:226-236).  One D2 launch with the
    prior as the proposal -- log p(mu) - log q(mu) cancels inside the kernel -- leaving ll_f - ll_b; density = log p(mu_new)."""
    S, B, K_, D = mu_old.shape
    q_mu = dec.prior_mu_mu.expand(S, B, K_, D).contiguous()
    q_sigma = dec.prior_mu_sigma.expand(S, B, K_, D).contiguous()
    mu = normal_sample(q_mu, q_sigma)
    if h2 is None:
        h2 = dec.hidden2(beta)
    w = kernels.logw_mu(x, z, h2, q_mu, q_sigma, mu, mu_old, dec)
    trace['density'].append(w.density.unsqueeze(0))
    return w.log_w, mu, trace
