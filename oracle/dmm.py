"""Oracle: DMM (ring mixture) reweighting arithmetic + the full APG sweep (CPU restatement).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows, in the reference's operation order (torch-CPU fp32):
  apgs/dmm/models.py:10-210        Enc_rws_mu / Enc_apg_local / Enc_apg_mu / Decoder
  apgs/dmm/objectives.py:9-195     apg_objective / oneshot / apg_update_mu / apg_update_local / resample_variables
Parameters are dicts keyed like the reference's state_dicts; random draws use the same torch calls in the
same order as the reference, so that one torch.manual_seed reproduces the reference's trace.
"""
import math

import torch
import torch.nn.functional as F
from torch.distributions.beta import Beta
from torch.distributions.normal import Normal
from torch.distributions.one_hot_categorical import OneHotCategorical

from . import dists
from .gmm import _Resampler, _resample


# --------------------------------------------------------------------------- parameters
def make_params(K=4, D=2, H_mu=32, NSS=8, H_local=32, H_dec=32):
    """default-initialised weights in the construction order of apgs/dmm/apg_training.py:73-76."""
    def lin(i, o):
        m = torch.nn.Linear(i, o)
        return m.weight.detach().clone(), m.bias.detach().clone()

    def mlp(names_dims):
        d = {}
        for name, i, o in names_dims:
            d[name + '.weight'], d[name + '.bias'] = lin(i, o)
        return d
    P = {}
    P['enc_rws_mu'] = mlp([('nss1.0', D, H_mu), ('nss1.2', H_mu, NSS), ('nss2.0', D, H_mu), ('nss2.2', H_mu, K),
                           ('mean_mu.0', NSS + 2 * D, H_mu), ('mean_mu.2', H_mu, D),
                           ('mean_log_sigma.0', NSS + 2 * D, H_mu), ('mean_log_sigma.2', H_mu, D)])
    P['enc_apg_local'] = mlp([('pi_log_prob.0', D, H_local), ('pi_log_prob.2', H_local, 1),
                              ('angle_log_con1.0', D, H_local), ('angle_log_con1.2', H_local, 1),
                              ('angle_log_con0.0', D, H_local), ('angle_log_con0.2', H_local, 1)])
    P['enc_apg_mu'] = mlp([('nss1.0', D + K, H_mu), ('nss1.2', H_mu, NSS), ('nss2.0', D + K, H_mu), ('nss2.2', H_mu, K),
                           ('mean_mu.0', NSS + 2 * D, H_mu), ('mean_mu.2', H_mu, D),
                           ('mean_log_sigma.0', NSS + 2 * D, H_mu), ('mean_log_sigma.2', H_mu, D)])
    P['dec'] = mlp([('recon_mu.0', 1, H_dec), ('recon_mu.2', H_dec, D)])
    P['dec']['radi'] = torch.ones(1)
    return P


def make_prior(K=4, D=2, recon_sigma=0.5):
    """apgs/dmm/models.py:163-172."""
    return {'mu_mu': torch.zeros(D), 'mu_sigma': torch.ones(D) * 10.0, 'pi': torch.ones(K) * (1. / K),
            'con1': torch.ones(1), 'con0': torch.ones(1), 'recon_sigma': torch.ones(1) * recon_sigma}


def _mlp2(p, name, x, act):
    h = act(F.linear(x, p[name + '.0.weight'], p[name + '.0.bias']))
    return F.linear(h, p[name + '.2.weight'], p[name + '.2.bias'])


# ------------------------------------------------------------------- D1: NSS pooling
def nss_pool(nss1, nss2, prior, EPS=1e-8):
    """models.py:134-137 (and :38-41): nss (S,B,K,NSS) = sum_n nss1[n,:] nss2[n,k] / (sum_n nss2[n,k] + EPS),
    concatenated with the prior mean / scale -> (S,B,K,NSS+2D)."""
    S, B, N, K = nss2.shape
    a = nss1.unsqueeze(-1).repeat(1, 1, 1, 1, K).transpose(-1, -2)
    b = nss2.unsqueeze(-1).repeat(1, 1, 1, 1, a.shape[-1])
    nss = (a * b).sum(2) / (b.sum(2) + EPS)
    return torch.cat((nss, prior['mu_mu'].repeat(S, B, K, 1), prior['mu_sigma'].repeat(S, B, K, 1)), -1)


def enc_mu_params(p, inp, prior):
    """Enc_rws_mu (inp = ob) / Enc_apg_mu (inp = cat(ob, z)) -> q_mu_mu, q_mu_sigma (S,B,K,D)."""
    nss1 = _mlp2(p, 'nss1', inp, torch.tanh)
    nss2 = F.softmax(_mlp2(p, 'nss2', inp, torch.tanh), -1)
    nss_prior = nss_pool(nss1, nss2, prior)
    return _mlp2(p, 'mean_mu', nss_prior, torch.tanh), _mlp2(p, 'mean_log_sigma', nss_prior, torch.tanh).exp()


def enc_local_probs(p, ob, mu, K):
    """models.py:78-79: q_probs (S,B,N,K)."""
    N = ob.shape[-2]
    ob_mu = ob.unsqueeze(2).repeat(1, 1, K, 1, 1) - mu.unsqueeze(-2).repeat(1, 1, 1, N, 1)
    return F.softmax(_mlp2(p, 'pi_log_prob', ob_mu, F.relu).squeeze(-1).transpose(-1, -2), -1)


def enc_local_cons(p, ob, mu, z):
    """models.py:86-88 / :99-101: Beta concentrations (S,B,N,1) given the assignment z."""
    D = ob.shape[-1]
    mu_expand = torch.gather(mu, -2, z.argmax(-1).unsqueeze(-1).repeat(1, 1, 1, D))
    return (_mlp2(p, 'angle_log_con1', ob - mu_expand, torch.tanh).exp(),
            _mlp2(p, 'angle_log_con0', ob - mu_expand, torch.tanh).exp())


def dec_hidden2(p, beta):
    """models.py:202-203: unit-normalised decoder output (S,B,N,D)."""
    hidden = _mlp2(p, 'recon_mu', beta * 2 * math.pi, torch.tanh)
    return hidden / (hidden ** 2).sum(-1).unsqueeze(-1).sqrt()


# ------------------------------------------------------------------ log-density pieces
def likelihood(prior, p_dec, ob, mu, z, hidden2):
    """models.py:204-209: Normal(hidden2 * radi + mu[z], recon_sigma).log_prob(ob) -> (S,B,N,D)."""
    S, B, N, D = ob.shape
    mu_expand = torch.gather(mu, -2, z.argmax(-1).unsqueeze(-1).repeat(1, 1, 1, D))
    recon_mu = hidden2 * p_dec['radi'] + mu_expand
    return dists.normal_log_prob(ob, recon_mu, prior['recon_sigma'].repeat(S, B, N, D))


def mu_log_prior(prior, mu):
    return dists.normal_log_prob(mu, prior['mu_mu'], prior['mu_sigma'])            # S,B,K,D


def mu_step_weights(prior, p_dec, ob, z, hidden2, q_mu, q_sigma, mu, mu_old, parts=False):
    """objectives.py:116-151 given the proposal parameters and the draws -> log_w, log_q_f, log_prior_f, ll_f."""
    log_q_f = dists.normal_log_prob(mu, q_mu, q_sigma).sum(-1).sum(-1)
    ll_f = likelihood(prior, p_dec, ob, mu, z, hidden2).sum(-1).sum(-1)
    log_priors_f = mu_log_prior(prior, mu).sum(-1).sum(-1)
    log_w_f = (log_priors_f + ll_f) - log_q_f
    log_q_b = dists.normal_log_prob(mu_old, q_mu, q_sigma).sum(-1).sum(-1)
    ll_b = likelihood(prior, p_dec, ob, mu_old, z, hidden2).sum(-1).sum(-1)
    log_prior_b = mu_log_prior(prior, mu_old).sum(-1).sum(-1)
    log_w_b = (log_prior_b + ll_b) - log_q_b
    if parts:
        return log_w_f - log_w_b, log_q_f, log_priors_f, ll_f, log_w_f, log_w_b
    return log_w_f - log_w_b, log_q_f, log_priors_f, ll_f


def local_step_weights(prior, p_dec, ob, mu, q_probs, z, beta, con1, con0, hidden2,
                       z_old, beta_old, con1_old, con0_old, hidden2_old, parts=False):
    """objectives.py:153-188 given every proposal parameter and draw.
    -> log_w_local (S,B,N), log_w (S,B), log_q_f (S,B,N), ll_f (S,B,N), log_p_f (S,B,N)."""
    log_q_f = dists.onehot_categorical_log_prob(z, q_probs) + dists.beta_log_prob(beta, con1, con0).sum(-1)
    ll_f = likelihood(prior, p_dec, ob, mu, z, hidden2).sum(-1)
    log_p_f = ll_f + dists.onehot_categorical_log_prob(z, prior['pi']) + dists.beta_log_prob(beta, prior['con1'], prior['con0']).sum(-1)
    log_w_f = log_p_f - log_q_f
    log_q_b = dists.onehot_categorical_log_prob(z_old, q_probs) + dists.beta_log_prob(beta_old, con1_old, con0_old).sum(-1)
    ll_b = likelihood(prior, p_dec, ob, mu, z_old, hidden2_old).sum(-1)
    log_p_b = ll_b + dists.onehot_categorical_log_prob(z_old, prior['pi']) + dists.beta_log_prob(beta_old, prior['con1'], prior['con0']).sum(-1)
    log_w_b = log_p_b - log_q_b
    log_w_local = log_w_f - log_w_b
    if parts:
        return log_w_local, log_w_local.sum(-1), log_q_f, ll_f, log_p_f, log_w_f.sum(-1), log_w_b.sum(-1)
    return log_w_local, log_w_local.sum(-1), log_q_f, ll_f, log_p_f


def oneshot_weights(prior, p_dec, ob, q_mu, q_sigma, mu, q_probs, z, beta, con1, con0, hidden2):
    """objectives.py:80-114 -> log_w, log_q, log_p, ll (S,B)."""
    log_q = (dists.normal_log_prob(mu, q_mu, q_sigma).sum(-1).sum(-1) + dists.onehot_categorical_log_prob(z, q_probs).sum(-1)
             + dists.beta_log_prob(beta, con1, con0).sum(-1).sum(-1))
    ll = likelihood(prior, p_dec, ob, mu, z, hidden2).sum(-1).sum(-1)
    log_p = (ll + mu_log_prior(prior, mu).sum(-1).sum(-1) + dists.onehot_categorical_log_prob(z, prior['pi']).sum(-1)
             + dists.beta_log_prob(beta, prior['con1'], prior['con0']).sum(-1).sum(-1))
    return log_p - log_q, log_q, log_p, ll


# ------------------------------------------------------------------------------ the sweep
def _sample_local(P, ob, mu, K):
    q_probs = enc_local_probs(P['enc_apg_local'], ob, mu, K)
    z = OneHotCategorical(q_probs).sample()
    con1, con0 = enc_local_cons(P['enc_apg_local'], ob, mu, z)
    beta = Beta(con1, con0).sample()
    return q_probs, z, con1, con0, beta


def apg_objective(P, prior, x, K, result_flags, num_sweeps, strategy='systematic', record=None):
    """objectives.py:9-56.  x (S,B,N,D) S-replicated.  -> {'loss_phi','loss_theta','ess','density'}."""
    S = x.shape[0]
    rs = _Resampler(strategy, S)
    trace = {'loss_phi': [], 'loss_theta': [], 'ess': [], 'density': []}
    pd = P['dec']

    def book(log_w, log_q, ll, dens):
        w = F.softmax(log_w, 0)
        if result_flags['loss_required']:
            trace['loss_phi'].append((w * (-log_q)).sum(0).mean().unsqueeze(0))
            trace['loss_theta'].append((w * (-ll)).sum(0).mean().unsqueeze(0))
        if result_flags['ess_required']:
            trace['ess'].append((1. / (w ** 2).sum(0)).unsqueeze(0))
        if result_flags['density_required']:
            trace['density'].append(dens.unsqueeze(0))

    # ---- one-shot (objectives.py:80-114)
    q_mu, q_sigma = enc_mu_params(P['enc_rws_mu'], x, prior)
    mu = Normal(q_mu, q_sigma).sample()
    q_probs, z, con1, con0, beta = _sample_local(P, x, mu, K)
    h2 = dec_hidden2(pd, beta)
    log_w, log_q, log_p, ll = oneshot_weights(prior, pd, x, q_mu, q_sigma, mu, q_probs, z, beta, con1, con0, h2)
    book(log_w, log_q, ll, log_p)
    if record is not None:
        record.append(dict(step='oneshot', q_mu=q_mu, q_sigma=q_sigma, mu=mu, q_probs=q_probs, z=z, beta=beta, con1=con1,
                           con0=con0, hidden2=h2, log_w=log_w, log_q=log_q, log_p=log_p, ll=ll))
    mu, z, beta = _resample(rs, log_w, mu, z, beta)
    for _ in range(num_sweeps - 1):
        # ---- mu update (objectives.py:116-151)
        q_mu, q_sigma = enc_mu_params(P['enc_apg_mu'], torch.cat((x, z), -1), prior)
        mu_new = Normal(q_mu, q_sigma).sample()
        h2 = dec_hidden2(pd, beta)
        log_w, log_q_f, log_priors_f, ll_f = mu_step_weights(prior, pd, x, z, h2, q_mu, q_sigma, mu_new, mu)
        book(log_w, log_q_f, ll_f, log_priors_f)
        if record is not None:
            record.append(dict(step='mu', z=z, beta=beta, hidden2=h2, q_mu=q_mu, q_sigma=q_sigma, mu=mu_new, mu_old=mu,
                               log_w=log_w, log_q_f=log_q_f, log_priors_f=log_priors_f, ll_f=ll_f))
        mu, z, beta = _resample(rs, log_w, mu_new, z, beta)
        # ---- local update (objectives.py:153-188)
        q_probs, z_new, con1, con0, beta_new = _sample_local(P, x, mu, K)
        h2_new = dec_hidden2(pd, beta_new)
        con1_b, con0_b = enc_local_cons(P['enc_apg_local'], x, mu, z)
        h2_old = dec_hidden2(pd, beta)
        lwl, log_w, log_q_f, ll_f, log_p_f = local_step_weights(prior, pd, x, mu, q_probs, z_new, beta_new, con1, con0, h2_new,
                                                                z, beta, con1_b, con0_b, h2_old)
        if result_flags['loss_required']:
            w_local = F.softmax(lwl, 0)
            trace['loss_phi'][-1] = trace['loss_phi'][-1] + (w_local * (-log_q_f)).sum(0).sum(-1).mean().unsqueeze(0)
            trace['loss_theta'][-1] = trace['loss_theta'][-1] + (w_local * (-ll_f)).sum(0).sum(-1).mean().unsqueeze(0)
        if result_flags['density_required']:
            trace['density'][-1] = trace['density'][-1] + log_p_f.sum(-1).unsqueeze(0)
        if record is not None:
            record.append(dict(step='local', mu=mu, q_probs=q_probs, z=z_new, beta=beta_new, con1=con1, con0=con0,
                               hidden2=h2_new, z_old=z, beta_old=beta, con1_old=con1_b, con0_old=con0_b, hidden2_old=h2_old,
                               log_w_local=lwl, log_w=log_w, log_q_f=log_q_f, ll_f=ll_f, log_p_f=log_p_f))
        mu, z, beta = _resample(rs, log_w, mu, z_new, beta_new)
    out = {}
    for k in ('loss_phi', 'loss_theta', 'ess', 'density'):
        flag = 'loss_required' if k.startswith('loss') else k + '_required'
        if result_flags[flag]:
            out[k] = torch.cat(trace[k], 0)
    return out


# ----------------------------------------------------------------------- synthetic data
def simulate(B, N=200, K=4, D=2, radius=2.0, noise=0.1, center_std=3.0):
    """Ring mixtures as apgs/dmm/sim_dmm.py:27-62 (args :93-99): K rings of radius 2 centred at N(0, 3^2), angles
    uniform, isotropic noise 0.1 -> x (B,N,D)."""
    centers = torch.randn(B, K, D) * center_std
    labels = torch.randint(0, K, (B, N))
    ang = torch.rand(B, N) * 2 * math.pi
    ring = torch.stack((torch.cos(ang), torch.sin(ang)), -1) * radius
    c = torch.gather(centers, 1, labels.unsqueeze(-1).expand(B, N, D))
    return c + ring + torch.randn(B, N, D) * noise
