"""Oracle: GMM reweighting arithmetic + the full APG sweep (CPU restatement).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows, in the reference's operation order (torch-CPU fp32):
  apgs/gmm/kls_gmm.py:92-119     data_to_stats / posterior_eta   (quirks Q1, Q2)
  apgs/gmm/models.py:10-115      Enc_rws_eta / Enc_apg_eta / Enc_apg_z
  apgs/gmm/models.py:139-185     Generative.eta_prior / z_prior / log_prob (quirk Q3)
  apgs/gmm/objectives.py:9-234   apg_objective / oneshot / apg_update_{eta,z,joint} /
                                 resample_variables
  apgs/gmm/objectives.py:236-330 gibbs_objective / gibbs_sweep / bpg_objective / bpg_update_eta (baselines, N4)
  apgs/gmm/kls_gmm.py:121-133    posterior_z

Parameters are plain dicts keyed like the reference's nn.Module state_dicts
('gamma.0.weight', 'ob.0.weight', 'pi_log_prob.0.weight', ...).  Random draws use the
same torch.distributions calls in the same order as the reference, so that under one
torch.manual_seed the whole trace is reproduced (tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F
from torch.distributions.gamma import Gamma
from torch.distributions.normal import Normal
from torch.distributions.one_hot_categorical import OneHotCategorical

from . import dists
from . import resampler as _rs


# --------------------------------------------------------------------------- parameters
def make_params(K=3, D=2, H=32):
    """Default-initialised encoder weights, constructed in the order of
    apgs/gmm/apg_training.py:76-79 (Enc_rws_eta, Enc_apg_z, Enc_apg_eta) so that a
    shared torch.manual_seed gives the reference's weights."""
    def lin(i, o):
        m = torch.nn.Linear(i, o)
        return m.weight.detach().clone(), m.bias.detach().clone()
    P = {}
    gw, gb = lin(D, K)
    ow, ob = lin(D, D)
    P['enc_rws_eta'] = {'gamma.0.weight': gw, 'gamma.0.bias': gb, 'ob.0.weight': ow, 'ob.0.bias': ob}
    w0, b0 = lin(3 * D, H)
    w2, b2 = lin(H, 1)
    P['enc_apg_z'] = {'pi_log_prob.0.weight': w0, 'pi_log_prob.0.bias': b0,
                      'pi_log_prob.2.weight': w2, 'pi_log_prob.2.bias': b2}
    gw, gb = lin(K + D, K)
    ow, ob = lin(K + D, D)
    P['enc_apg_eta'] = {'gamma.0.weight': gw, 'gamma.0.bias': gb, 'ob.0.weight': ow, 'ob.0.bias': ob}
    return P


def make_prior(K=3, D=2):
    """apgs/gmm/models.py:124-128."""
    return {'mu': torch.zeros(K, D), 'nu': torch.ones(K, D) * 0.1,
            'alpha': torch.ones(K, D) * 2, 'beta': torch.ones(K, D) * 2,
            'pi': torch.ones(K) * (1. / K)}


# ------------------------------------------------------------------- G1: NSS posterior
def nss_posterior(xp, gamma, prior):
    """kls_gmm.py:92-119.  xp (S,B,N,D), gamma (S,B,N,K) -> alpha,beta,mu,nu (S,B,K,D).
    Q1: zero counts are replaced by 1 in EVERY use of the count (aliasing at :112-113).
    Q2: the last term of beta uses (x_bar - prior_nu)^2 (:118)."""
    D = xp.shape[-1]
    s1 = gamma.sum(2).unsqueeze(-1)                                   # S,B,K,1
    g5 = gamma.unsqueeze(-1)                                          # S,B,N,K,1
    x5 = xp.unsqueeze(-2)                                             # S,B,N,1,D
    s2 = (g5 * x5).sum(2)                                             # S,B,K,D
    s3 = (g5 * (x5 ** 2)).sum(2)
    cnt = s1.repeat(1, 1, 1, D)
    cnt[cnt == 0.0] = 1.0
    x_bar = s2 / cnt
    alpha = prior['alpha'] + cnt / 2
    nu = prior['nu'] + cnt
    mu = (prior['mu'] * prior['nu'] + s2) / (cnt + prior['nu'])
    beta = (prior['beta'] + (s3 - (s2 ** 2) / cnt) / 2.
            + (cnt * prior['nu'] / (cnt + prior['nu'])) * ((x_bar - prior['nu']) ** 2) / 2.)
    return alpha, beta, mu, nu


# ------------------------------------------------------------------------ encoder heads
def enc_eta_heads(p, inp):
    """gamma = softmax(Linear(inp)), x' = Linear(inp)   (models.py:17-22, 56-65)."""
    gamma = F.softmax(F.linear(inp, p['gamma.0.weight'], p['gamma.0.bias']), -1)
    xp = F.linear(inp, p['ob.0.weight'], p['ob.0.bias'])
    return xp, gamma


def enc_rws_eta_posterior(p, x, prior):
    xp, gamma = enc_eta_heads(p, x)
    return nss_posterior(xp, gamma, prior)


def enc_apg_eta_posterior(p, x, z, prior):
    xp, gamma = enc_eta_heads(p, torch.cat((x, z), -1))
    return nss_posterior(xp, gamma, prior)


def enc_apg_z_probs(p, x, tau, mu):
    """models.py:102-109: per-cluster MLP on cat(x, mu_k, tau_k), softmax over k."""
    N = x.shape[-2]
    logits = []
    for k in range(mu.shape[-2]):
        feat = torch.cat((x, mu[:, :, k, :].unsqueeze(-2).repeat(1, 1, N, 1),
                          tau[:, :, k, :].unsqueeze(-2).repeat(1, 1, N, 1)), -1)
        h = torch.tanh(F.linear(feat, p['pi_log_prob.0.weight'], p['pi_log_prob.0.bias']))
        logits.append(F.linear(h, p['pi_log_prob.2.weight'], p['pi_log_prob.2.bias']))
    return F.softmax(torch.cat(logits, -1), -1)


# ------------------------------------------------------------------ log-density pieces
def eta_log_q(tau, mu, post):
    """log q(tau, mu) under Normal-Gamma(alpha,beta,mu_q,nu_q), summed over (K,D)
    (models.py:69-78 + objectives.py:167).  Q3: scale = 1/sqrt(nu*tau)."""
    alpha, beta, mu_q, nu_q = post
    lp_tau = dists.gamma_log_prob(tau, alpha, beta)
    lp_mu = dists.normal_log_prob(mu, mu_q, 1. / (nu_q * tau).sqrt())
    return lp_mu.sum(-1).sum(-1) + lp_tau.sum(-1).sum(-1)


def eta_log_prior(tau, mu, prior):
    """models.py:139-150 + objectives.py:168."""
    lp_tau = dists.gamma_log_prob(tau, prior['alpha'], prior['beta'])
    lp_mu = dists.normal_log_prob(mu, prior['mu'], 1. / (prior['nu'] * tau).sqrt())
    return lp_mu.sum(-1).sum(-1) + lp_tau.sum(-1).sum(-1)


def log_likelihood(x, z, tau, mu, aggregate):
    """models.py:172-185: Normal(mu[z_n], 1/sqrt(tau[z_n])).log_prob(x_n) summed over D
    (and over N when aggregate)."""
    sigma = 1. / tau.sqrt()
    labels = z.argmax(-1).unsqueeze(-1).repeat(1, 1, 1, x.shape[-1])
    ll = dists.normal_log_prob(x, torch.gather(mu, 2, labels), torch.gather(sigma, 2, labels)).sum(-1)
    return ll.sum(-1) if aggregate else ll


def z_log_q(z, q_probs):
    return dists.onehot_categorical_log_prob(z, q_probs)                 # S,B,N


def z_log_prior(z, prior):
    return dists.onehot_categorical_log_prob(z, prior['pi'])             # S,B,N


# --------------------------------------------------- G2 / G3 / G4: importance weights
def eta_step_weights(x, z, post, prior, tau, mu, tau_old, mu_old, parts=False):
    """objectives.py:161-195 given the posterior parameters and the draws.
    Returns log_w, log_q_f, log_p_f (the 'density' increment), ll_f   -- all (S,B)."""
    log_q_f = eta_log_q(tau, mu, post)
    log_p_f = eta_log_prior(tau, mu, prior)
    ll_f = log_likelihood(x, z, tau, mu, True)
    log_w_f = ll_f + log_p_f - log_q_f
    log_q_b = eta_log_q(tau_old, mu_old, post)
    log_p_b = eta_log_prior(tau_old, mu_old, prior)
    ll_b = log_likelihood(x, z, tau_old, mu_old, True)
    log_w_b = ll_b + log_p_b - log_q_b
    if parts:
        return log_w_f - log_w_b, log_q_f, log_p_f, ll_f, log_w_f, log_w_b
    return log_w_f - log_w_b, log_q_f, log_p_f, ll_f


def z_step_weights(x, tau, mu, q_probs, prior, z, z_old, parts=False):
    """objectives.py:197-227 given q_probs and the draws.
    Returns log_w_local (S,B,N), log_w (S,B), log_q_f (S,B,N), ll_f + log_p_f (S,B,N)."""
    log_q_f = z_log_q(z, q_probs)
    log_p_f = z_log_prior(z, prior)
    ll_f = log_likelihood(x, z, tau, mu, False)
    log_w_f = ll_f + log_p_f - log_q_f
    log_q_b = z_log_q(z_old, q_probs)
    log_p_b = z_log_prior(z_old, prior)
    ll_b = log_likelihood(x, z_old, tau, mu, False)
    log_w_b = ll_b + log_p_b - log_q_b
    log_w_local = log_w_f - log_w_b
    if parts:
        return log_w_local, log_w_local.sum(-1), log_q_f, ll_f + log_p_f, log_w_f.sum(-1), log_w_b.sum(-1)
    return log_w_local, log_w_local.sum(-1), log_q_f, ll_f + log_p_f


def oneshot_weights(x, post, prior, tau, mu, q_probs, z):
    """objectives.py:74-110 given parameters and draws -> log_w, log_q, log_p (S,B)."""
    log_q_eta = eta_log_q(tau, mu, post)
    log_p_eta = eta_log_prior(tau, mu, prior)
    log_q_z = z_log_q(z, q_probs).sum(-1)
    log_p_z = z_log_prior(z, prior).sum(-1)
    ll = log_likelihood(x, z, tau, mu, True)
    log_p = ll + log_p_eta + log_p_z
    log_q = log_q_eta + log_q_z
    return log_p - log_q, log_q, log_p


def joint_step_weights(x, post, prior, tau, mu, tau_old, mu_old, q_probs, z, z_old, parts=False):
    """objectives.py:112-158 given parameters and draws.  q_probs is evaluated at the
    NEW (tau, mu) for both the forward and the backward z-term (:123,:131)."""
    log_q_f_eta = eta_log_q(tau, mu, post)
    log_p_f_eta = eta_log_prior(tau, mu, prior)
    log_q_f_z = z_log_q(z, q_probs).sum(-1)
    log_p_f_z = z_log_prior(z, prior).sum(-1)
    ll_f = log_likelihood(x, z, tau, mu, True)
    log_w_f = ll_f + log_p_f_eta - log_q_f_eta - log_q_f_z + log_p_f_z
    log_q_b_z = z_log_q(z_old, q_probs).sum(-1)
    log_p_b_z = z_log_prior(z_old, prior).sum(-1)
    log_q_b_eta = eta_log_q(tau_old, mu_old, post)
    log_p_b_eta = eta_log_prior(tau_old, mu_old, prior)
    ll_b = log_likelihood(x, z_old, tau_old, mu_old, True)
    log_w_b = ll_b + log_p_b_eta - log_q_b_eta + log_p_b_z - log_q_b_z
    if parts:
        return (log_w_f - log_w_b, log_q_f_eta + log_q_f_z, ll_f + log_p_f_eta + log_p_f_z, log_w_f, log_w_b)
    return (log_w_f - log_w_b, log_q_f_eta + log_q_f_z, ll_f + log_p_f_eta + log_p_f_z)


# ------------------------------------------------------------------------- the sweep
def _sample_eta(post):
    alpha, beta, mu_q, nu_q = post
    tau = Gamma(alpha, beta).sample()
    mu = Normal(mu_q, 1. / (nu_q * tau).sqrt()).sample()
    return tau, mu


class _Resampler:
    """apgs/resampler.py:7-54 restated on top of oracle.resampler (torch-CPU RNG)."""

    def __init__(self, strategy, S):
        self.strategy, self.S = strategy, S

    def ancestors(self, log_w):
        S, B = log_w.shape
        w = F.softmax(log_w, 0)
        if self.strategy == 'systematic':
            u = torch.rand(B, 1)                                          # Q5
            a = _rs.systematic_ancestors(w.numpy(), u.numpy().reshape(-1))
        else:
            u = torch.rand(B, S, dtype=torch.float64)
            a = _rs.multinomial_ancestors(_rs.categorical_probs(w).numpy(), u.numpy())
        return torch.from_numpy(a)

    def gather(self, var, a):
        idx = a.reshape(a.shape + (1,) * (var.dim() - 2)).expand_as(var)
        return torch.gather(var, 0, idx)


def _resample(rs, log_w, *latents):
    a = rs.ancestors(log_w)
    return tuple(rs.gather(v, a) for v in latents)


def apg_objective(P, prior, x, result_flags, num_sweeps, block='decomposed',
                  strategy='systematic', record=None, mirror_cost=False):
    """objectives.py:9-53.  x is (S,B,N,D) (S-replicated, as the reference passes it).
    Returns the trace dict {'loss','ess','density'} (mode outputs are not restated).
    `record`, when a list, receives a dict of every tensor entering/leaving each
    reweight+resample step (teacher-forcing fixtures for the CUDA parity tests)."""
    S = x.shape[0]
    rs = _Resampler(strategy, S)
    trace = {'loss': [], 'ess': [], 'density': []}

    def book(log_w, log_q_for_loss, density):
        w = F.softmax(log_w, 0)
        if result_flags['loss_required']:
            trace['loss'].append((w * (-log_q_for_loss)).sum(0).mean().unsqueeze(0))
        if result_flags['ess_required']:
            trace['ess'].append((1. / (w ** 2).sum(0)).unsqueeze(0))
        if result_flags['density_required']:
            trace['density'].append(density.unsqueeze(0))

    # ---- one-shot (objectives.py:74-110)
    post = enc_rws_eta_posterior(P['enc_rws_eta'], x, prior)
    tau, mu = _sample_eta(post)
    q_probs = enc_apg_z_probs(P['enc_apg_z'], x, tau, mu)
    z = OneHotCategorical(q_probs).sample()
    log_w, log_q, log_p = oneshot_weights(x, post, prior, tau, mu, q_probs, z)
    book(log_w, log_q, log_p)
    if record is not None:
        record.append(dict(step='oneshot', post=post, tau=tau, mu=mu, q_probs=q_probs, z=z,
                           log_w=log_w, log_q=log_q, log_p=log_p))
    tau, mu, z = _resample(rs, log_w, tau, mu, z)

    for _ in range(num_sweeps - 1):
        if block == 'decomposed':
            # ---- eta update (objectives.py:161-195)
            post = enc_apg_eta_posterior(P['enc_apg_eta'], x, z, prior)
            tau_new, mu_new = _sample_eta(post)
            if mirror_cost:
                post = enc_apg_eta_posterior(P['enc_apg_eta'], x, z, prior)
            log_w, log_q_f, log_p_f, ll_f = eta_step_weights(x, z, post, prior, tau_new, mu_new, tau, mu)
            book(log_w, log_q_f, log_p_f)
            if record is not None:
                record.append(dict(step='eta', z=z, post=post, tau=tau_new, mu=mu_new, tau_old=tau,
                                   mu_old=mu, log_w=log_w, log_q_f=log_q_f, log_p_f=log_p_f, ll_f=ll_f))
            tau, mu, z = _resample(rs, log_w, tau_new, mu_new, z)
            # ---- z update (objectives.py:197-227)
            q_probs = enc_apg_z_probs(P['enc_apg_z'], x, tau, mu)
            z_new = OneHotCategorical(q_probs).sample()
            if mirror_cost:
                q_probs = enc_apg_z_probs(P['enc_apg_z'], x, tau, mu)
            log_w_local, log_w, log_q_f, dens = z_step_weights(x, tau, mu, q_probs, prior, z_new, z)
            if result_flags['loss_required']:
                w_local = F.softmax(log_w_local, 0)
                trace['loss'][-1] = trace['loss'][-1] + (w_local * (-log_q_f)).sum(0).sum(-1).mean().unsqueeze(0)
            if result_flags['density_required']:
                trace['density'][-1] = trace['density'][-1] + dens.sum(-1).unsqueeze(0)
            if record is not None:
                record.append(dict(step='z', tau=tau, mu=mu, q_probs=q_probs, z=z_new, z_old=z,
                                   log_w_local=log_w_local, log_w=log_w, log_q_f=log_q_f, dens=dens))
            tau, mu, z = _resample(rs, log_w, tau, mu, z_new)
        elif block == 'joint':
            post = enc_apg_eta_posterior(P['enc_apg_eta'], x, z, prior)
            tau_new, mu_new = _sample_eta(post)
            q_probs = enc_apg_z_probs(P['enc_apg_z'], x, tau_new, mu_new)
            z_new = OneHotCategorical(q_probs).sample()
            if mirror_cost:
                q_probs = enc_apg_z_probs(P['enc_apg_z'], x, tau_new, mu_new)
                post = enc_apg_eta_posterior(P['enc_apg_eta'], x, z, prior)
            log_w, log_q_f, dens = joint_step_weights(x, post, prior, tau_new, mu_new, tau, mu,
                                                      q_probs, z_new, z)
            book(log_w, log_q_f, dens)
            if record is not None:
                record.append(dict(step='joint', post=post, tau=tau_new, mu=mu_new, tau_old=tau, mu_old=mu,
                                   q_probs=q_probs, z=z_new, z_old=z, log_w=log_w, log_q_f=log_q_f, dens=dens))
            tau, mu, z = _resample(rs, log_w, tau_new, mu_new, z_new)
        else:
            raise ValueError(block)
    out = {}
    for k in ('loss', 'ess', 'density'):
        if result_flags[k + '_required']:
            out[k] = torch.cat(trace[k], 0)
    return out


# ------------------------------------------------- baselines around the path
def posterior_z_logits(x, tau, mu, prior):
    """kls_gmm.py:121-133: log posterior responsibilities (S,B,N,K) given eta."""
    N, K = x.shape[-2], mu.shape[-2]
    sigma = 1. / tau.sqrt()
    mu_e = mu.unsqueeze(-2).repeat(1, 1, 1, N, 1)
    sigma_e = sigma.unsqueeze(-2).repeat(1, 1, 1, N, 1)
    ob_e = x.unsqueeze(2).repeat(1, 1, K, 1, 1)
    log_gammas = Normal(mu_e, sigma_e).log_prob(ob_e).sum(-1).transpose(-1, -2) + prior['pi'].log()
    return F.softmax(log_gammas, dim=-1).log()


def _oneshot(P, prior, x):
    post = enc_rws_eta_posterior(P['enc_rws_eta'], x, prior)
    tau, mu = _sample_eta(post)
    q_probs = enc_apg_z_probs(P['enc_apg_z'], x, tau, mu)
    z = OneHotCategorical(q_probs).sample()
    log_w, log_q, log_p = oneshot_weights(x, post, prior, tau, mu, q_probs, z)
    return log_w, log_p, tau, mu, z


def gibbs_objective(P, prior, x, num_sweeps):
    """objectives.py:236-274: one-shot initialisation, then exact conjugate Gibbs sweeps; trace of the log joint
    (num_sweeps, S, B)."""
    _, log_p, tau, mu, z = _oneshot(P, prior, x)
    dens = [log_p.unsqueeze(0)]
    for _ in range(num_sweeps - 1):
        post = nss_posterior(x, z, prior)                        # posterior_eta(x, z=z, ...): the same formulas (Q1, Q2)
        tau, mu = _sample_eta(post)
        logits = posterior_z_logits(x, tau, mu, prior)
        z = OneHotCategorical(logits=logits).sample()
        lp_tau = dists.gamma_log_prob(tau, prior['alpha'], prior['beta']).sum(-1).sum(-1)
        lp_mu = dists.normal_log_prob(mu, prior['mu'], 1. / (prior['nu'] * tau).sqrt()).sum(-1).sum(-1)
        log_joint = (log_likelihood(x, z, tau, mu, True) + lp_tau + lp_mu                        # the order of :270
                     + dists.onehot_categorical_log_prob(z, prior['pi']).sum(-1))
        dens.append(log_joint.unsqueeze(0))
    return {'density': torch.cat(dens, 0)}


def bpg_objective(P, prior, x, num_sweeps, strategy='systematic'):
    """objectives.py:291-330: APG whose eta block proposes from the PRIOR (weight = likelihood ratio); density trace
    (num_sweeps, S, B): one-shot log joint, then the prior log-density of each proposed eta plus the z-step density."""
    S, B = x.shape[0], x.shape[1]
    rs = _Resampler(strategy, S)
    log_w, log_p, tau, mu, z = _oneshot(P, prior, x)
    dens = [log_p.unsqueeze(0)]
    tau, mu, z = _resample(rs, log_w, tau, mu, z)
    K, D = prior['alpha'].shape
    for _ in range(num_sweeps - 1):
        rep = lambda t: t.unsqueeze(0).unsqueeze(0).repeat(S, B, 1, 1)
        tau_new = Gamma(rep(prior['alpha']), rep(prior['beta'])).sample()
        mu_new = Normal(rep(prior['mu']), 1. / (prior['nu'] * tau_new).sqrt()).sample()
        log_p_f = eta_log_prior(tau_new, mu_new, prior)
        log_w = log_likelihood(x, z, tau_new, mu_new, True) - log_likelihood(x, z, tau, mu, True)
        dens.append(log_p_f.unsqueeze(0))
        tau, mu, z = _resample(rs, log_w, tau_new, mu_new, z)
        q_probs = enc_apg_z_probs(P['enc_apg_z'], x, tau, mu)
        z_new = OneHotCategorical(q_probs).sample()
        log_w_local, log_w, log_q_f, d = z_step_weights(x, tau, mu, q_probs, prior, z_new, z)
        dens[-1] = dens[-1] + d.sum(-1).unsqueeze(0)
        tau, mu, z = _resample(rs, log_w, tau, mu, z_new)
    return {'density': torch.cat(dens, 0)}


# ---------------------------------------------------------------------- synthetic data
def simulate(B, N=60, K=3, D=2, alpha=2.0, beta=2.0, mu0=0.0, nu0=0.1):
    """Instances drawn as apgs/gmm/sim_gmm.py:30-38 (priors :78-81): returns x (B,N,D)."""
    tau = Gamma(torch.ones(B, K, D) * alpha, torch.ones(B, K, D) * beta).sample()
    mean = Normal(torch.ones(B, K, D) * mu0, 1. / (tau * nu0).sqrt()).sample()
    labels = torch.randint(0, K, (B, N))
    idx = labels.unsqueeze(-1).expand(B, N, D)
    return Normal(torch.gather(mean, 1, idx), torch.gather(1. / tau.sqrt(), 1, idx)).sample()
