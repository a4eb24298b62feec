"""Oracle: bouncing-MNIST reweighting arithmetic + the full APG sweep (CPU restatement).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows, in the reference's operation order (torch-CPU fp32):
  apgs/bmnist/affine_transformer.py:41-65   crop / paste (closed form of oracle/stn.py, vectorised in torch)
  apgs/bmnist/models.py:7-162               Enc_coor / Enc_digit / Dec_coor / Dec_digit / MBern_log_prob
  apgs/bmnist/objectives.py:5-257           resample_variables / apg_objective / propose_one_movement /
                                            oneshot / apg_where / apg_what
Parameters are dicts keyed like the reference's state_dicts; random draws use the same torch calls in the
same order as the reference, so that one torch.manual_seed reproduces the reference's trace.
"""
import math

import torch
import torch.nn.functional as F
from torch.distributions.normal import Normal

from . import dists
from . import resampler as _rs
from .gmm import _Resampler, _resample


# --------------------------------------------------------------------------- parameters
def make_params(FP=96, DP=28, H_digit=400, H_coor=400, ZW=2, ZD=10):
    """default-initialised weights in the construction order of apgs/bmnist/apg_training.py:62-66."""
    def lin(i, o):
        m = torch.nn.Linear(i, o)
        return m.weight.detach().clone(), m.bias.detach().clone()

    def pack(prefix_layers):
        d = {}
        for name, (w, b) in prefix_layers:
            d[name + '.weight'], d[name + '.bias'] = w, b
        return d
    npx = (FP - DP + 1) ** 2
    P = {}
    P['enc_coor'] = pack([('enc_hidden.0', lin(npx, H_coor)),
                          ('where_mean.0', lin(H_coor, H_coor // 2)), ('where_mean.2', lin(H_coor // 2, ZW)),
                          ('where_log_std.0', lin(H_coor, H_coor // 2)), ('where_log_std.2', lin(H_coor // 2, ZW))])
    P['enc_digit'] = pack([('enc_hidden.0', lin(DP * DP, H_digit)), ('enc_hidden.2', lin(H_digit, H_digit // 2)),
                           ('q_mean.0', lin(H_digit // 2, ZD)), ('q_log_std.0', lin(H_digit // 2, ZD))])
    P['dec_digit'] = pack([('digit_mean.0', lin(ZD, H_digit // 2)), ('digit_mean.2', lin(H_digit // 2, H_digit)),
                           ('digit_mean.4', lin(H_digit, DP * DP))])
    return P


def make_prior(ZW=2):
    """apgs/bmnist/models.py:88-90."""
    return {'mu0': torch.zeros(ZW), 'Sigma0': torch.ones(ZW) * 1.0, 'Sigmat': torch.ones(ZW) * 0.2}


# --------------------------------------------------------------------------- networks
def enc_coor_params(p, conved):
    """models.py:27-31 -> q_mean, q_std."""
    hidden = F.relu(F.linear(conved, p['enc_hidden.0.weight'], p['enc_hidden.0.bias']))
    m = F.relu(F.linear(hidden, p['where_mean.0.weight'], p['where_mean.0.bias']))
    q_mean = torch.tanh(F.linear(m, p['where_mean.2.weight'], p['where_mean.2.bias']))
    s = F.relu(F.linear(hidden, p['where_log_std.0.weight'], p['where_log_std.0.bias']))
    q_std = F.linear(s, p['where_log_std.2.weight'], p['where_log_std.2.bias']).exp()
    return q_mean, q_std


def enc_digit_params(p, cropped):
    """models.py:56-61: hidden averaged over T (dim 2) -> q_mu, q_std (S,B,K,ZD)."""
    h = F.relu(F.linear(cropped, p['enc_hidden.0.weight'], p['enc_hidden.0.bias']))
    h = F.relu(F.linear(h, p['enc_hidden.2.weight'], p['enc_hidden.2.bias'])).mean(2)
    return F.linear(h, p['q_mean.0.weight'], p['q_mean.0.bias']), F.linear(h, p['q_log_std.0.weight'], p['q_log_std.0.bias']).exp()


def dec_digit_mean(p, z_what):
    """models.py:138-141 -> digit templates (S,B,K,DP,DP)."""
    h = F.relu(F.linear(z_what, p['digit_mean.0.weight'], p['digit_mean.0.bias']))
    h = F.relu(F.linear(h, p['digit_mean.2.weight'], p['digit_mean.2.bias']))
    d = torch.sigmoid(F.linear(h, p['digit_mean.4.weight'], p['digit_mean.4.bias']))
    DP = int(math.sqrt(d.shape[-1]))
    return d.view(d.shape[:-1] + (DP, DP))


# --------------------------------------------------------------------------- transformer
def _axis_index(n_out, scale, t, n_in):
    lin = torch.linspace(-1, 1, n_out)
    v = lin * torch.tensor(scale, dtype=torch.float32) + t.unsqueeze(-1)
    u = (v + 1) * ((n_in - 1) / 2)
    idx = torch.round(u)
    ok = (idx >= 0) & (idx <= n_in - 1)
    return idx.clamp(0, n_in - 1).long(), ok


def _sample_separable(img, row, rok, col, cok):
    """img (..., H, W); row/col (..., n_out) -> (..., n_out, n_out) nearest samples, zero where invalid."""
    n_out = row.shape[-1]
    W = img.shape[-1]
    rows = torch.gather(img, -2, row.unsqueeze(-1).expand(row.shape + (W,)))
    out = torch.gather(rows, -1, col.unsqueeze(-2).expand(col.shape[:-1] + (n_out, n_out)))
    return out * (rok.unsqueeze(-1) & cok.unsqueeze(-2)).to(out.dtype)


def digit_to_frame(digit, z_where, FP=96, DP=28):
    """affine_transformer.py:41-52: (S,B,K,DP,DP) x (S,B,T,K,2) -> (S,B,T,K,FP,FP)."""
    T = z_where.shape[2]
    scale = float(torch.tensor(FP / DP, dtype=torch.float32))
    tr = (FP - DP) / DP
    col, cok = _axis_index(FP, scale, -1 * (z_where[..., 0] * tr), DP)
    row, rok = _axis_index(FP, scale, z_where[..., 1] * tr, DP)
    img = digit.unsqueeze(2).expand(-1, -1, T, -1, -1, -1)
    return _sample_separable(img, row, rok, col, cok)


def frame_to_digit(frames, z_where, FP=96, DP=28):
    """affine_transformer.py:54-65: (S,B,T,FP,FP) x (S,B,T,K,2) -> (S,B,T,K,DP,DP)."""
    K = z_where.shape[3]
    scale = float(torch.tensor(DP / FP, dtype=torch.float32))
    tr = (FP - DP) / FP
    col, cok = _axis_index(DP, scale, z_where[..., 0] * tr, FP)
    row, rok = _axis_index(DP, scale, -1 * (z_where[..., 1] * tr), FP)
    img = frames.unsqueeze(3).expand(-1, -1, -1, K, -1, -1)
    return _sample_separable(img, row, rok, col, cok)


def recon_ll(digit, z_where, frames, EPS=1e-9):
    """models.py:148-162 -> recon (S,B,T,FP,FP), ll (S,B,T)."""
    FP, DP = frames.shape[-1], digit.shape[-1]
    recon = torch.clamp(digit_to_frame(digit, z_where, FP, DP).sum(-3), min=0.0, max=1.0)
    ll = (torch.log(recon + EPS) * frames + torch.log(1 - recon + EPS) * (1 - frames)).sum(-1).sum(-1)
    return recon, ll


def what_log_prior(z_what):
    """models.py:150: N(0,1) log-density summed over the code dimension -> (S,B,K)."""
    return dists.normal_log_prob(z_what, torch.zeros(z_what.shape[-1]), torch.ones(z_what.shape[-1])).sum(-1)


def coor_log_prior(prior, z_t, z_t_1=None):
    """models.py:97-104 -> (S,B)."""
    if z_t_1 is None:
        return dists.normal_log_prob(z_t, prior['mu0'], prior['Sigma0']).sum(-1)
    return dists.normal_log_prob(z_t, z_t_1, prior['Sigmat']).sum(-1)


# ----------------------------------------------------------------- one frame of z_where
def propose_one_movement(P, prior, frame, template, z_where_t_1, z_where_old_t, z_where_old_t_1, FP=96, DP=28,
                         forced=None):
    """objectives.py:60-107.  Returns dict with log_p_f, log_q_f, (log_p_b, log_q_b), z_where (S,B,K,2),
    q_mean, q_std (S,B,K,2).  `forced` (S,B,K,2) replaces the draws (teacher forcing)."""
    S, B, K = template.shape[:3]
    frame_left = frame
    zs, means, stds = [], [], []
    log_q_f, log_p_f, log_q_b, log_p_b = [], [], [], []
    for k in range(K):
        template_k = template[:, :, k, :, :]
        conved = F.conv2d(frame_left.view(S * B, FP, FP).unsqueeze(0), template_k.reshape(S * B, DP, DP).unsqueeze(1), groups=int(S * B))
        CP = conved.shape[-1]
        conved = F.softmax(conved.squeeze(0).view(S, B, CP, CP).view(S, B, CP * CP), -1)
        q_mean, q_std = enc_coor_params(P['enc_coor'], conved)
        z_k = Normal(q_mean, q_std).sample() if forced is None else forced[:, :, k, :]
        zs.append(z_k.unsqueeze(2)); means.append(q_mean.unsqueeze(2)); stds.append(q_std.unsqueeze(2))
        log_q_f.append(dists.normal_log_prob(z_k, q_mean, q_std).sum(-1).unsqueeze(-1))
        log_p_f.append(coor_log_prior(prior, z_k, None if z_where_t_1 is None else z_where_t_1[:, :, k, :]).unsqueeze(-1))
        recon_k = digit_to_frame(template_k.unsqueeze(2), z_k.unsqueeze(2).unsqueeze(2), FP, DP).squeeze(2).squeeze(2)
        frame_left = frame_left - recon_k
        if z_where_old_t is not None:
            log_q_b.append(dists.normal_log_prob(z_where_old_t[:, :, k, :], q_mean, q_std).sum(-1).unsqueeze(-1))
            log_p_b.append(coor_log_prior(prior, z_where_old_t[:, :, k, :],
                                          None if z_where_old_t_1 is None else z_where_old_t_1[:, :, k, :]).unsqueeze(-1))
    out = {'z_where': torch.cat(zs, 2), 'q_mean': torch.cat(means, 2), 'q_std': torch.cat(stds, 2),
           'log_p_f': torch.cat(log_p_f, -1).sum(-1), 'log_q_f': torch.cat(log_q_f, -1).sum(-1)}
    if z_where_old_t is not None:
        out['log_p_b'] = torch.cat(log_p_b, -1).sum(-1)
        out['log_q_b'] = torch.cat(log_q_b, -1).sum(-1)
    return out


def where_step_weight(pm, ll_f, ll_b):
    """objectives.py:198-206."""
    log_w_f = pm['log_p_f'] - pm['log_q_f']
    log_w_b = pm['log_p_b'] - pm['log_q_b']
    return log_w_f - log_w_b + ll_f - ll_b


def what_step_weight(q_mu, q_std, z_what, z_what_old, ll_f, ll_b):
    """objectives.py:233-241 -> log_w, log_q_f, density (all (S,B))."""
    log_q_f = dists.normal_log_prob(z_what, q_mu, q_std).sum(-1).sum(-1)
    log_p_f = what_log_prior(z_what)
    log_q_b = dists.normal_log_prob(z_what_old, q_mu, q_std).sum(-1).sum(-1)
    log_p_b = what_log_prior(z_what_old)
    log_w = ll_f.sum(-1) + log_p_f.sum(-1) - log_q_f - (ll_b.sum(-1) + log_p_b.sum(-1) - log_q_b)
    return log_w, log_q_f, ll_f.sum(-1) + log_p_f.sum(-1)


# ------------------------------------------------------------------------------ the sweep
def apg_objective(P, prior, frames, K, result_flags, num_sweeps, mnist_mean, strategy='systematic', record=None):
    """objectives.py:11-58.  frames (S,B,T,FP,FP) S-replicated; mnist_mean (S,B,K,DP,DP).
    Returns {'loss_phi','loss_theta','ess','density'} (mode outputs not restated)."""
    S, B, T, FP, _ = frames.shape
    DP = mnist_mean.shape[-1]
    rs = _Resampler(strategy, S)
    trace = {'loss_phi': [], 'loss_theta': [], 'ess': [], 'density': []}

    # ---- one-shot (objectives.py:109-165)
    z_where, log_p_where, log_q_where = [], None, None
    z_t = None
    for t in range(T):
        pm = propose_one_movement(P, prior, frames[:, :, t], mnist_mean, z_t, None, None, FP, DP)
        z_t = pm['z_where']
        if t == 0:
            log_p_where, log_q_where = pm['log_p_f'], pm['log_q_f']
        # quirk: the reference adds the t = 0 terms twice (objectives.py:124-139)
        log_q_where = log_q_where + pm['log_q_f']
        log_p_where = log_p_where + pm['log_p_f']
        z_where.append(z_t.unsqueeze(2))
    z_where = torch.cat(z_where, 2)
    cropped = frame_to_digit(frames, z_where, FP, DP).view(S, B, T, K, DP * DP)
    q_mu, q_std = enc_digit_params(P['enc_digit'], cropped)
    z_what = Normal(q_mu, q_std).sample()
    log_q_what = dists.normal_log_prob(z_what, q_mu, q_std).sum(-1).sum(-1)
    recon, ll = recon_ll(dec_digit_mean(P['dec_digit'], z_what), z_where, frames)
    log_p = log_p_where + what_log_prior(z_what).sum(-1) + ll.sum(-1)
    log_q = log_q_where + log_q_what
    log_w = log_p - log_q
    w = F.softmax(log_w, 0)
    if result_flags['loss_required']:
        trace['loss_phi'].append((w * (-log_q)).sum(0).mean().unsqueeze(0))
        trace['loss_theta'].append((w * (-ll.sum(-1))).sum(0).mean().unsqueeze(0))
    if result_flags['ess_required']:
        trace['ess'].append((1. / (w ** 2).sum(0)).unsqueeze(0))
    if result_flags['density_required']:
        trace['density'].append(log_p.unsqueeze(0))
    if record is not None:
        record.append(dict(step='oneshot', z_where=z_where, z_what=z_what, q_mu=q_mu, q_std=q_std, ll=ll,
                           log_p_where=log_p_where, log_q_where=log_q_where, log_w=log_w, log_p=log_p, log_q=log_q))
    z_where, z_what = _resample(rs, log_w, z_where, z_what)

    for _ in range(num_sweeps - 1):
        # ---- apg_where (objectives.py:167-225)
        template = dec_digit_mean(P['dec_digit'], z_what)
        z_where_old = z_where
        # quirk: apg_where resamples its LOCAL z_what at every frame but returns only z_where
        # (objectives.py:212,225); apg_what then receives the caller's un-resampled z_what (:45)
        z_what_outer = z_what
        LOSS_phi, LOSS_theta = [], []
        log_prior = 0.0
        z_t = None
        z_new = None
        for t in range(T):
            frame_t = frames[:, :, t]
            pm = propose_one_movement(P, prior, frame_t, template, z_t, z_where_old[:, :, t],
                                      None if t == 0 else z_where_old[:, :, t - 1], FP, DP)
            z_t = pm['z_where']
            if result_flags['density_required']:
                log_prior = log_prior + pm['log_p_f']
            digit = dec_digit_mean(P['dec_digit'], z_what)
            _, ll_f = recon_ll(digit, z_t.unsqueeze(2), frame_t.unsqueeze(2))
            _, ll_b = recon_ll(digit, z_where_old[:, :, t].unsqueeze(2), frame_t.unsqueeze(2))
            log_w = where_step_weight(pm, ll_f.squeeze(-1), ll_b.squeeze(-1))
            w = F.softmax(log_w, 0)
            z_new = z_t.unsqueeze(2) if t == 0 else torch.cat((z_new, z_t.unsqueeze(2)), 2)
            if record is not None:
                record.append(dict(step='where', t=t, pm=pm, z_what=z_what, ll_f=ll_f, ll_b=ll_b, log_w=log_w,
                                   z_old_t=z_where_old[:, :, t], z_old_t_1=None if t == 0 else z_where_old[:, :, t - 1]))
            z_new, z_what = _resample(rs, log_w, z_new, z_what)
            # quirk: the reference keeps proposing frame t+1 from the UN-resampled z_where_t (objectives.py:189)
            if result_flags['loss_required']:
                LOSS_phi.append((w * (-pm['log_q_f'])).sum(0).mean().unsqueeze(-1))
                LOSS_theta.append((w * (-ll_f.squeeze(-1))).sum(0).mean().unsqueeze(-1))
        z_where = z_new
        z_what = z_what_outer
        if result_flags['loss_required']:
            trace['loss_phi'].append(torch.cat(LOSS_phi, -1).sum(-1).unsqueeze(0))
            trace['loss_theta'].append(torch.cat(LOSS_theta, -1).sum(-1).unsqueeze(0))
        if result_flags['density_required']:
            trace['density'].append(log_prior.unsqueeze(0))
        # ---- apg_what (objectives.py:228-257)
        cropped = frame_to_digit(frames, z_where, FP, DP).view(S, B, T, K, DP * DP)
        q_mu, q_std = enc_digit_params(P['enc_digit'], cropped)
        z_what_new = Normal(q_mu, q_std).sample()
        _, ll_f = recon_ll(dec_digit_mean(P['dec_digit'], z_what_new), z_where, frames)
        _, ll_b = recon_ll(dec_digit_mean(P['dec_digit'], z_what), z_where, frames)
        log_w, log_q_f, dens = what_step_weight(q_mu, q_std, z_what_new, z_what, ll_f, ll_b)
        w = F.softmax(log_w, 0)
        if result_flags['loss_required']:
            trace['loss_phi'][-1] = trace['loss_phi'][-1] + (w * (-log_q_f)).sum(0).mean().unsqueeze(0)
            trace['loss_theta'][-1] = trace['loss_theta'][-1] + (w * (-ll_f.sum(-1))).sum(0).mean().unsqueeze(0)
        if result_flags['ess_required']:
            trace['ess'].append((1. / (w ** 2).sum(0)).unsqueeze(0))
        if result_flags['density_required']:
            trace['density'][-1] = trace['density'][-1] + dens.unsqueeze(0)
        if record is not None:
            record.append(dict(step='what', z_where=z_where, q_mu=q_mu, q_std=q_std, z_what=z_what_new,
                               z_what_old=z_what, ll_f=ll_f, ll_b=ll_b, log_w=log_w, log_q_f=log_q_f, dens=dens))
        z_where, z_what = _resample(rs, log_w, z_where, z_what_new)
    out = {}
    for k in ('loss_phi', 'loss_theta', 'ess', 'density'):
        flag = 'loss_required' if k.startswith('loss') else k + '_required'
        if result_flags[flag]:
            out[k] = torch.cat(trace[k], 0)
    return out


# ----------------------------------------------------------------------- synthetic data
def simulate(B, T=10, K=3, FP=96, DP=28, dt=0.3, templates=None):
    """Bouncing trajectories as apgs/bmnist/sim_bmnist.py:48-77 (reflect at +-1); digits are smooth random
    blobs (MNIST itself is not available offline) pasted with the transformer and clamped -> (B,T,FP,FP)."""
    if templates is None:
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, DP), torch.linspace(-1, 1, DP), indexing='ij')
        templates = []
        for k in range(K):
            a, b_ = 0.35 + 0.1 * k, 0.55 - 0.1 * k
            templates.append(torch.exp(-((xx / a) ** 2 + (yy / b_) ** 2) * 2.0).clamp(0, 1))
        templates = torch.stack(templates, 0)
    pos = torch.rand(B, K, 2) * 1.6 - 0.8
    vel = torch.randn(B, K, 2)
    vel = vel / vel.norm(dim=-1, keepdim=True) * dt
    traj = []
    for _ in range(T):
        traj.append(pos.clone())
        pos = pos + vel
        over, under = pos > 1, pos < -1
        pos = torch.where(over, 2 - pos, pos)
        pos = torch.where(under, -2 - pos, pos)
        vel = torch.where(over | under, -vel, vel)
    z = torch.stack(traj, 1).unsqueeze(0)                                   # (1,B,T,K,2)
    digit = templates.unsqueeze(0).unsqueeze(0).expand(1, B, K, DP, DP)
    return torch.clamp(digit_to_frame(digit, z, FP, DP).sum(-3), 0, 1)[0]
