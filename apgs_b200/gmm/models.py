"""GMM proposals and generative model: same classes, constructor arguments, parameter names
(state_dict keys) and forward() contract as apgs/gmm/models.py, with the arithmetic after the
Linear layers done by the sm_100a kernels.

    Enc_rws_eta   apgs/gmm/models.py:10-48      Enc_apg_eta  :50-88
    Enc_apg_z     apgs/gmm/models.py:91-115     Generative   :117-185

The nn.Linear layers and the random draws stay in PyTorch (north_star).  Each encoder has a
`params(...)` method returning the proposal parameters (what the fused log-weight kernels eat)
and the reference's `forward(...)` returning a Trace of nodes {value, dist, log_prob}.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions.gamma import Gamma
from torch.distributions.normal import Normal
from torch.distributions.one_hot_categorical import OneHotCategorical as cat

from ..rng import normal_sample
from ..trace import Trace
from . import kernels


def _x_inst(ob):
    """(S,B,N,D) S-replicated data (apgs/gmm/apg_training.py:19) or (B,N,D) -> (B,N,D)."""
    return ob[0] if ob.dim() == 4 else ob


def needs_grad(*modules):
    """True when autograd is recording and any parameter of the given encoders requires grad (training)."""
    return torch.is_grad_enabled() and any(p.requires_grad for m in modules for p in m.parameters())


def posterior_eta_torch(xp, gamma, prior_ng):
    """Differentiable PyTorch form of G1 (apgs/gmm/kls_gmm.py:92-119, quirks Q1/Q2 included) used ONLY to build the
    autograd graph of log q in training mode; the forward numbers of the sweep come from the kernels."""
    prior_alpha, prior_beta, prior_mu, prior_nu = prior_ng
    s1 = gamma.sum(2).unsqueeze(-1)
    s2 = torch.einsum('sbnk,sbnd->sbkd', gamma, xp)
    s3 = torch.einsum('sbnk,sbnd->sbkd', gamma, xp ** 2)
    cnt = s1.expand_as(s2)
    cnt = torch.where(cnt == 0.0, torch.ones_like(cnt), cnt)                       # Q1
    x_bar = s2 / cnt
    alpha = prior_alpha + cnt / 2
    nu = prior_nu + cnt
    mu = (prior_mu * prior_nu + s2) / (cnt + prior_nu)
    beta = prior_beta + (s3 - (s2 ** 2) / cnt) / 2. + (cnt * prior_nu / (cnt + prior_nu)) * ((x_bar - prior_nu) ** 2) / 2.   # Q2
    return alpha, beta, mu, nu


def _posterior_from_stats(s1, s2, s3, prior_ng):
    """kls_gmm.py:106-119 from the three neural statistics (s1 (..,K,1), s2 / s3 (..,K,D)); quirks Q1 / Q2 included."""
    prior_alpha, prior_beta, prior_mu, prior_nu = prior_ng
    cnt = s1.expand_as(s2)
    cnt = torch.where(cnt == 0.0, torch.ones_like(cnt), cnt)                       # Q1
    x_bar = s2 / cnt
    alpha = prior_alpha + cnt / 2
    nu = prior_nu + cnt
    mu = (prior_mu * prior_nu + s2) / (cnt + prior_nu)
    beta = prior_beta + (s3 - (s2 ** 2) / cnt) / 2. + (cnt * prior_nu / (cnt + prior_nu)) * ((x_bar - prior_nu) ** 2) / 2.   # Q2
    return alpha, beta, mu, nu


def eta_log_q_torch(post, tau, mu):
    """differentiable log q(tau, mu) summed over (K, D) (apgs/gmm/objectives.py:167)."""
    alpha, beta, mu_q, nu_q = post
    lq = Gamma(alpha, beta, validate_args=False).log_prob(tau).sum(-1).sum(-1)
    return lq + Normal(mu_q, 1. / (nu_q * tau).sqrt(), validate_args=False).log_prob(mu).sum(-1).sum(-1)


class _PriorNG(object):
    def __init__(self, prior_ng):
        self.prior_alpha, self.prior_beta, self.prior_mu, self.prior_nu = prior_ng


def _eta_trace(post, sampled, tau_old, mu_old):
    q_alpha, q_beta, q_mu, q_nu = post
    q = Trace()
    if sampled:
        tau = Gamma(q_alpha, q_beta, validate_args=False).sample()
        q.gamma(q_alpha, q_beta, value=tau, name='precisions')
        mu = normal_sample(q_mu, 1. / (q_nu * tau).sqrt())
        q.normal(q_mu, 1. / (q_nu * tau).sqrt(), value=mu, name='means')
    else:
        q.gamma(q_alpha, q_beta, value=tau_old, name='precisions')
        q.normal(q_mu, 1. / (q_nu * tau_old).sqrt(), value=mu_old, name='means')
    q.post = post            # (alpha, beta, mu, nu): inputs of the fused log-weight kernels
    return q


class Enc_rws_eta(nn.Module):
    """One-shot (i.e. RWS) encoder of {mean, precision}, apgs/gmm/models.py:10-48."""

    def __init__(self, K, D):
        super(self.__class__, self).__init__()
        self.gamma = nn.Sequential(nn.Linear(D, K), nn.Softmax(-1))
        self.ob = nn.Sequential(nn.Linear(D, D))

    fused = True

    def params(self, ob, prior_ng, S=None):
        x = _x_inst(ob)
        S = ob.shape[0] if ob.dim() == 4 else S
        # the heads depend on the data only: evaluate once per instance, not once per particle
        if self.fused and kernels.supported(self.gamma[0].weight.shape[0], x.shape[-1]) & kernels.ETA_FUSED:
            post = kernels.eta_posterior_fused(x, None, (self.gamma[0].weight, self.gamma[0].bias,
                                                         self.ob[0].weight, self.ob[0].bias), _PriorNG(prior_ng))
        else:
            xp, gam = self.ob(x), self.gamma(x)                               # (B,N,D), (B,N,K)
            post = kernels.nss_posterior(xp.unsqueeze(0), gam.unsqueeze(0), _PriorNG(prior_ng))   # (1,B,K,D)
        return tuple(v.expand(S, -1, -1, -1).contiguous() for v in post)

    def params_torch(self, ob, prior_ng, S=None):
        """same parameters through PyTorch ops only (autograd graph for training)."""
        x = _x_inst(ob)
        S = ob.shape[0] if ob.dim() == 4 else S
        post = posterior_eta_torch(self.ob(x).unsqueeze(0), self.gamma(x).unsqueeze(0), prior_ng)
        return tuple(v.expand(S, -1, -1, -1) for v in post)

    def forward(self, ob, prior_ng, sampled=True, tau_old=None, mu_old=None):
        # the reference's forward() is differentiable: when the parameters take part in autograd the Trace is built from
        # the PyTorch-ops parameters, so q['means'].log_prob / q['precisions'].log_prob carry the graph a custom loss or
        # the reference's own objectives.py need; otherwise (evaluation) from the kernel
        post = self.params_torch(ob, prior_ng) if needs_grad(self) else self.params(ob, prior_ng)
        return _eta_trace(post, sampled, tau_old, mu_old)


class Enc_apg_eta(nn.Module):
    """Conditional proposal of {mean, precision} given z, apgs/gmm/models.py:50-88."""

    def __init__(self, K, D):
        super(self.__class__, self).__init__()
        self.gamma = nn.Sequential(nn.Linear(K + D, K), nn.Softmax(-1))
        self.ob = nn.Sequential(nn.Linear(K + D, D))

    def heads(self, ob, z):
        """Linear(cat(ob, z)) without materialising the S-replicated concatenation:
        W [x; z] + b = (W_x x + b) (per instance) + W_z z (per particle)."""
        x = _x_inst(ob)
        D = x.shape[-1]
        Wg, bg = self.gamma[0].weight, self.gamma[0].bias
        Wo, bo = self.ob[0].weight, self.ob[0].bias
        gam = F.softmax(F.linear(z, Wg[:, D:]) + F.linear(x, Wg[:, :D], bg), -1)
        xp = F.linear(z, Wo[:, D:]) + F.linear(x, Wo[:, :D], bo)
        return xp, gam

    fused = True

    def can_fuse(self):
        """the fused heads + statistics kernels are compiled for a fixed list of (K, D) (kernels.supported)."""
        K, D = self.gamma[0].weight.shape[0], self.ob[0].weight.shape[0]
        return bool(self.fused and kernels.supported(K, D) & kernels.ETA_FUSED)

    def head_weights(self):
        """(Wg, bg, Wo, bo) as the fused kernels take them."""
        return (self.gamma[0].weight, self.gamma[0].bias, self.ob[0].weight, self.ob[0].bias)

    def params(self, ob, z, prior_ng, center=None, z_ancestors=None, center_ancestors=None, tables=None):
        """posterior parameters (alpha, beta, mu, nu); with `center` (the old mu) the fused kernel's pass over z also
        returns the centred one-hot statistics of z as a fifth tensor (see kernels.logw_eta_stats).  z_ancestors /
        center_ancestors: read the labels / the centre through ancestor maps (kernels.eta_posterior_fused)."""
        if self.can_fuse():
            return kernels.eta_posterior_fused(ob, z, (self.gamma[0].weight, self.gamma[0].bias,
                                                       self.ob[0].weight, self.ob[0].bias), _PriorNG(prior_ng),
                                               center=center, z_ancestors=z_ancestors, center_ancestors=center_ancestors,
                                               tables=tables)
        xp, gam = self.heads(ob, z)
        return kernels.nss_posterior(xp, gam, _PriorNG(prior_ng))

    def params_torch(self, ob, z, prior_ng):
        """same parameters through PyTorch ops only (autograd graph for training)."""
        xp, gam = self.heads(ob, z)
        return posterior_eta_torch(xp, gam, prior_ng)

    def params_graph(self, ob, z, prior_ng, nss_only=False):
        """the same differentiable parameters with the S-sized work on the kernels: z is one-hot, so the
        heads take only K values per data point -- tabulate {gamma_k, gamma_k x'_d, gamma_k x'_d^2} per (instance, point,
        label) with ordinary PyTorch ops (a (B,N,K,.) tensor: S times less work than heads()), sum the rows the labels
        select (train.LabelTableSum: one kernel forward, one backward) and apply the Normal-Gamma update."""
        from .train import LabelTableSum
        x = _x_inst(ob)
        B, N, D = x.shape
        K = self.gamma[0].weight.shape[0]
        eye = torch.eye(K, dtype=x.dtype, device=x.device)
        inp = torch.cat((x.unsqueeze(2).expand(B, N, K, D), eye.expand(B, N, K, K)), -1)      # cat(x_n, e_l) for every label l
        gam, xp = self.gamma(inp), self.ob(inp)                                              # (B,N,K,K), (B,N,K,D)
        g5 = gam.unsqueeze(-1)
        table = torch.cat((gam, (g5 * xp.unsqueeze(-2)).flatten(-2), (g5 * (xp ** 2).unsqueeze(-2)).flatten(-2)), -1)
        nss = LabelTableSum.apply(table.contiguous(), z)                                     # (S,B,K + 2 K D)
        if nss_only:
            return nss
        S = nss.shape[0]
        s1 = nss[..., :K].unsqueeze(-1)
        s2 = nss[..., K:K + K * D].reshape(S, B, K, D)
        s3 = nss[..., K + K * D:].reshape(S, B, K, D)
        return _posterior_from_stats(s1, s2, s3, prior_ng)

    def log_q_graph(self, ob, z, prior_ng, tau, mu):
        """log q(tau, mu | x, z) (S,B) with a graph to the head weights: head table (PyTorch) -> LabelTableSum -> EtaLogQ
        (one kernel for the Normal-Gamma update + log-densities + Jacobian) when (K, D) is compiled, else PyTorch ops."""
        K, D = tau.shape[-2], tau.shape[-1]
        if kernels.supported(K, D) & kernels.BASE:
            from .train import EtaLogQ
            return EtaLogQ.apply(self.params_graph(ob, z, prior_ng, nss_only=True), tau, mu, *prior_ng)
        return eta_log_q_torch(self.params_graph(ob, z, prior_ng), tau, mu)

    def forward(self, ob, z, prior_ng, sampled=True, tau_old=None, mu_old=None):
        # differentiable like the reference's forward() when the parameters take part in autograd (see Enc_rws_eta)
        post = self.params_torch(ob, z, prior_ng) if needs_grad(self) else self.params(ob, z, prior_ng)
        return _eta_trace(post, sampled, tau_old, mu_old)


class Enc_apg_z(nn.Module):
    """Conditional proposal of the assignments z, apgs/gmm/models.py:91-115."""
    chunk_elems = 1 << 28          # bound on the (s,b,n,k,h) activation held at once (floats)
    fused = True                   # evaluation path: MLP + softmax + draw fused with the weights (csrc/gmm_fused.cu)

    def __init__(self, K, D, num_hidden):
        super(self.__class__, self).__init__()
        self.K = K
        self.pi_log_prob = nn.Sequential(nn.Linear(3 * D, num_hidden), nn.Tanh(), nn.Linear(num_hidden, 1))

    def probs(self, ob, tau, mu):
        """q_probs (S,B,N,K) = softmax_k MLP(cat(x_n, mu_k, tau_k)).  The first Linear is split
        into its x / mu / tau column blocks so that nothing is replicated over N or S."""
        x = _x_inst(ob)
        D = x.shape[-1]
        W1, b1 = self.pi_log_prob[0].weight, self.pi_log_prob[0].bias
        W2, b2 = self.pi_log_prob[2].weight, self.pi_log_prob[2].bias
        hx = F.linear(x, W1[:, :D])                                          # (B,N,H)
        he = F.linear(mu, W1[:, D:2 * D]) + F.linear(tau, W1[:, 2 * D:], b1)  # (S,B,K,H)
        S, B, K, H = he.shape
        N = x.shape[-2]
        step = max(1, int(self.chunk_elems // max(1, B * N * K * H)))
        outs = []
        for s0 in range(0, S, step):
            h = torch.tanh(hx[None, :, :, None, :] + he[s0:s0 + step, :, None, :, :])   # (s,B,N,K,H)
            outs.append(F.linear(h, W2, b2).squeeze(-1))
        logits = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return F.softmax(logits, -1)

    def mlp(self):
        return (self.pi_log_prob[0].weight, self.pi_log_prob[0].bias, self.pi_log_prob[2].weight, self.pi_log_prob[2].bias)

    def can_fuse(self, need_grad=False):
        """the fused kernel is forward-only: use it unless gradients w.r.t. the MLP are needed."""
        if not self.fused:
            return False
        D = self.pi_log_prob[0].weight.shape[1] // 3
        if self.K is not None and not kernels.supported(self.K, D) & kernels.ZFUSED:
            return False                # (K, D) outside the compiled list: PyTorch MLP + the un-fused weight path
        return not (need_grad and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()))

    def sample_fused(self, ob, tau, mu, generative, want_q=True, label_stride=0, seed=None):
        """z ~ q(z | x, tau, mu) and q_probs from the fused kernel (sample-only mode).  With label_stride the same draw
        also comes back as uint8 labels (S,B,label_stride): returns (z, q_probs, labels)."""
        if label_stride:
            (z, lab), q_probs, _ = kernels.z_step_fused(ob, tau, mu, self.mlp(), generative, want_q=want_q,
                                                        also_labels=label_stride, seed=seed)
            return z, q_probs, lab
        z, q_probs, _ = kernels.z_step_fused(ob, tau, mu, self.mlp(), generative, want_q=want_q, seed=seed)
        return z, q_probs

    def forward(self, ob, tau, mu, sampled=True, z_old=None):
        q = Trace()
        q_probs = self.probs(ob, tau, mu)
        if sampled:
            z = cat(q_probs).sample()
            _ = q.variable(cat, probs=q_probs, value=z, name='states')
        else:
            _ = q.variable(cat, probs=q_probs, value=z_old, name='states')
        q.q_probs = q_probs
        return q


class Generative():
    """The generative model of GMM, apgs/gmm/models.py:117-185."""

    def __init__(self, K, D, CUDA, device):
        super().__init__()
        self.K = K
        dev = torch.device(device) if CUDA else torch.device("cpu")
        self.prior_mu = torch.zeros((K, D), device=dev)
        self.prior_nu = torch.ones((K, D), device=dev) * 0.1
        self.prior_alpha = torch.ones((K, D), device=dev) * 2
        self.prior_beta = torch.ones((K, D), device=dev) * 2
        self.prior_pi = torch.ones(K, device=dev) * (1. / K)
        self.prior_ng = (self.prior_alpha, self.prior_beta, self.prior_mu, self.prior_nu)

    def eta_prior(self, q):
        p = Trace()
        p.gamma(self.prior_alpha, self.prior_beta, value=q['precisions'], name='precisions')
        p.normal(self.prior_mu, 1. / (self.prior_nu * p['precisions'].value).sqrt(), value=q['means'], name='means')
        return p

    def z_prior(self, q):
        p = Trace()
        _ = p.variable(cat, probs=self.prior_pi, value=q['states'], name='states')
        return p

    def log_prob(self, ob, z, tau, mu, aggregate=False):
        """log p(x | z, tau, mu): (S,B,N), or (S,B) when aggregate -- via the z-step kernel's
        likelihood path would need q_probs, so the un-fused form is evaluated with torch ops here;
        the sweep objectives never call this (they use the fused kernels)."""
        if ob.dim() == 3:
            ob = ob.unsqueeze(0).expand(z.shape[0], -1, -1, -1)
        sigma = 1. / tau.sqrt()
        labels = z.argmax(-1).unsqueeze(-1).expand(-1, -1, -1, ob.shape[-1])
        ll = Normal(torch.gather(mu, 2, labels), torch.gather(sigma, 2, labels)).log_prob(ob).sum(-1)
        return ll.sum(-1) if aggregate else ll


def init_apg_models(K, D, num_hidden_z, CUDA, device, load_version=None, lr=None):
    """apgs/gmm/apg_training.py:70-99 (same construction order -> same weights under a seed)."""
    enc_rws_eta = Enc_rws_eta(K, D)
    enc_apg_z = Enc_apg_z(K, D, num_hidden_z)
    enc_apg_eta = Enc_apg_eta(K, D)
    generative = Generative(K, D, CUDA, device)
    if CUDA:
        enc_rws_eta.to(device)
        enc_apg_z.to(device)
        enc_apg_eta.to(device)
    if load_version is not None:
        weights = torch.load("weights/cp-%s" % load_version)
        enc_rws_eta.load_state_dict(weights['enc-rws-eta'])
        enc_apg_z.load_state_dict(weights['enc-apg-z'])
        enc_apg_eta.load_state_dict(weights['enc-apg-eta'])
    if lr is not None:
        assert isinstance(lr, float)
        optimizer = torch.optim.Adam(list(enc_rws_eta.parameters()) + list(enc_apg_z.parameters())
                                     + list(enc_apg_eta.parameters()), lr=lr, betas=(0.9, 0.99))
        return (enc_rws_eta, enc_apg_z, enc_apg_eta, generative), optimizer
    for m in (enc_rws_eta, enc_apg_z, enc_apg_eta):
        for p in m.parameters():
            p.requires_grad = False
    return (enc_rws_eta, enc_apg_z, enc_apg_eta, generative)


def save_apg_models(models, save_version):
    """apgs/gmm/apg_training.py:101-115: weights/cp-<version> with keys enc-rws-eta / enc-apg-z / enc-apg-eta."""
    from .. import checkpoint
    return checkpoint.save("gmm", models, save_version)
