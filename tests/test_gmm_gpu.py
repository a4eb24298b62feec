"""GPU parity: GMM reweighting kernels (G1-G4) and the sweep, through the C ABI, against the
oracle and the reference-generated golden fixtures.

Tolerances (north_star: 1e-5 relative, fp32):
  * every log-density sum (log_w_f, log_w_b, log_q_f, density): rtol 1e-5;
  * log_w = log_w_f - log_w_b is a difference of large sums: atol = 1e-5 * (|f| + |b|);
  * posterior beta contains s3 - s2^2/cnt, a cancellation the reference itself evaluates in fp32:
    atol = 4e-6 * (|s3| + s2^2/cnt) on top of rtol 1e-5 (the kernel is checked to be at least as
    close to an fp64 evaluation as the reference is).
"""
import numpy as np
import pytest
import torch

from oracle import gmm as og

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RT = 1e-5


def _t(g, k):
    return torch.from_numpy(g[k])


def _params(g):
    P = {}
    for k in g.files:
        if k.startswith("P/"):
            _, mod, name = k.split("/", 2)
            P.setdefault(mod, {})[name] = torch.from_numpy(g[k])
    return P


class _Gen:
    def __init__(self, K=3, D=2):
        pr = og.make_prior(K, D)
        self.prior_alpha, self.prior_beta = pr["alpha"].to(DEV), pr["beta"].to(DEV)
        self.prior_mu, self.prior_nu, self.prior_pi = pr["mu"].to(DEV), pr["nu"].to(DEV), pr["pi"].to(DEV)
        self.prior_ng = (self.prior_alpha, self.prior_beta, self.prior_mu, self.prior_nu)


def _close(a, b, rtol=RT, atol=0.0, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    atol_t = atol if torch.is_tensor(atol) else torch.full_like(b, atol)
    err = (a - b).abs() - (atol_t.cpu().double() + rtol * b.abs())
    assert float(err.max()) <= 0, "%s: worst excess %g (max abs diff %g)" % (what, float(err.max()), float((a - b).abs().max()))


def _dev(*ts):
    return [t.to(DEV) for t in ts]


def _eta_mag(tau, mu, post):
    """sum over (K,D) of the magnitudes of the terms a Normal-Gamma log-density adds up: the
    1e-5 relative bound applies to this scale, not to the (possibly cancelling) total."""
    from oracle import dists
    al, be, mq, nq = post
    sc = 1. / (nq * tau).sqrt()
    m = (torch.xlogy(al, be).abs() + torch.xlogy(al - 1, tau).abs() + (be * tau).abs() + torch.lgamma(al).abs()
         + ((mu - mq) ** 2 / (2 * sc ** 2)).abs() + sc.log().abs() + dists.LOG_SQRT_2PI)
    return m.sum(-1).sum(-1)


# --------------------------------------------------------------------------------------- G1
def _nss_check(xp, gamma, ref_post):
    from apgs_b200.gmm import kernels
    a, b, m, n = kernels.nss_posterior(xp.to(DEV), gamma.to(DEV), _Gen(gamma.shape[-1], xp.shape[-1]))
    ra, rb, rm, rn = ref_post
    _close(a, ra, what="alpha")
    _close(n, rn, what="nu")
    g5, x5 = gamma.double().unsqueeze(-1), xp.double().unsqueeze(-2)
    s2a, s3 = (g5 * x5.abs()).sum(2), (g5 * x5 ** 2).sum(2)       # magnitudes entering the fp32 sums
    cnt = gamma.double().sum(2).unsqueeze(-1).expand_as(s3).clone()
    cnt[cnt == 0] = 1
    cond = s3.abs() + s2a ** 2 / cnt
    _close(m, rm, atol=2e-6 * s2a / cnt, what="mu")
    _close(b, rb, atol=4e-6 * cond, what="beta")
    # at least as accurate as the reference w.r.t. an fp64 evaluation of the same formula
    pr = {k: v.double() for k, v in og.make_prior(gamma.shape[-1], xp.shape[-1]).items()}
    ea = og.nss_posterior(xp.double(), gamma.double(), pr)
    err_k = (b.cpu().double() - ea[1]).abs().max()
    err_r = (rb.double() - ea[1]).abs().max()
    assert float(err_k) <= float(err_r) * 1.5 + 1e-6 * float(ea[1].abs().max())


def test_G1_nss_posterior_vs_reference_golden(golden):
    g = golden("gmm")
    for tag in ("soft", "hard"):       # "hard" has an empty cluster -> quirk Q1
        _nss_check(_t(g, "g1_xp"), _t(g, "g1_%s_gamma" % tag),
                   [_t(g, "g1_%s_%s" % (tag, n)) for n in ("alpha", "beta", "mu", "nu")])


@pytest.mark.parametrize("S,B,N,K,D", [(7, 5, 60, 3, 2), (33, 31, 17, 4, 2), (1, 1, 3, 3, 3), (64, 16, 200, 4, 2), (3, 2, 61, 5, 2)])
def test_G1_nss_posterior_vs_oracle(S, B, N, K, D):
    torch.manual_seed(S * 7 + N)
    xp = torch.randn(S, B, N, D) * 3 + 2
    gamma = torch.softmax(torch.randn(S, B, N, K) * 3, -1)
    _nss_check(xp, gamma, og.nss_posterior(xp, gamma, og.make_prior(K, D)))


# ------------------------------------------------------------------------------ G2 / G3 / G4
def test_G2_eta_step_vs_reference_golden(golden):
    from apgs_b200.gmm import kernels
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["eta_z"].shape[0]
    x = _t(g, "x")
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    post = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, _t(g, "eta_z"), prior)
    ref = og.eta_step_weights(xs, _t(g, "eta_z"), post, prior, _t(g, "eta_tau"), _t(g, "eta_mu"),
                              _t(g, "eta_tau_old"), _t(g, "eta_mu_old"), parts=True)
    assert torch.equal(ref[0], _t(g, "eta_log_w"))            # oracle == reference (bit exact on CPU)
    w = kernels.logw_eta(x.to(DEV), _t(g, "eta_z").to(DEV), _dev(*post), _Gen(), *_dev(_t(g, "eta_tau"), _t(g, "eta_mu"),
                         _t(g, "eta_tau_old"), _t(g, "eta_mu_old")), parts=True)
    _close(w.log_w_f, ref[4], what="log_w_f")
    _close(w.log_w_b, ref[5], what="log_w_b")
    _close(w.log_q_f, ref[1], rtol=0, atol=RT * _eta_mag(_t(g, "eta_tau"), _t(g, "eta_mu"), post), what="log_q_f")
    _close(w.density, _t(g, "eta_density")[0], what="density")
    _close(w.log_w, _t(g, "eta_log_w"), rtol=0, atol=RT * (ref[4].abs() + ref[5].abs()), what="log_w")
    # the S-replicated layout the reference passes gives the same numbers
    w2 = kernels.logw_eta(xs.to(DEV), _t(g, "eta_z").to(DEV), _dev(*post), _Gen(), *_dev(_t(g, "eta_tau"), _t(g, "eta_mu"),
                          _t(g, "eta_tau_old"), _t(g, "eta_mu_old")))
    assert torch.equal(w2.log_w, w.log_w)


def _heads(P):
    p = P["enc_apg_eta"]
    return [p[k].to(DEV) for k in ("gamma.0.weight", "gamma.0.bias", "ob.0.weight", "ob.0.bias")]


def _stats_ref(x, z, center):
    """centred one-hot statistics of z in fp64: (S,B,K,1+2D)."""
    zz, t = z.double().unsqueeze(-1), x.double().unsqueeze(0).unsqueeze(-2) - center.double().unsqueeze(2)   # (S,B,N,K,D)
    return torch.cat([z.double().sum(2).unsqueeze(-1), (zz * t).sum(2), (zz * t * t).sum(2)], -1)


def test_G2_eta_step_from_statistics_vs_reference_golden(golden):
    """the product's eta step: one pass over z (posterior + centred statistics), weights from the statistics."""
    from apgs_b200.gmm import kernels
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["eta_z"].shape[0]
    x, z, mu_old = _t(g, "x"), _t(g, "eta_z"), _t(g, "eta_mu_old")
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    post = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, z, prior)
    ref = og.eta_step_weights(xs, z, post, prior, _t(g, "eta_tau"), _t(g, "eta_mu"), _t(g, "eta_tau_old"), mu_old, parts=True)
    *post_k, stats = kernels.eta_posterior_fused(x.to(DEV), z.to(DEV), _heads(P), _Gen(), center=mu_old.to(DEV))
    post_plain = kernels.eta_posterior_fused(x.to(DEV), z.to(DEV), _heads(P), _Gen())
    for a, b in zip(post_k, post_plain):
        assert torch.equal(a, b)                                   # the extra output does not change the posterior
    sr = _stats_ref(x, z, mu_old)
    assert torch.equal(stats[..., 0].cpu().double(), sr[..., 0])   # counts are exact
    mag = _stats_ref(x.abs(), z, -mu_old.abs())                    # magnitudes entering the sums
    _close(stats, sr, rtol=0, atol=2e-6 * mag + 1e-30, what="centred statistics")
    w = kernels.logw_eta_stats(stats, mu_old.to(DEV), tuple(post_k), _Gen(), *_dev(_t(g, "eta_tau"), _t(g, "eta_mu"),
                               _t(g, "eta_tau_old"), mu_old), parts=True)
    _close(w.log_w_f, ref[4], what="log_w_f")
    _close(w.log_w_b, ref[5], what="log_w_b")
    _close(w.log_q_f, ref[1], rtol=0, atol=RT * _eta_mag(_t(g, "eta_tau"), _t(g, "eta_mu"), post), what="log_q_f")
    _close(w.density, _t(g, "eta_density")[0], what="density")
    _close(w.log_w, _t(g, "eta_log_w"), rtol=0, atol=RT * (ref[4].abs() + ref[5].abs()), what="log_w")


@pytest.mark.parametrize("S,B,N,K,D,spread,tau_scale", [(5, 3, 60, 3, 2, 2.0, 3.0), (37, 9, 41, 4, 2, 2.0, 3.0), (2, 1, 7, 3, 3, 2.0, 3.0),
                                                       (40, 6, 60, 3, 2, 0.05, 400.0), (33, 5, 200, 4, 2, 10.0, 0.05)])
def test_G2_eta_step_from_statistics_random_shapes(S, B, N, K, D, spread, tau_scale):
    """statistics route vs the oracle's pointwise sums: ragged tiles, other (K,D), and tight clusters far from the
    origin (tau ~ 400, |x| ~ 10), where un-centred statistics would cancel catastrophically."""
    from apgs_b200.gmm import kernels
    torch.manual_seed(S + N + K)
    prior, gen = og.make_prior(K, D), _Gen(K, D)
    lab = torch.randint(0, K, (S, B, N))
    z = torch.nn.functional.one_hot(lab, K).float()
    if S > 4:
        z[1, 0] = torch.nn.functional.one_hot(torch.zeros(N, dtype=torch.long), K).float()    # empty clusters
    centres = torch.randn(B, K, D) * 10
    x = centres[:, torch.randint(0, K, (N,))][torch.arange(B), :, :] + torch.randn(B, N, D) * spread
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    mu_old = centres.unsqueeze(0) + torch.randn(S, B, K, D) * spread
    mu = mu_old + torch.randn(S, B, K, D) * spread * 0.5
    tau, tau_old = (torch.rand(S, B, K, D) + 0.1) * tau_scale, (torch.rand(S, B, K, D) + 0.1) * tau_scale
    post = (torch.rand(S, B, K, D) * 20 + 1, torch.rand(S, B, K, D) * 20 + 0.5, mu_old + torch.randn(S, B, K, D) * spread,
            torch.rand(S, B, K, D) * 20 + 0.1)
    heads = [torch.randn(K, D + K), torch.randn(K), torch.randn(D, D + K), torch.randn(D)]
    out = kernels.eta_posterior_fused(x.to(DEV), z.to(DEV), [h.to(DEV) for h in heads], gen, center=mu_old.to(DEV))
    stats = out[4]
    w = kernels.logw_eta_stats(stats, mu_old.to(DEV), _dev(*post), gen, *_dev(tau, mu, tau_old, mu_old), parts=True)
    # fp64 truth and the magnitudes that are added (|ll| + |log p| + |log q|), as in SURVEY.md H2
    d = lambda t: t.double()
    pr64 = {k: d(v) for k, v in prior.items()}
    p64 = tuple(d(v) for v in post)
    ref = og.eta_step_weights(d(xs), d(z), p64, pr64, d(tau), d(mu), d(tau_old), d(mu_old), parts=True)
    ref32 = og.eta_step_weights(xs, z, post, prior, tau, mu, tau_old, mu_old, parts=True)
    ll_mag = lambda t, m: (d(z).unsqueeze(-1) * (0.5 * d(t).unsqueeze(2) * (d(xs).unsqueeze(-2) - d(m).unsqueeze(2)) ** 2
                                                 + (0.5 * d(t).log().abs() + 0.9189385).unsqueeze(2))).sum((2, 3, 4))
    mag_f = ll_mag(tau, mu) + og.eta_log_prior(d(tau), d(mu), pr64).abs() + og.eta_log_q(d(tau), d(mu), p64).abs()
    mag_b = ll_mag(tau_old, mu_old) + og.eta_log_prior(d(tau_old), d(mu_old), pr64).abs() + og.eta_log_q(d(tau_old), d(mu_old), p64).abs()
    _close(w.log_w_f, ref[4], rtol=0, atol=RT * mag_f, what="eta f (stats)")
    _close(w.log_w_b, ref[5], rtol=0, atol=RT * mag_b, what="eta b (stats)")
    _close(w.log_w, ref[0], rtol=0, atol=RT * (mag_f + mag_b), what="eta log_w (stats)")
    # and at least as close to the fp64 truth as the reference's own fp32 arithmetic is
    err_k = (w.log_w.cpu().double() - ref[0]).abs().max()
    err_r = (ref32[0].double() - ref[0]).abs().max()
    assert float(err_k) <= 2.0 * float(err_r) + 1e-6 * float((mag_f + mag_b).max())


def test_G3_z_step_vs_reference_golden(golden):
    from apgs_b200.gmm import kernels
    from apgs_b200 import ops
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["z_old"].shape[0]
    x = _t(g, "x")
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    q_probs = og.enc_apg_z_probs(P["enc_apg_z"], xs, _t(g, "z_tau"), _t(g, "z_mu"))
    ref = og.z_step_weights(xs, _t(g, "z_tau"), _t(g, "z_mu"), q_probs, prior, _t(g, "z_new"), _t(g, "z_old"), parts=True)
    assert torch.equal(ref[1], _t(g, "z_log_w"))
    w = kernels.logw_z(x.to(DEV), *_dev(_t(g, "z_tau"), _t(g, "z_mu"), q_probs), _Gen(), *_dev(_t(g, "z_new"), _t(g, "z_old")),
                       parts=True, local=True)
    _close(w.log_w_f, ref[4], what="sum log_w_f")
    _close(w.log_w_b, ref[5], what="sum log_w_b")
    _close(w.density, _t(g, "z_density")[0], what="density")
    _close(w.log_q_local, ref[2], atol=1e-6, what="log_q_local")
    # per-point weights: f - b per point, each term O(1..10)
    _close(w.log_w_local, ref[0], rtol=RT, atol=2e-5, what="log_w_local")
    _close(w.log_w, _t(g, "z_log_w"), rtol=0, atol=RT * (ref[4].abs() + ref[5].abs()), what="log_w")
    # W2 + loss term as the reference computes it (objectives.py:217,220)
    w_local, _ = ops.local_weight_stats(w.log_w_local, want_sum=False)
    loss = (w_local * (-w.log_q_local)).sum(0).sum(-1).mean()
    np.testing.assert_allclose(float(loss), float(_t(g, "z_loss")[0]), rtol=2e-5)


def test_G4_oneshot_and_joint_vs_reference_golden(golden):
    from apgs_b200.gmm import kernels
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["os_z"].shape[0]
    x = _t(g, "x")
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    post = og.enc_rws_eta_posterior(P["enc_rws_eta"], xs, prior)
    q_probs = og.enc_apg_z_probs(P["enc_apg_z"], xs, _t(g, "os_tau"), _t(g, "os_mu"))
    lw, lq, lp = og.oneshot_weights(xs, post, prior, _t(g, "os_tau"), _t(g, "os_mu"), q_probs, _t(g, "os_z"))
    w = kernels.logw_oneshot(x.to(DEV), _dev(*post), _Gen(), *_dev(_t(g, "os_tau"), _t(g, "os_mu"), q_probs, _t(g, "os_z")))
    _close(w.density, _t(g, "os_density")[0], what="oneshot log_p")
    _close(w.log_q_f, lq, rtol=RT, atol=RT * _eta_mag(_t(g, "os_tau"), _t(g, "os_mu"), post), what="oneshot log_q")
    _close(w.log_w, _t(g, "os_log_w"), rtol=0, atol=RT * (lp.abs() + lq.abs()), what="oneshot log_w")
    # joint
    post = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, _t(g, "j_z_old"), prior)
    q_probs = og.enc_apg_z_probs(P["enc_apg_z"], xs, _t(g, "j_tau"), _t(g, "j_mu"))
    ref = og.joint_step_weights(xs, post, prior, _t(g, "j_tau"), _t(g, "j_mu"), _t(g, "j_tau_old"), _t(g, "j_mu_old"),
                                q_probs, _t(g, "j_z"), _t(g, "j_z_old"), parts=True)
    assert torch.equal(ref[0], _t(g, "j_log_w"))
    w = kernels.logw_joint(x.to(DEV), _dev(*post), _Gen(), *_dev(_t(g, "j_tau"), _t(g, "j_mu"), _t(g, "j_tau_old"),
                           _t(g, "j_mu_old"), q_probs, _t(g, "j_z"), _t(g, "j_z_old")), parts=True)
    _close(w.log_w_f, ref[3], what="joint log_w_f")
    _close(w.log_w_b, ref[4], what="joint log_w_b")
    _close(w.density, _t(g, "j_density")[0], what="joint density")
    _close(w.log_w, _t(g, "j_log_w"), rtol=0, atol=RT * (ref[3].abs() + ref[4].abs()), what="joint log_w")


@pytest.mark.parametrize("S,B,N,K,D", [(5, 3, 60, 3, 2), (37, 9, 41, 4, 2), (2, 1, 7, 3, 3), (100, 20, 60, 3, 2)])
def test_steps_vs_oracle_random_shapes(S, B, N, K, D):
    """ragged tiles (S*B not a multiple of the 32-particle tile), other (K,D), -inf-free random inputs."""
    from apgs_b200.gmm import kernels
    torch.manual_seed(S + N + K)
    prior = og.make_prior(K, D)
    x = torch.randn(B, N, D) * 2
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    one_hot = lambda: torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    z, z_old = one_hot(), one_hot()
    tau, tau_old = torch.rand(S, B, K, D) * 3 + 0.1, torch.rand(S, B, K, D) * 3 + 0.1
    mu, mu_old = torch.randn(S, B, K, D) * 2, torch.randn(S, B, K, D) * 2
    post = (torch.rand(S, B, K, D) * 20 + 1, torch.rand(S, B, K, D) * 20 + 0.5, torch.randn(S, B, K, D), torch.rand(S, B, K, D) * 20 + 0.1)
    q_probs = torch.softmax(torch.randn(S, B, N, K) * 3, -1)
    q_probs[0, 0, 0] = torch.tensor([1.0] + [0.0] * (K - 1))           # exercises the eps clamp (Q4)
    gen = _Gen(K, D)
    # random proposal parameters make the three summands of log_w_f huge and cancelling: the bound
    # is relative to the magnitudes that are added (|ll| + |log p| + |log q|), as in SURVEY.md H2
    mag_f = (og.log_likelihood(xs, z, tau, mu, True).abs() + og.eta_log_prior(tau, mu, prior).abs()
             + og.eta_log_q(tau, mu, post).abs())
    mag_b = (og.log_likelihood(xs, z, tau_old, mu_old, True).abs() + og.eta_log_prior(tau_old, mu_old, prior).abs()
             + og.eta_log_q(tau_old, mu_old, post).abs())
    ref = og.eta_step_weights(xs, z, post, prior, tau, mu, tau_old, mu_old, parts=True)
    w = kernels.logw_eta(x.to(DEV), z.to(DEV), _dev(*post), gen, *_dev(tau, mu, tau_old, mu_old), parts=True)
    _close(w.log_w_f, ref[4], rtol=0, atol=RT * mag_f, what="eta f")
    _close(w.log_w_b, ref[5], rtol=0, atol=RT * mag_b, what="eta b")
    _close(w.log_w, ref[0], rtol=0, atol=RT * (mag_f + mag_b), what="eta log_w")
    ref = og.z_step_weights(xs, tau, mu, q_probs, prior, z, z_old, parts=True)
    w = kernels.logw_z(x.to(DEV), *_dev(tau, mu, q_probs), gen, *_dev(z, z_old), parts=True, local=True)
    _close(w.log_w_f, ref[4], what="z f")
    _close(w.log_w_b, ref[5], what="z b")
    _close(w.log_w, ref[1], rtol=0, atol=RT * (ref[4].abs() + ref[5].abs()), what="z log_w")
    _close(w.log_w_local, ref[0], rtol=RT, atol=1e-4, what="z local")
    ref = og.joint_step_weights(xs, post, prior, tau, mu, tau_old, mu_old, q_probs, z, z_old, parts=True)
    w = kernels.logw_joint(x.to(DEV), _dev(*post), gen, *_dev(tau, mu, tau_old, mu_old, q_probs, z, z_old), parts=True)
    mag_zf = og.z_log_q(z, q_probs).sum(-1).abs() + og.z_log_prior(z, prior).sum(-1).abs()
    mag_zb = og.z_log_q(z_old, q_probs).sum(-1).abs() + og.z_log_prior(z_old, prior).sum(-1).abs()
    mag_bj = (og.log_likelihood(xs, z_old, tau_old, mu_old, True).abs() + og.eta_log_prior(tau_old, mu_old, prior).abs()
              + og.eta_log_q(tau_old, mu_old, post).abs())
    _close(w.log_w_f, ref[3], rtol=0, atol=RT * (mag_f + mag_zf), what="joint f")
    _close(w.log_w_b, ref[4], rtol=0, atol=RT * (mag_bj + mag_zb), what="joint b")
    lw, lq, lp = og.oneshot_weights(xs, post, prior, tau, mu, q_probs, z)
    w = kernels.logw_oneshot(x.to(DEV), _dev(*post), gen, *_dev(tau, mu, q_probs, z))
    _close(w.density, lp, rtol=0, atol=RT * (mag_f + mag_zf), what="oneshot p")
    _close(w.log_q_f, lq, rtol=RT, atol=RT * (_eta_mag(tau, mu, post) + mag_zf), what="oneshot q")


# ------------------------------------------------------------------------- encoders + sweep
def _models_from_params(P, K=3, D=2, H=32):
    from apgs_b200.gmm import models as M
    ms = M.init_apg_models(K, D, H, True, DEV)
    for m, key in zip(ms[:3], ("enc_rws_eta", "enc_apg_z", "enc_apg_eta")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})     # same state_dict keys as the reference
    return ms


def test_encoder_parameters_match_oracle(golden):
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    enc_rws_eta, enc_apg_z, enc_apg_eta, gen = _models_from_params(P)
    S = g["eta_z"].shape[0]
    x = _t(g, "x")
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    with torch.no_grad():
        post = enc_apg_eta.params(xs.to(DEV), _t(g, "eta_z").to(DEV), gen.prior_ng)
        ref = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, _t(g, "eta_z"), prior)
        for a, b, nm in zip(post, ref, ("alpha", "beta", "mu", "nu")):
            _close(a, b, rtol=2e-5, atol=2e-5, what="enc_apg_eta " + nm)
        post = enc_rws_eta.params(xs.to(DEV), gen.prior_ng)
        ref = og.enc_rws_eta_posterior(P["enc_rws_eta"], xs, prior)
        for a, b, nm in zip(post, ref, ("alpha", "beta", "mu", "nu")):
            _close(a, b, rtol=2e-5, atol=2e-5, what="enc_rws_eta " + nm)
        for fused in (True, False):          # fused heads (csrc) and the PyTorch Linear heads give the same parameters
            enc_apg_eta.fused = enc_rws_eta.fused = fused
            post = enc_apg_eta.params(xs.to(DEV), _t(g, "eta_z").to(DEV), gen.prior_ng)
            for a, b, nm in zip(post, og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, _t(g, "eta_z"), prior), "abmn"):
                _close(a, b, rtol=2e-5, atol=2e-5, what="enc_apg_eta fused=%s %s" % (fused, nm))
            post = enc_rws_eta.params(x.to(DEV), gen.prior_ng, S=S)
            for a, b, nm in zip(post, og.enc_rws_eta_posterior(P["enc_rws_eta"], xs, prior), "abmn"):
                _close(a, b, rtol=2e-5, atol=2e-5, what="enc_rws_eta fused=%s %s" % (fused, nm))
        qp = enc_apg_z.probs(x.to(DEV), _t(g, "z_tau").to(DEV), _t(g, "z_mu").to(DEV))
        _close(qp, og.enc_apg_z_probs(P["enc_apg_z"], xs, _t(g, "z_tau"), _t(g, "z_mu")), rtol=2e-5, atol=1e-7, what="q_probs")
        # reference forward() contract: Trace nodes with value / dist / log_prob
        q = enc_apg_eta(xs.to(DEV), z=_t(g, "eta_z").to(DEV), prior_ng=gen.prior_ng, sampled=False,
                        tau_old=_t(g, "eta_tau").to(DEV), mu_old=_t(g, "eta_mu").to(DEV))
        _close(q["precisions"].dist.concentration, _t(g, "eta_alpha"), rtol=2e-5, what="trace alpha")
        _close(q["means"].dist.scale, _t(g, "eta_scale_q"), rtol=2e-5, what="trace scale")
        assert q["means"].log_prob.shape == (S, x.shape[0], 3, 2)


@pytest.mark.parametrize("block", ["decomposed", "joint"])
def test_sweep_statistics_match_oracle(block):
    """End-to-end: CPU and CUDA generators differ, so the comparison is statistical:
    mean ESS and mean log-joint of the last sweep over many instances, within 5 standard errors."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 40, 300, 4
    torch.manual_seed(0)
    P = og.make_params()
    torch.manual_seed(1)
    x = og.simulate(B)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    torch.manual_seed(2)
    ref = og.apg_objective(P, og.make_prior(), x.repeat(S, 1, 1, 1), flags, sweeps, block, "systematic")
    models = _models_from_params(P)
    torch.manual_seed(3)
    with torch.no_grad():
        tr = O.apg_objective(models, x.to(DEV), flags, sweeps, block, Resampler("systematic", S, True, DEV))
    for k in ("loss", "ess", "density"):
        assert tuple(tr[k].shape) == tuple(ref[k].shape), k
        assert bool(torch.isfinite(tr[k]).all()), k
    for k in ("ess", "density"):
        a = tr[k][-1].cpu().reshape(-1, B).mean(0) if k == "density" else tr[k][-1].cpu()
        b = ref[k][-1].reshape(-1, B).mean(0) if k == "density" else ref[k][-1]
        se = (a.var() / B + b.var() / B).sqrt()
        assert abs(float(a.mean() - b.mean())) < 5 * float(se), (k, float(a.mean()), float(b.mean()), float(se))


def test_full_size_sweep_runs():
    """BASELINE.json configs[1] shape (S=1000, B=1024), 2 sweeps: finite, ESS in [1,S]."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B = 1000, 1024
    torch.manual_seed(0)
    models = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    flags = {"loss_required": False, "ess_required": True, "mode_required": False, "density_required": True}
    with torch.no_grad():
        tr = O.apg_objective(models, x, flags, 2, "decomposed", Resampler("systematic", S, True, DEV))
    assert tuple(tr["ess"].shape) == (2, B) and tuple(tr["density"].shape) == (2, S, B)
    assert bool(torch.isfinite(tr["density"]).all())
    assert float(tr["ess"].min()) >= 1 - 1e-3 and float(tr["ess"].max()) <= S + 1e-1


# ------------------------------------------------------------------ N1: fused z step
def _mlp(P):
    p = P["enc_apg_z"]
    return [p[k].to(DEV) for k in ("pi_log_prob.0.weight", "pi_log_prob.0.bias", "pi_log_prob.2.weight", "pi_log_prob.2.bias")]


def test_N1_fused_z_step_teacher_forced_vs_reference_golden(golden):
    """the fused kernel (MLP + softmax + weights) given the reference's own draw z_new."""
    from apgs_b200.gmm import kernels
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["z_old"].shape[0]
    x = _t(g, "x")
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    q_ref = og.enc_apg_z_probs(P["enc_apg_z"], xs, _t(g, "z_tau"), _t(g, "z_mu"))
    ref = og.z_step_weights(xs, _t(g, "z_tau"), _t(g, "z_mu"), q_ref, prior, _t(g, "z_new"), _t(g, "z_old"), parts=True)
    z_new, q, w = kernels.z_step_fused(x.to(DEV), _t(g, "z_tau").to(DEV), _t(g, "z_mu").to(DEV), _mlp(P), _Gen(),
                                       z_old=_t(g, "z_old").to(DEV), z_forced=_t(g, "z_new").to(DEV),
                                       want_q=True, parts=True, local=True)
    assert torch.equal(z_new.cpu(), _t(g, "z_new"))
    _close(q, q_ref, rtol=2e-5, atol=1e-7, what="q_probs")
    _close(w.log_w_f, ref[4], what="sum log_w_f")
    _close(w.log_w_b, ref[5], what="sum log_w_b")
    _close(w.density, _t(g, "z_density")[0], what="density")
    _close(w.log_q_local, ref[2], rtol=2e-5, atol=2e-6, what="log_q_local")
    _close(w.log_w_local, ref[0], rtol=RT, atol=3e-5, what="log_w_local")
    _close(w.log_w, _t(g, "z_log_w"), rtol=0, atol=RT * (ref[4].abs() + ref[5].abs()), what="log_w")


# H = 10, 7: hidden widths that are not a multiple of four (the fused loop pads its tables to quads of hidden units)
@pytest.mark.parametrize("S,B,N,K,D,H", [(6, 5, 60, 3, 2, 32), (37, 3, 41, 4, 2, 16), (33, 2, 9, 3, 3, 48),
                                         (5, 4, 23, 3, 2, 10), (40, 2, 60, 2, 2, 7)])
def test_N1_fused_sampling_is_inverse_cdf_of_oracle_probs(S, B, N, K, D, H):
    from apgs_b200.gmm import kernels
    torch.manual_seed(S + H)
    P, prior = og.make_params(K, D, H), og.make_prior(K, D)
    x = torch.randn(B, N, D) * 2
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    tau, mu = torch.rand(S, B, K, D) * 3 + 0.1, torch.randn(S, B, K, D) * 2
    u = torch.rand(S, B, N)
    q_ref = og.enc_apg_z_probs(P["enc_apg_z"], xs, tau, mu)
    z_new, q, _ = kernels.z_step_fused(x.to(DEV), tau.to(DEV), mu.to(DEV), _mlp(P), _Gen(K, D), uniforms=u.to(DEV), want_q=True)
    _close(q, q_ref, rtol=2e-5, atol=1e-7, what="q_probs")
    z_new = z_new.cpu()
    assert bool(((z_new == 0) | (z_new == 1)).all()) and bool((z_new.sum(-1) == 1).all()), "one-hot"
    cdf = torch.cumsum(q_ref, -1)
    lab_ref = (u.unsqueeze(-1) * cdf[..., -1:] >= cdf[..., :-1]).sum(-1)
    bad = (z_new.argmax(-1) != lab_ref)
    # a label may differ only where u sits on a CDF boundary (q differs by ~1e-6 relative)
    assert float(bad.float().mean()) <= 1e-3
    gap = (u.unsqueeze(-1) * cdf[..., -1:] - cdf[..., :-1]).abs().min(-1).values
    assert bool((gap[bad] < 1e-5).all())


@pytest.mark.parametrize("xs_scale,w_scale", [(12.0, 4.0), (40.0, 10.0), (300.0, 30.0)])
def test_N1_fused_saturating_preactivations_take_the_plain_loop(xs_scale, w_scale):
    """pre-activations far beyond the tanh saturation point (exponents > 2^30, up to +-inf in the Ex table) must
    fail the range test of the factorised loop, run the plain loop and still reproduce the reference probabilities."""
    from apgs_b200.gmm import kernels
    S, B, N, K, D, H = 35, 4, 60, 3, 2, 32
    torch.manual_seed(int(xs_scale))
    P, prior = og.make_params(K, D, H), og.make_prior(K, D)
    enc = {k: v.clone() for k, v in P["enc_apg_z"].items()}
    enc["pi_log_prob.0.weight"] = enc["pi_log_prob.0.weight"] * w_scale
    x = torch.randn(B, N, D) * xs_scale
    x[1] = torch.randn(N, D) * 0.5                      # one instance stays in the clamp-free regime
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    tau, mu = torch.rand(S, B, K, D) * 3 + 0.1, torch.randn(S, B, K, D) * xs_scale
    mu[:, 1] = torch.randn(S, K, D) * 0.5
    u = torch.rand(S, B, N)
    q_ref = og.enc_apg_z_probs(enc, xs, tau, mu)
    mlp = [enc[k].to(DEV) for k in ("pi_log_prob.0.weight", "pi_log_prob.0.bias", "pi_log_prob.2.weight", "pi_log_prob.2.bias")]
    z_new, q, _ = kernels.z_step_fused(x.to(DEV), tau.to(DEV), mu.to(DEV), mlp, _Gen(K, D), uniforms=u.to(DEV), want_q=True)
    assert bool(torch.isfinite(q).all())
    _close(q, q_ref, rtol=3e-5, atol=2e-7, what="q_probs (saturating)")
    z_new = z_new.cpu()
    assert bool((z_new.sum(-1) == 1).all())


def test_N1_fused_huge_output_weights_take_the_plain_loop():
    """output-layer weights beyond 2^24 would overflow the numerators of the fraction-form loop: range test (d)."""
    from apgs_b200.gmm import kernels
    S, B, N, K, D, H = 9, 3, 60, 3, 2, 32
    torch.manual_seed(4)
    P = og.make_params(K, D, H)
    enc = {k: v.clone() for k, v in P["enc_apg_z"].items()}
    enc["pi_log_prob.2.weight"] = enc["pi_log_prob.2.weight"] * 1e8
    enc["pi_log_prob.2.weight"][0, 3:] = 0.0                       # three live units: logits stay comparable
    x = torch.randn(B, N, D)
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    tau, mu = torch.rand(S, B, K, D) * 3 + 0.1, torch.randn(S, B, K, D)
    u = torch.rand(S, B, N)
    q_ref = og.enc_apg_z_probs(enc, xs, tau, mu)
    mlp = [enc[k].to(DEV) for k in ("pi_log_prob.0.weight", "pi_log_prob.0.bias", "pi_log_prob.2.weight", "pi_log_prob.2.bias")]
    z_new, q, _ = kernels.z_step_fused(x.to(DEV), tau.to(DEV), mu.to(DEV), mlp, _Gen(K, D), uniforms=u.to(DEV), want_q=True)
    assert bool(torch.isfinite(q).all())
    assert bool((z_new.cpu().sum(-1) == 1).all())
    # logits of magnitude 1e8 carry fp32 rounding of ~1e1, so q is one-hot: compare the argmax (a near-tie may flip)
    assert float((q.cpu().argmax(-1) != q_ref.argmax(-1)).float().mean()) <= 2e-3


def test_sweep_statistics_torch_encoder_path():
    """the un-fused Enc_apg_z path (PyTorch MLP + torch categorical draw) stays available and agrees statistically."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 40, 300, 3
    torch.manual_seed(0)
    P = og.make_params()
    torch.manual_seed(1)
    x = og.simulate(B)
    flags = {"loss_required": False, "ess_required": True, "mode_required": True, "density_required": True}
    out = {}
    for fused in (True, False):
        models = _models_from_params(P)
        models[0].fused = models[1].fused = models[2].fused = fused
        torch.manual_seed(3)
        with torch.no_grad():
            out[fused] = O.apg_objective(models, x.to(DEV), flags, sweeps, "decomposed", Resampler("systematic", S, True, DEV))
    for k in ("ess", "density"):
        a = out[True][k][-1].reshape(-1, B).mean(0).cpu()
        b = out[False][k][-1].reshape(-1, B).mean(0).cpu()
        se = (a.var() / B + b.var() / B).sqrt()
        assert abs(float(a.mean() - b.mean())) < 5 * float(se), k
    assert tuple(out[True]["E_z"].shape) == tuple(out[False]["E_z"].shape)


# ------------------------------------------------------------------ N2: training mode
def _train_models(P):
    from apgs_b200.gmm import models as M
    ms, opt = M.init_apg_models(3, 2, 32, True, DEV, lr=1e-3)
    for m, key in zip(ms[:3], ("enc_rws_eta", "enc_apg_z", "enc_apg_eta")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})
    return ms, opt


def test_N2_eta_and_z_step_gradients_match_oracle_autograd():
    """loss.backward() through the drop-in steps gives the gradients autograd gives the oracle (= the reference's
    computation) for the same draws: the kernels supply the detached weights, PyTorch the graph of log q."""
    from apgs_b200.gmm import objectives as O
    S, B, N, K, D = 8, 6, 60, 3, 2
    torch.manual_seed(0)
    P, prior = og.make_params(), og.make_prior()
    x = og.simulate(B)
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    z = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    tau_old, mu_old = torch.rand(S, B, K, D) + 0.5, torch.randn(S, B, K, D)
    (enc_rws_eta, enc_apg_z, enc_apg_eta, gen), _ = _train_models(P)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    # ---- eta step
    trace = {"loss": [], "ess": [], "E_tau": [], "E_mu": [], "E_z": [], "density": []}
    torch.manual_seed(1)
    log_w, tau, mu, trace = O.apg_update_eta(enc_apg_eta, gen, x.to(DEV), z.to(DEV), tau_old.to(DEV), mu_old.to(DEV), trace, flags)
    trace["loss"][0].sum().backward()
    Pg = {k: v.clone().requires_grad_(True) for k, v in P["enc_apg_eta"].items()}
    post = og.enc_apg_eta_posterior(Pg, xs, z, prior)
    lw, lq, _, _ = og.eta_step_weights(xs, z, post, prior, tau.cpu(), mu.cpu(), tau_old, mu_old)
    loss_ref = (torch.softmax(lw.detach(), 0) * (-lq)).sum(0).mean()
    loss_ref.backward()
    np.testing.assert_allclose(float(trace["loss"][0].detach()), float(loss_ref.detach()), rtol=2e-4)
    for name, p in enc_apg_eta.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), Pg[name].grad.numpy(), rtol=2e-3, atol=2e-5 * float(Pg[name].grad.abs().max()) + 1e-7)
    # ---- z step
    trace = {"loss": [torch.zeros(1, device=DEV)], "ess": [], "E_tau": [], "E_mu": [], "E_z": [], "density": [torch.zeros(1, S, B, device=DEV)]}
    log_w, z_new, trace = O.apg_update_z(enc_apg_z, gen, x.to(DEV), tau, mu, z.to(DEV), trace, flags)
    trace["loss"][0].sum().backward()
    Pz = {k: v.clone().requires_grad_(True) for k, v in P["enc_apg_z"].items()}
    q_probs = og.enc_apg_z_probs(Pz, xs, tau.cpu(), mu.cpu())
    lwl, _, lqf, _ = og.z_step_weights(xs, tau.cpu(), mu.cpu(), q_probs, prior, z_new.cpu(), z)
    loss_ref = (torch.softmax(lwl.detach(), 0) * (-lqf)).sum(0).sum(-1).mean()
    loss_ref.backward()
    np.testing.assert_allclose(float(trace["loss"][0].detach()), float(loss_ref.detach()), rtol=2e-4)
    scale = max(float(v.grad.abs().max()) for v in Pz.values())     # the shared output bias has an exactly-zero gradient
    for name, p in enc_apg_z.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), Pz[name].grad.numpy(), rtol=5e-3, atol=1e-4 * scale)


def test_N2_training_loop_runs_like_reference_train():
    """the reference's train() inner loop (apgs/gmm/apg_training.py:17-27) on the drop-in: finite loss, every encoder
    parameter receives a gradient, Adam steps change the weights."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 10, 20, 5                       # BASELINE.json configs[0]
    torch.manual_seed(0)
    P = og.make_params()
    models, opt = _train_models(P)
    torch.manual_seed(1)
    data = og.simulate(B).to(DEV)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    rs = Resampler("systematic", S, True, DEV)
    before = [p.detach().clone() for m in models[:3] for p in m.parameters()]
    for it in range(3):
        opt.zero_grad()
        x = data.repeat(S, 1, 1, 1)
        trace = O.apg_objective(models, x, flags, sweeps, "decomposed", rs)
        loss = trace["loss"].sum()
        loss.backward()
        assert bool(torch.isfinite(loss))
        for m in models[:3]:
            for p in m.parameters():
                assert p.grad is not None and bool(torch.isfinite(p.grad).all())
        opt.step()
    after = [p.detach() for m in models[:3] for p in m.parameters()]
    assert all(not torch.equal(a, b) for a, b in zip(before, after))


# ------------------------------------------------------------ N4: Gibbs / BPG baselines on the kernels
def test_N4_bpg_eta_step_and_gibbs_sweep_arithmetic_vs_oracle():
    from apgs_b200.gmm import objectives as O
    S, B, N, K, D = 9, 4, 60, 3, 2
    torch.manual_seed(0)
    prior = og.make_prior()
    x = og.simulate(B)
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    z = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    tau_old, mu_old = torch.rand(S, B, K, D) + 0.5, torch.randn(S, B, K, D)
    gen = _models_from_params(og.make_params())[3]
    # BPG: weight = likelihood ratio, density = log prior of the proposed eta (objectives.py:308-320)
    trace = {"density": []}
    log_w, tau, mu, trace = O.bpg_update_eta(gen, xs.to(DEV), z.to(DEV), tau_old.to(DEV), mu_old.to(DEV), trace)
    ll_f = og.log_likelihood(xs, z, tau.cpu(), mu.cpu(), True)
    ll_b = og.log_likelihood(xs, z, tau_old, mu_old, True)
    _close(log_w, ll_f - ll_b, rtol=0, atol=RT * (ll_f.abs() + ll_b.abs()), what="bpg log_w")
    _close(trace["density"][0][0], og.eta_log_prior(tau.cpu(), mu.cpu(), prior), rtol=RT, atol=1e-4, what="bpg density")
    # Gibbs: the log joint of the sweep's own draws (objectives.py:249-274)
    trace = {"density": []}
    tau, mu, z_new, trace = O.gibbs_sweep(gen, xs.to(DEV), z.to(DEV), trace)
    assert bool((z_new.sum(-1) == 1).all())
    lj = (og.log_likelihood(xs, z_new.cpu(), tau.cpu(), mu.cpu(), True) + og.eta_log_prior(tau.cpu(), mu.cpu(), prior)
          + og.z_log_prior(z_new.cpu(), prior).sum(-1))
    _close(trace["density"][0][0], lj, rtol=RT, atol=1e-5 * lj.abs().max(), what="gibbs log joint")
    # the responsibilities are the reference's (kls_gmm.py:121-133)
    logits = O.posterior_z(xs.to(DEV), tau, mu, gen.prior_pi)
    _close(logits.exp(), og.posterior_z_logits(xs, tau.cpu(), mu.cpu(), prior).exp(), rtol=1e-4, atol=1e-7, what="posterior_z")


def test_N4_gibbs_and_bpg_objectives_match_reference_traces_statistically(golden):
    """whole baselines against the reference's golden density traces (different generators: Monte-Carlo comparison of
    the per-sweep mean log joint over the same 20 instances)."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    g = golden("gmm")
    S, sweeps = 10, 5
    x = _t(g, "full_x")
    B = x.shape[0]
    models = _models_from_params(_params(g))
    flags = {"loss_required": False, "ess_required": False, "mode_required": False, "density_required": True}
    reps = 12
    with torch.no_grad():
        torch.manual_seed(3)
        gib = torch.stack([O.gibbs_objective(models, x.to(DEV).repeat(S, 1, 1, 1), flags, sweeps)["density"] for _ in range(reps)])
        bpg = torch.stack([O.bpg_objective(models, x.to(DEV).repeat(4 * S, 1, 1, 1), flags, sweeps,
                                           Resampler("systematic", 4 * S, True, DEV))["density"] for _ in range(reps)])
    for name, mine, ref in (("gibbs", gib, _t(g, "gibbs_density")), ("bpg", bpg, _t(g, "bpg_density"))):
        assert tuple(mine.shape[1:]) == tuple(ref.shape), name
        assert bool(torch.isfinite(mine).all()), name
        a = mine.cpu().mean(2).permute(1, 0, 2).reshape(sweeps, -1)          # (sweeps, reps * B): particle means
        b = ref.mean(1)                                                      # (sweeps, B)
        for m in range(sweeps):
            se = (a[m].var() / a.shape[1] + b[m].var() / B).sqrt()
            assert abs(float(a[m].mean() - b[m].mean())) < 6 * float(se) + 1e-3 * abs(float(b[m].mean())), (name, m)


# ------------------------------------------------------------------ compact labels inside a sweep
@pytest.mark.parametrize("S,B,N,K,D,H", [(37, 5, 60, 3, 2, 32), (9, 3, 41, 4, 2, 16), (33, 2, 9, 3, 2, 8), (11, 3, 30, 3, 2, 10)])
def test_label_mode_equals_one_hot_mode_bit_for_bit(S, B, N, K, D, H):
    """the label entries (z as one byte per point) give exactly the outputs of the one-hot entries on the same inputs
    and uniforms: z step (draw, weights, density, q_probs) and eta posterior + statistics."""
    from apgs_b200.gmm import kernels
    torch.manual_seed(S + N)
    P = og.make_params(K, D, H)
    gen = _Gen(K, D)
    x = (torch.randn(B, N, D) * 2).to(DEV)
    tau, mu = (torch.rand(S, B, K, D) * 3 + 0.1).to(DEV), (torch.randn(S, B, K, D) * 2).to(DEV)
    lab_old = torch.randint(0, K, (S, B, N), dtype=torch.uint8, device=DEV)
    z_old = torch.nn.functional.one_hot(lab_old.long(), K).float()
    u = torch.rand(S, B, N, device=DEV)
    mlp = [P["enc_apg_z"][k].to(DEV) for k in ("pi_log_prob.0.weight", "pi_log_prob.0.bias", "pi_log_prob.2.weight", "pi_log_prob.2.bias")]
    z1, q1, w1 = kernels.z_step_fused(x, tau, mu, mlp, gen, z_old=z_old, uniforms=u, want_q=True, parts=True, local=True)
    z2, q2, w2 = kernels.z_step_fused(x, tau, mu, mlp, gen, z_old=lab_old, uniforms=u, want_q=True, parts=True, local=True)
    assert z2.dtype == torch.uint8 and tuple(z2.shape) == (S, B, N)
    assert torch.equal(z1.argmax(-1), z2.long()) and torch.equal(q1, q2)
    for name in ("log_w", "log_q_f", "density", "log_w_f", "log_w_b", "log_w_local", "log_q_local"):
        assert torch.equal(getattr(w1, name), getattr(w2, name)), name
    heads = [P["enc_apg_eta"][k].to(DEV) for k in ("gamma.0.weight", "gamma.0.bias", "ob.0.weight", "ob.0.bias")]
    a = kernels.eta_posterior_fused(x, z_old, heads, gen, center=mu)
    b = kernels.eta_posterior_fused(x, lab_old, heads, gen, center=mu)
    assert len(a) == len(b) == 5
    for t1, t2 in zip(a, b):
        assert torch.equal(t1, t2)
    c = kernels.eta_posterior_fused(x, lab_old, heads, gen)
    assert len(c) == 4 and all(torch.equal(t1, t2) for t1, t2 in zip(a[:4], c))
    # label rows padded to 16 bytes (what the objective carries, so that the gather moves 16-byte vectors)
    LN = (N + 15) // 16 * 16
    lab_pad = torch.full((S, B, LN), 255, dtype=torch.uint8, device=DEV)
    lab_pad[..., :N] = lab_old
    z3, q3, w3 = kernels.z_step_fused(x, tau, mu, mlp, gen, z_old=lab_pad, uniforms=u, want_q=True)
    assert tuple(z3.shape) == (S, B, LN) and torch.equal(z3[..., :N], z2) and torch.equal(q3, q2) and torch.equal(w3.log_w, w2.log_w)
    d = kernels.eta_posterior_fused(x, lab_pad, heads, gen, center=mu)
    assert all(torch.equal(t1, t2) for t1, t2 in zip(b, d))


def test_objective_with_compact_labels_matches_one_hot_objective():
    """same seed, same draws: the whole decomposed objective gives identical traces with z carried as labels or one-hot."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 64, 24, 4
    torch.manual_seed(0)
    models = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    flags = {"loss_required": False, "ess_required": True, "mode_required": True, "density_required": True}
    out = {}
    for compact in (True, False):
        O.compact_labels = compact
        try:
            torch.manual_seed(5)
            with torch.no_grad():
                out[compact] = O.apg_objective(models, x, flags, sweeps, "decomposed", Resampler("systematic", S, True, DEV))
        finally:
            O.compact_labels = True
    for k in ("ess", "density", "E_z", "E_tau", "E_mu"):
        assert torch.equal(out[True][k], out[False][k]), k


def test_lean_oneshot_matches_the_oneshot_weight_kernel():
    """the one-shot of the compact-label objective (labels-only sampler + eta terms from the statistics weight kernel on
    empty statistics) against kernels.logw_oneshot on the SAME draw: same terms, another summation order."""
    from apgs_b200.gmm import kernels
    from apgs_b200.gmm import objectives as O
    S, B, N, K, D = 37, 6, 60, 3, 2
    torch.manual_seed(0)
    enc_rws_eta, enc_apg_z, enc_apg_eta, gen = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    with torch.no_grad():
        post = enc_rws_eta.params(x.unsqueeze(0).expand(S, -1, -1, -1), gen.prior_ng)
        tau, mu = O._sample_eta(post)
        u = torch.rand(S, B, N, device=DEV)
        LN = 64
        (z, lab), q, _ = kernels.z_step_fused(x, tau, mu, enc_apg_z.mlp(), gen, uniforms=u, want_q=True, also_labels=LN)
        ref = kernels.logw_oneshot(x, post, gen, tau, mu, q, z)
        (none, lab2), q2, wz = kernels.z_step_fused(x, tau, mu, enc_apg_z.mlp(), gen, uniforms=u, want_q=False,
                                                   also_labels=LN, want_onehot=False)
        assert none is None and q2 is None and torch.equal(lab[..., :N], lab2[..., :N])
        we = kernels.logw_eta_stats(torch.zeros(S, B, K, 1 + 2 * D, device=DEV), mu, post, gen, tau, mu, tau, mu)
        log_p = wz.density + we.density
        log_q = we.log_q_f + wz.log_q_f
    _close(log_p, ref.density, rtol=3e-6, atol=1e-4, what="log p")
    _close(log_q, ref.log_q_f, rtol=3e-6, atol=1e-4, what="log q")
    _close(log_p - log_q, ref.log_w, rtol=0, atol=3e-6 * (ref.density.abs() + ref.log_q_f.abs()).cpu() + 1e-4, what="log w")
    # and the objective takes this path with the budget_analysis flags: finite traces of the right shape
    flags = {"loss_required": False, "ess_required": True, "mode_required": False, "density_required": True}
    from apgs_b200.resampler import Resampler
    with torch.no_grad():
        tr = O.apg_objective((enc_rws_eta, enc_apg_z, enc_apg_eta, gen), x, flags, 3, "decomposed", Resampler("systematic", S, True, DEV))
    assert tuple(tr["ess"].shape) == (3, B) and bool(torch.isfinite(tr["ess"]).all()) and bool(torch.isfinite(tr["density"]).all())


def test_fused_eta_draw_is_bit_identical_to_torch_sampling():
    """draw_logw_eta_stats forms (tau, mu) from the standard variates with the IEEE operations of
    Gamma(alpha, beta).sample() / Normal(q_mu, 1/sqrt(q_nu tau)).sample(): same seed -> the same draws, bit for bit,
    and the same weights as scoring those draws through logw_eta_stats."""
    from apgs_b200.gmm import kernels, objectives as O
    S, B, N, K, D = 21, 7, 60, 3, 2
    torch.manual_seed(0)
    gen = _Gen(K, D)
    x = (torch.randn(B, N, D) * 2).to(DEV)
    lab = torch.randint(0, K, (S, B, N), dtype=torch.uint8, device=DEV)
    tau_old, mu_old = (torch.rand(S, B, K, D) + 0.5).to(DEV), torch.randn(S, B, K, D).to(DEV)
    P = og.make_params(K, D)
    heads = [P["enc_apg_eta"][k].to(DEV) for k in ("gamma.0.weight", "gamma.0.bias", "ob.0.weight", "ob.0.bias")]
    *post, stats = kernels.eta_posterior_fused(x, lab, heads, gen, center=mu_old)
    post = tuple(post)
    torch.manual_seed(11)
    tau_ref, mu_ref = O._sample_eta(post)
    torch.manual_seed(11)
    g, e = torch._standard_gamma(post[0]), torch.randn_like(post[2])
    w, tau, mu = kernels.draw_logw_eta_stats(stats, mu_old, post, gen, g, e, tau_old, mu_old, parts=True)
    assert torch.equal(tau, tau_ref) and torch.equal(mu, mu_ref)
    w2 = kernels.logw_eta_stats(stats, mu_old, post, gen, tau_ref, mu_ref, tau_old, mu_old, parts=True)
    for name in ("log_w", "log_q_f", "density", "log_w_f", "log_w_b"):
        assert torch.equal(getattr(w, name), getattr(w2, name)), name


def test_gmm_label_kernels_stay_inside_their_buffers():
    """raw C-ABI calls of the two label-mode kernels with outputs embedded in sentinel guard bands (the pool has no
    compute-sanitizer): padded label rows, ragged tiles (S not a multiple of 32), every optional output."""
    import ctypes as C
    from apgs_b200 import _lib
    from apgs_b200._lib import ptr
    lib = _lib.load()
    S, B, N, K, D, H, LN = 37, 5, 60, 3, 2, 32, 64
    torch.manual_seed(3)
    P = og.make_params(K, D, H)
    G = 2048

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=DEV)
        return buf, buf[G:G + n]
    x = (torch.randn(B, N, D) * 2).to(DEV)
    tau, mu = (torch.rand(S, B, K, D) * 3 + 0.1).to(DEV), (torch.randn(S, B, K, D) * 2).to(DEV)
    lab = torch.randint(0, K, (S, B, LN), dtype=torch.uint8, device=DEV)
    u = torch.rand(S, B, N, device=DEV)
    mlp = [P["enc_apg_z"][k].to(DEV).contiguous() for k in ("pi_log_prob.0.weight", "pi_log_prob.0.bias", "pi_log_prob.2.weight", "pi_log_prob.2.bias")]
    gen = _Gen(K, D)
    from apgs_b200.gmm.kernels import _prior_struct
    rs, keep = _prior_struct(gen)
    bufs = {}
    bufs["zl"] = guarded(S * B * LN, torch.uint8, 0xEE)
    bufs["q"] = guarded(S * B * N * K, torch.float32, -7.0)
    for name in ("log_w", "log_q_f", "density", "log_w_f", "log_w_b"):
        bufs[name] = guarded(S * B, torch.float32, -7.0)
    for name in ("log_w_local", "log_q_local"):
        bufs[name] = guarded(S * B * N, torch.float32, -7.0)
    o = _lib.GmmOut()
    for name in ("log_w", "log_q_f", "density", "log_w_f", "log_w_b", "log_w_local", "log_q_local"):
        setattr(o, name, bufs[name][1].data_ptr())
    rc = lib.apgs_gmm_z_step_fused_labels_f32(ptr(x), ptr(tau), ptr(mu), ptr(mlp[0]), ptr(mlp[1]), ptr(mlp[2]), ptr(mlp[3]), H,
                                              C.byref(rs), ptr(lab), ptr(u), S, B, N, K, D, ptr(bufs["zl"][1]), LN, None,
                                              ptr(bufs["q"][1]), C.byref(o), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.apgs_last_error()
    heads = [P["enc_apg_eta"][k].to(DEV).contiguous() for k in ("gamma.0.weight", "gamma.0.bias", "ob.0.weight", "ob.0.bias")]
    pr = [getattr(gen, "prior_" + n) for n in ("alpha", "beta", "mu", "nu")]
    for name in ("pa", "pb", "pm", "pn"):
        bufs[name] = guarded(S * B * K * D, torch.float32, -7.0)
    bufs["stats"] = guarded(S * B * K * (1 + 2 * D), torch.float32, -7.0)
    rc = lib.apgs_gmm_eta_posterior_stats_labels_f32(ptr(x), ptr(lab), *[ptr(h) for h in heads], S, B, N, K, D,
                                                     *[ptr(v) for v in pr], *[ptr(bufs[n][1]) for n in ("pa", "pb", "pm", "pn")],
                                                     ptr(mu), ptr(bufs["stats"][1]), LN, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.apgs_last_error()
    torch.cuda.synchronize()
    for name, (buf, view) in bufs.items():
        fill = 0xEE if buf.dtype == torch.uint8 else -7.0
        assert bool((buf[:G] == fill).all()) and bool((buf[-G:] == fill).all()), name
    zl = bufs["zl"][1].view(S, B, LN)
    assert bool((zl[..., :N] < K).all()) and bool((zl[..., N:] == 0xEE).all())            # pad bytes are never written
    for name in ("q", "log_w", "density", "pa", "pb", "pm", "pn", "stats", "log_w_local"):
        assert bool(torch.isfinite(bufs[name][1]).all()) and not bool((bufs[name][1] == -7.0).all()), name


# ------------------------------------------------------------------ shapes outside the compiled (K, D) list
@pytest.mark.parametrize("K,D", [(6, 2), (3, 4), (7, 3)])
def test_unsupported_cluster_shapes_run_through_the_generic_path(K, D):
    """the reference accepts any --num_clusters / data dimension; kernels are compiled for a fixed list (apgs_gmm_supported).
    Other shapes must run (stock PyTorch CUDA ops, gmm/torch_path.py) and agree with the oracle -- not raise UNSUPPORTED."""
    from apgs_b200.gmm import kernels, objectives as O
    from apgs_b200.resampler import Resampler
    assert kernels.supported(3, 2) & kernels.BASE and kernels.supported(3, 2) & kernels.ZFUSED
    assert not kernels.supported(K, D) & kernels.BASE
    S, B, N, H = 8, 5, 30, 16
    torch.manual_seed(K + D)
    P, prior = og.make_params(K, D, H), og.make_prior(K, D)
    x = og.simulate(B, N, K, D)
    xs = x.repeat(S, 1, 1, 1)
    z = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    z_old = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    post = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, z, prior)
    tau, mu = torch.rand(S, B, K, D) + 0.5, torch.randn(S, B, K, D)
    tau_old, mu_old = torch.rand(S, B, K, D) + 0.5, torch.randn(S, B, K, D)
    gen = _Gen(K, D)
    ref = og.eta_step_weights(xs, z, post, prior, tau, mu, tau_old, mu_old, parts=True)
    w = kernels.logw_eta(x.to(DEV), z.to(DEV), _dev(*post), gen, *_dev(tau, mu, tau_old, mu_old), parts=True)
    _close(w.log_w_f, ref[4], rtol=2e-5, atol=1e-3, what="generic eta log_w_f")
    _close(w.log_w, ref[0], rtol=0, atol=2e-5 * (ref[4].abs() + ref[5].abs()) + 1e-3, what="generic eta log_w")
    q = og.enc_apg_z_probs(P["enc_apg_z"], xs, tau, mu)
    refz = og.z_step_weights(xs, tau, mu, q, prior, z, z_old, parts=True)
    wz = kernels.logw_z(x.to(DEV), *_dev(tau, mu, q), gen, *_dev(z, z_old), parts=True, local=True)
    _close(wz.log_w_local, refz[0], rtol=2e-5, atol=1e-4, what="generic z log_w_local")
    _close(wz.log_w, refz[1], rtol=0, atol=2e-5 * (refz[4].abs() + refz[5].abs()) + 1e-3, what="generic z log_w")
    pk = kernels.nss_posterior(*_dev(*og.enc_eta_heads(P["enc_apg_eta"], torch.cat((xs, z), -1))), gen)
    for a, b, name in zip(pk, post, ("alpha", "beta", "mu", "nu")):
        _close(a, b, rtol=1e-4, atol=1e-4, what="generic posterior " + name)
    # and the whole objective, both blocks
    models = _models_from_params(P, K, D, H)
    assert not models[1].can_fuse() and not models[2].can_fuse()
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    for block in ("decomposed", "joint"):
        with torch.no_grad():
            tr = O.apg_objective(models, x.to(DEV), flags, 3, block, Resampler("systematic", S, True, DEV))
        assert tuple(tr["ess"].shape) == (3, B) and bool(torch.isfinite(tr["density"]).all()) and bool(torch.isfinite(tr["loss"]).all())
    with pytest.raises(Exception):
        kernels.logw_eta(x, z, post, gen, tau, mu, tau_old, mu_old)          # CPU tensors are still rejected


def test_encoder_forward_is_differentiable_like_the_reference():
    """Enc_*_eta.forward() builds its Trace from PyTorch ops when parameters take part in autograd, so a custom
    loss on q[...].log_prob reaches the encoder weights (the kernel path is used only without autograd)."""
    from apgs_b200.gmm import models as M
    K, D, S, B, N = 3, 2, 6, 4, 60
    torch.manual_seed(0)
    (enc_rws_eta, enc_apg_z, enc_apg_eta, gen), opt = M.init_apg_models(K, D, 32, True, DEV, lr=1e-3)
    x = og.simulate(B, N, K, D).to(DEV).repeat(S, 1, 1, 1)
    z = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N), device=DEV), K).float()
    q = enc_apg_eta(x, z, gen.prior_ng)
    loss = -(q["means"].log_prob.sum() + q["precisions"].log_prob.sum())
    loss.backward()
    g = [p.grad for p in enc_apg_eta.parameters()]
    assert all(t is not None and bool(torch.isfinite(t).all()) and float(t.abs().sum()) > 0 for t in g)
    q0 = enc_rws_eta(x, gen.prior_ng)
    (-(q0["means"].log_prob.sum())).backward()
    assert all(p.grad is not None for p in enc_rws_eta.parameters())
    with torch.no_grad():                                         # evaluation: the kernel path, same numbers
        qe = enc_apg_eta(x, z, gen.prior_ng, sampled=False, tau_old=q["precisions"].value, mu_old=q["means"].value)
    _close(qe["means"].log_prob, q["means"].log_prob, rtol=1e-4, atol=1e-4, what="kernel vs torch log q(mu)")
