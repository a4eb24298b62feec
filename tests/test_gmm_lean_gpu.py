"""GPU parity of the instantiations the BENCHMARK runs ("What's weak" #1).

bench.py's 10-sweep GMM step repeats, per sweep, the LEAN label-mode z step (`k_gmm_z_fused<..., LEAN, 32, 60>`: labels in,
labels out, only log_w / density requested, H = 32, N = 60), the N = 60 label-mode statistics of the eta step and the fused
draw + eta weights.  The tests of tests/test_gmm_gpu.py reach those kernels mostly through their non-lean variants
(want_q / parts / local), so this file pins the lean launches themselves, at S = 1000:

  * against the oracle (oracle/gmm.py == the reference bit for bit on CPU): labels = inverse CDF of the oracle's q_probs on
    the same uniforms (a label may differ only where u sits on a CDF boundary), log_w / density against
    oracle.gmm.z_step_weights on the kernel's own draw (teacher-forced), 1e-5 relative to the summed magnitudes;
  * bit for bit against the non-lean variant on the same inputs;
  * at the full BASELINE.json shapes (C2: S = 1000, B = 1024) on slices of instances incl. b = 0 and b = B - 1;
  * through apg_objective with the bench flags, compact labels vs the one-hot objective.
"""
import numpy as np
import pytest
import torch

from oracle import gmm as og

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RT = 1e-5
BENCH_FLAGS = {"loss_required": False, "ess_required": True, "mode_required": False, "density_required": True}


class _Gen:
    def __init__(self, K=3, D=2):
        pr = og.make_prior(K, D)
        self.prior_alpha, self.prior_beta = pr["alpha"].to(DEV), pr["beta"].to(DEV)
        self.prior_mu, self.prior_nu, self.prior_pi = pr["mu"].to(DEV), pr["nu"].to(DEV), pr["pi"].to(DEV)
        self.prior_ng = (self.prior_alpha, self.prior_beta, self.prior_mu, self.prior_nu)


def _mlp(P):
    p = P["enc_apg_z"]
    return [p[k].to(DEV) for k in ("pi_log_prob.0.weight", "pi_log_prob.0.bias", "pi_log_prob.2.weight", "pi_log_prob.2.bias")]


def _heads(P):
    p = P["enc_apg_eta"]
    return [p[k].to(DEV) for k in ("gamma.0.weight", "gamma.0.bias", "ob.0.weight", "ob.0.bias")]


def _close(a, b, rtol=RT, atol=0.0, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    atol_t = atol.cpu().double() if torch.is_tensor(atol) else torch.full_like(b, atol)
    err = (a - b).abs() - (atol_t + rtol * b.abs())
    assert float(err.max()) <= 0, "%s: worst excess %g (max abs diff %g)" % (what, float(err.max()), float((a - b).abs().max()))
    return float(((a - b).abs() / (atol_t + rtol * b.abs()).clamp_min(1e-300)).max())


def _pad_labels(lab, LN=64):
    LN = max(LN, (lab.shape[-1] + 15) // 16 * 16)
    out = torch.full(tuple(lab.shape[:2]) + (LN,), 255, dtype=torch.uint8, device=lab.device)
    out[..., :lab.shape[-1]] = lab
    return out


def _latents(S, B, K=3, D=2, N=60, seed=0):
    """a particle population like the one a sweep carries: data from the simulator, eta around the data."""
    torch.manual_seed(seed)
    x = og.simulate(B, N, K, D)
    tau = torch.rand(S, B, K, D) * 3 + 0.2
    mu = torch.randn(S, B, K, D) * 2
    lab_old = torch.randint(0, K, (S, B, N), dtype=torch.uint8)
    u = torch.rand(S, B, N)
    return x, tau, mu, lab_old, u


def _check_z_lean_vs_oracle(P, x, tau, mu, lab_old, u, lab_new, w, K=3):
    """x (b,N,D) ... CPU tensors of a slice of instances; lab_new / w the lean kernel's outputs for the same slice."""
    S = tau.shape[0]
    prior = og.make_prior(K, x.shape[-1])
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    q_ref = og.enc_apg_z_probs(P["enc_apg_z"], xs, tau, mu)
    cdf = torch.cumsum(q_ref, -1)
    lab_ref = (u.unsqueeze(-1) * cdf[..., -1:] >= cdf[..., :-1]).sum(-1)
    bad = lab_new.long() != lab_ref
    assert float(bad.float().mean()) <= 1e-3, "label mismatch fraction %g" % float(bad.float().mean())
    gap = (u.unsqueeze(-1) * cdf[..., -1:] - cdf[..., :-1]).abs().min(-1).values
    assert bool((gap[bad] < 1e-5).all()), "a label differs away from a CDF boundary"
    z_new = torch.nn.functional.one_hot(lab_new.long(), K).float()
    z_old = torch.nn.functional.one_hot(lab_old.long(), K).float()
    ref = og.z_step_weights(xs, tau, mu, q_ref, prior, z_new, z_old, parts=True)
    r1 = _close(w["log_w"], ref[1], rtol=0, atol=RT * (ref[4].abs() + ref[5].abs()), what="lean log_w")
    r2 = _close(w["density"], ref[3].sum(-1), what="lean density")
    return int(bad.sum()), r1, r2


def test_lean_z_step_vs_oracle_and_vs_nonlean_S1000():
    """the launch a sweep repeats: labels in / out, log_w + density only, H = 32, N = 60, S = 1000."""
    from apgs_b200.gmm import kernels
    S, B, N, K, D = 1000, 64, 60, 3, 2
    torch.manual_seed(0)
    P = og.make_params(K, D, 32)
    x, tau, mu, lab_old, u = _latents(S, B, K, D, N, seed=1)
    gen = _Gen(K, D)
    xd, td, md, ud = x.to(DEV), tau.to(DEV), mu.to(DEV), u.to(DEV)
    lab_pad = _pad_labels(lab_old.to(DEV))
    # lean: only log_w / log_q_f / density
    z_lean, q_none, w_lean = kernels.z_step_fused(xd, td, md, _mlp(P), gen, z_old=lab_pad, uniforms=ud)
    assert q_none is None and w_lean.log_w_f is None and w_lean.log_w_local is None
    assert z_lean.dtype == torch.uint8 and tuple(z_lean.shape) == (S, B, 64)
    # non-lean variants of the same step: every optional output requested / one-hot z
    z_full, q_full, w_full = kernels.z_step_fused(xd, td, md, _mlp(P), gen, z_old=lab_pad, uniforms=ud,
                                                  want_q=True, parts=True, local=True)
    assert torch.equal(z_lean[..., :N], z_full[..., :N]), "lean labels != non-lean labels"
    for name in ("log_w", "density", "log_q_f"):
        assert torch.equal(getattr(w_lean, name), getattr(w_full, name)), "lean %s is not bit-identical to the non-lean variant" % name
    z_hot, _, w_hot = kernels.z_step_fused(xd, td, md, _mlp(P), gen, z_old=torch.nn.functional.one_hot(lab_old.long(), K).float().to(DEV),
                                           uniforms=ud)
    assert torch.equal(z_hot.argmax(-1), z_lean[..., :N].long()) and torch.equal(w_hot.log_w, w_lean.log_w)
    # the oracle on 16 of the instances (the (S,b,N,K,H) activations of the reference MLP are 23 MB per instance)
    sel = [0, 1, 2, 3, 4, 5, 6, 7, 31, 32, 33, 59, 60, 61, 62, 63]
    w_sel = {"log_w": w_lean.log_w[:, sel].cpu(), "density": w_lean.density[:, sel].cpu()}
    nbad, r1, r2 = _check_z_lean_vs_oracle(P, x[sel], tau[:, sel], mu[:, sel], lab_old[:, sel], u[:, sel],
                                           z_lean[:, sel, :N].cpu(), w_sel, K)
    print("lean z step S=1000: %d boundary labels of %d; worst log_w err / bound %.3f, density %.3f" % (nbad, S * len(sel) * N, r1, r2))


def _eta_mags(xs, z, post, prior, tau, mu):
    """|ll| + |log p(eta)| + |log q(eta)|: the magnitudes one half of the eta-step weight adds up."""
    return (og.log_likelihood(xs, z, tau, mu, True).abs() + og.eta_log_prior(tau, mu, prior).abs()
            + og.eta_log_q(tau, mu, post).abs())


def test_lean_eta_statistics_and_draw_weights_vs_oracle_S1000():
    """k_gmm_nss_im<..., STATS, 60> in label mode + k_gmm_eta_weights_stats with the fused draw, S = 1000."""
    from apgs_b200.gmm import kernels
    S, B, N, K, D = 1000, 24, 60, 3, 2
    torch.manual_seed(0)
    P = og.make_params(K, D, 32)
    prior = og.make_prior(K, D)
    x, tau_old, mu_old, lab, _ = _latents(S, B, K, D, N, seed=2)
    # centres near the data, as in a sweep (the statistics are centred at the old mu)
    gen = _Gen(K, D)
    lab_pad = _pad_labels(lab.to(DEV))
    *post, stats = kernels.eta_posterior_fused(x.to(DEV), lab_pad, _heads(P), gen, center=mu_old.to(DEV))
    z = torch.nn.functional.one_hot(lab.long(), K).float()
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    ref_post = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, z, prior)
    xp, gamma = og.enc_eta_heads(P["enc_apg_eta"], torch.cat((xs, z), -1))
    g5, x5 = gamma.double().unsqueeze(-1), xp.double().unsqueeze(-2)
    s2a, s3 = (g5 * x5.abs()).sum(2), (g5 * x5 ** 2).sum(2)
    cnt = gamma.double().sum(2).unsqueeze(-1).expand_as(s3).clone()
    cnt[cnt == 0] = 1
    _close(post[0], ref_post[0], what="alpha")
    _close(post[3], ref_post[3], what="nu")
    _close(post[2], ref_post[2], atol=2e-6 * s2a / cnt, what="mu")
    _close(post[1], ref_post[1], atol=4e-6 * (s3.abs() + s2a ** 2 / cnt), what="beta")
    # centred one-hot statistics
    dx = xs.unsqueeze(3) - mu_old.unsqueeze(2)                                    # S,B,N,K,D
    zz = z.unsqueeze(-1)
    ref_stats = torch.cat((z.sum(2).unsqueeze(-1), (zz * dx).sum(2), (zz * dx * dx).sum(2)), -1)
    mag = torch.cat((z.sum(2).unsqueeze(-1), (zz * dx.abs()).sum(2), (zz * dx * dx).sum(2)), -1)
    _close(stats, ref_stats, rtol=0, atol=4e-6 * mag + 1e-6, what="centred statistics")
    # the fused draw + eta weights on the kernel's own posterior parameters
    torch.manual_seed(5)
    g_std = torch._standard_gamma(post[0])
    e_std = torch.randn_like(post[2])
    w, tau, mu = kernels.draw_logw_eta_stats(stats, mu_old.to(DEV), tuple(post), gen, g_std, e_std, tau_old.to(DEV), mu_old.to(DEV), parts=True)
    post_c = [p.cpu() for p in post]
    ref = og.eta_step_weights(xs, z, post_c, prior, tau.cpu(), mu.cpu(), tau_old, mu_old, parts=True)
    # each half is ll + log p(eta) - log q(eta): three sums of O(10^2..10^4) that may cancel (the old eta is far out in
    # the tail of the new proposal), so 1e-5 is relative to the summed magnitudes
    mag_f, mag_b = _eta_mags(xs, z, post_c, prior, tau.cpu(), mu.cpu()), _eta_mags(xs, z, post_c, prior, tau_old, mu_old)
    _close(w.log_w_f, ref[4], rtol=0, atol=RT * mag_f, what="eta log_w_f")
    _close(w.log_w_b, ref[5], rtol=0, atol=RT * mag_b, what="eta log_w_b")
    _close(w.density, ref[2], what="eta density")
    _close(w.log_w, ref[0], rtol=0, atol=RT * (mag_f + mag_b), what="eta log_w")
    # without `parts` (the launch the sweep makes) the same numbers come out
    w2, tau2, mu2 = kernels.draw_logw_eta_stats(stats, mu_old.to(DEV), tuple(post), gen, g_std, e_std, tau_old.to(DEV), mu_old.to(DEV))
    assert torch.equal(tau2, tau) and torch.equal(mu2, mu)
    for name in ("log_w", "density", "log_q_f"):
        assert torch.equal(getattr(w2, name), getattr(w, name)), name


def test_full_shape_C2_slices_vs_oracle():
    """BASELINE.json configs[1]: S = 1000, B = 1024 through the lean launches; instances 0, 511, 1022, 1023 against the oracle."""
    from apgs_b200.gmm import kernels
    S, B, N, K, D = 1000, 1024, 60, 3, 2
    torch.manual_seed(0)
    P = og.make_params(K, D, 32)
    prior = og.make_prior(K, D)
    gen = _Gen(K, D)
    g = torch.Generator(device=DEV).manual_seed(3)
    torch.manual_seed(4)
    x = og.simulate(B, N, K, D)
    xd = x.to(DEV)
    tau = torch.rand(S, B, K, D, device=DEV, generator=g) * 3 + 0.2
    mu = torch.randn(S, B, K, D, device=DEV, generator=g) * 2
    lab_old = torch.randint(0, K, (S, B, 64), dtype=torch.uint8, device=DEV, generator=g)
    u = torch.rand(S, B, N, device=DEV, generator=g)
    z_new, _, w = kernels.z_step_fused(xd, tau, mu, _mlp(P), gen, z_old=lab_old, uniforms=u)
    sel = [0, 511, 1022, 1023]
    w_sel = {"log_w": w.log_w[:, sel].cpu(), "density": w.density[:, sel].cpu()}
    _check_z_lean_vs_oracle(P, x[sel], tau[:, sel].cpu(), mu[:, sel].cpu(), lab_old[:, sel, :N].cpu(), u[:, sel].cpu(),
                            z_new[:, sel, :N].cpu(), w_sel, K)
    # eta step at the full shape on the labels just drawn
    *post, stats = kernels.eta_posterior_fused(xd, z_new, _heads(P), gen, center=mu)
    g_std = torch._standard_gamma(post[0])
    e_std = torch.randn_like(post[2])
    tau_old = tau
    we, tau2, mu2 = kernels.draw_logw_eta_stats(stats, mu, tuple(post), gen, g_std, e_std, tau_old, mu)
    zs = torch.nn.functional.one_hot(z_new[:, sel, :N].long().cpu(), K).float()
    xs = x[sel].unsqueeze(0).repeat(S, 1, 1, 1)
    ref_post = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, zs, prior)
    _close(post[0][:, sel], ref_post[0], what="alpha (C2)")
    _close(post[3][:, sel], ref_post[3], what="nu (C2)")
    post_c = [p[:, sel].cpu() for p in post]
    ref = og.eta_step_weights(xs, zs, post_c, prior, tau2[:, sel].cpu(), mu2[:, sel].cpu(), tau_old[:, sel].cpu(), mu[:, sel].cpu(), parts=True)
    mags = (_eta_mags(xs, zs, post_c, prior, tau2[:, sel].cpu(), mu2[:, sel].cpu())
            + _eta_mags(xs, zs, post_c, prior, tau_old[:, sel].cpu(), mu[:, sel].cpu()))
    _close(we.log_w[:, sel], ref[0], rtol=0, atol=RT * mags, what="eta log_w (C2)")
    _close(we.density[:, sel], ref[2], what="eta density (C2)")
    # resample at the full shape: ancestors of the kernel == oracle ancestors given the kernel's normalised weights
    from apgs_b200 import ops
    from oracle import resampler as orc
    ub = torch.rand(B, device=DEV, generator=g)
    anc, wn, ess, lse = ops.sample_ancestors(log_w=we.log_w, strategy="systematic", uniforms=ub, want_stats=True)
    a_ref = orc.systematic_ancestors(wn[:, sel].cpu().numpy(), ub[sel].cpu().numpy())
    assert np.array_equal(anc[:, sel].cpu().numpy(), a_ref), "ancestors (C2 slices)"
    outs = ops.gather_ancestors([tau2, mu2, z_new], anc)
    for o, v in zip(outs, (tau2, mu2, z_new)):
        ref_g = orc.gather_ancestors(v[:, sel].cpu().numpy(), a_ref)
        assert np.array_equal(o[:, sel].cpu().numpy(), ref_g), "gather (C2 slices)"
    lse_ref = torch.logsumexp(we.log_w[:, sel].cpu().double(), 0)
    _close(lse[sel], lse_ref, rtol=RT, what="lse (C2)")


def _models_from_params(P, K=3, D=2, H=32):
    from apgs_b200.gmm import models as M
    ms = M.init_apg_models(K, D, H, True, DEV)
    for m, key in zip(ms[:3], ("enc_rws_eta", "enc_apg_z", "enc_apg_eta")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})
    return ms


def test_objective_bench_flags_compact_labels_vs_one_hot():
    """apg_objective with the flags bench.py uses (ess + density, no mode, no loss).  With the lean one-shot switched off
    in both runs, carrying z as labels (lean z-step / N=60 statistics launches) or one-hot (their non-lean variants)
    gives identical traces bit for bit; the lean one-shot itself only re-associates one sum (checked to tolerance on
    sweep 0, which is independent of the resampling that follows)."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 1000, 12, 4
    torch.manual_seed(0)
    models = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    out = {}
    for tag, compact, lean1 in (("hot", False, False), ("lab", True, False), ("lean", True, True)):
        O.compact_labels, O.lean_oneshot = compact, lean1
        try:
            torch.manual_seed(5)
            with torch.no_grad():
                out[tag] = O.apg_objective(models, x, BENCH_FLAGS, sweeps, "decomposed", Resampler("systematic", S, True, DEV))
        finally:
            O.compact_labels, O.lean_oneshot = True, True
    for k in ("ess", "density"):
        assert torch.equal(out["lab"][k], out["hot"][k]), k
    d0, d1 = out["lean"]["density"][0], out["hot"]["density"][0]
    _close(d0, d1, rtol=3e-6, atol=1e-4, what="one-shot density, lean vs regular")
    for k in ("ess", "density"):
        assert tuple(out["lean"][k].shape) == tuple(out["hot"][k].shape) and bool(torch.isfinite(out["lean"][k]).all())


# ------------------------------------------------------------------ W1 from the resampling launch, log-marginal
@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_resample_stats_equal_weight_stats_and_log_marginal(strategy):
    """Resampler.resample(want_stats=True): w / ESS / lse come out of the ancestor kernel's own softmax -- bit-identical
    to apgs_weight_stats_f32 on the same log-weights, lse = torch.logsumexp within 1e-5 (north_star: log-marginal)."""
    from apgs_b200 import ops
    from apgs_b200.resampler import Resampler
    S, B = 1000, 96
    torch.manual_seed(7)
    lw = (torch.randn(S, B) * 4).to(DEV)
    lat = [torch.randn(S, B, 3, 2, device=DEV), torch.randint(0, 3, (S, B, 64), dtype=torch.uint8, device=DEV)]
    rs = Resampler(strategy, S, True, DEV)
    u = rs.draw_uniforms(S, B, DEV)
    w0, ess0, lse0 = ops.weight_stats(lw, want_w=True)
    outs, anc, (w1, ess1, lse1) = rs.resample(lat, lw, uniforms=u, want_stats=True, want_index=True)
    assert torch.equal(w0, w1) and torch.equal(ess0, ess1) and torch.equal(lse0, lse1)
    outs2, _, (w2, ess2, lse2) = rs.resample(lat, lw, uniforms=u, want_stats=True, want_w=False)
    assert w2 is None and torch.equal(ess2, ess0) and torch.equal(lse2, lse0)
    plain = rs.resample(lat, lw, uniforms=u)
    for a, b, c in zip(outs, outs2, plain):
        assert torch.equal(a, b) and torch.equal(a, c)
    _close(lse1, torch.logsumexp(lw.cpu().double(), 0), rtol=RT, what="lse")
    _close(ess1, 1. / (torch.softmax(lw.cpu().double(), 0) ** 2).sum(0), rtol=RT, what="ess")


def test_objective_reports_log_marginal_estimate():
    """trace['log_z'][m, b] = logsumexp_s log_w - log S of the step whose ESS is trace['ess'][m, b]."""
    from apgs_b200.gmm import objectives as O, kernels
    from apgs_b200.resampler import Resampler
    from apgs_b200 import dist as adist
    S, B, sweeps = 50, 20, 3
    torch.manual_seed(0)
    models = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    with torch.no_grad():
        tr = O.apg_objective(models, x, BENCH_FLAGS, sweeps, "decomposed", Resampler("systematic", S, True, DEV))
    assert tuple(tr["log_z"].shape) == (sweeps, B) and bool(torch.isfinite(tr["log_z"]).all())
    # ESS and log Z-hat are consistent: ess <= S, and for the one-shot both come from the same log-weights: recompute them
    # through the public step function on the same seed
    torch.manual_seed(9)
    with torch.no_grad():
        trace = {'loss': [], 'ess': [], 'E_tau': [], 'E_mu': [], 'E_z': [], 'density': []}
        log_w, tau, mu, z, trace = O.oneshot(models[0], models[1], models[3], x.unsqueeze(0).expand(S, -1, -1, -1), trace, BENCH_FLAGS)
    ref = torch.logsumexp(log_w.cpu().double(), 0) - np.log(S)
    _close(trace["log_z"][0][0], ref, rtol=RT, what="log_z of the one-shot")
    m = adist.allreduce_metrics(tr)
    _close(m["log_z"], tr["log_z"].double().mean(-1), rtol=1e-12, what="metric means")


# ------------------------------------------------------------------ B-sharded == single
@pytest.mark.parametrize("world", [2, 3])
def test_instance_sharded_kernels_equal_the_full_batch_bit_for_bit(world):
    """every kernel of the sweep run on B-slices (its uniforms / variates sliced from ONE global draw via
    dist.shard_uniforms / shard_instances) reproduces the full-B run bit for bit: z step, eta statistics, draw + eta
    weights, ancestors, gather.  (No op mixes instances, SURVEY.md 8e -- this is the GPU-side check of it.)"""
    from apgs_b200 import ops, dist as adist
    from apgs_b200.gmm import kernels
    S, B, N, K, D = 200, 37, 60, 3, 2
    torch.manual_seed(0)
    P = og.make_params(K, D, 32)
    gen = _Gen(K, D)
    x, tau, mu, lab, u = [t.to(DEV) for t in _latents(S, B, K, D, N, seed=3)]
    lab = _pad_labels(lab)
    g_std = torch._standard_gamma(torch.full((S, B, K, D), 3.0, device=DEV))
    e_std = torch.randn(S, B, K, D, device=DEV)
    ub = torch.rand(B, device=DEV)

    def sweep(x, tau, mu, lab, u, g_std, e_std, ub):
        z_new, _, wz = kernels.z_step_fused(x, tau, mu, _mlp(P), gen, z_old=lab, uniforms=u)
        outs, anc, st = ops.resample([tau, mu, z_new], log_w=wz.log_w, uniforms=ub, want_index=True, want_stats=True)
        tau1, mu1, z1 = outs
        *post, stats = kernels.eta_posterior_fused(x, z1, _heads(P), gen, center=mu1)
        we, tau2, mu2 = kernels.draw_logw_eta_stats(stats, mu1, tuple(post), gen, g_std, e_std, tau1, mu1)
        return dict(z_new=z_new[..., :N], log_w_z=wz.log_w, dens_z=wz.density, anc=anc, ess=st[1], lse=st[2], tau1=tau1, mu1=mu1,
                    z1=z1[..., :N], alpha=post[0], beta=post[1], stats=stats, log_w_eta=we.log_w, tau2=tau2, mu2=mu2)
    full = sweep(x, tau, mu, lab, u, g_std, e_std, ub)
    parts = []
    for r in range(world):
        sl = lambda t, dim: adist.shard_instances(t, r, world, dim=dim).contiguous()      # noqa: E731
        parts.append(sweep(sl(x, 0), sl(tau, 1), sl(mu, 1), sl(lab, 1), sl(u, 1), sl(g_std, 1), sl(e_std, 1),
                           adist.shard_uniforms(ub, r, world).contiguous()))
    for k, v in full.items():
        dim = 0 if v.dim() == 1 else 1
        assert torch.equal(torch.cat([p[k] for p in parts], dim), v), k


# ------------------------------------------------------------------ z step that emits the next eta step's inputs
@pytest.mark.parametrize("S,B,N,K,D,H", [(1000, 16, 60, 3, 2, 32), (37, 5, 41, 3, 2, 16), (9, 3, 30, 2, 2, 10), (33, 4, 60, 4, 2, 32),
                                         (70, 2, 61, 3, 2, 32)])
def test_z_step_with_eta_statistics_equals_the_separate_kernels(S, B, N, K, D, H):
    """apgs_gmm_z_step_fused_labels_eta_f32: labels / weights bit-identical to the plain lean launch, and the emitted
    posterior parameters + centred statistics bit-identical to apgs_gmm_eta_posterior_stats_labels_f32 on those labels
    (and therefore within the oracle tolerances checked above)."""
    from apgs_b200.gmm import kernels
    torch.manual_seed(S + N + K)
    P = og.make_params(K, D, H)
    gen = _Gen(K, D)
    x, tau, mu, lab_old, u = [t.to(DEV) for t in _latents(S, B, K, D, N, seed=S)]
    LN = (N + 15) // 16 * 16
    lab_pad = _pad_labels(lab_old, LN)
    z0, _, w0 = kernels.z_step_fused(x, tau, mu, _mlp(P), gen, z_old=lab_pad, uniforms=u)
    z1, _, w1, post, stats = kernels.z_step_fused(x, tau, mu, _mlp(P), gen, z_old=lab_pad, uniforms=u, eta_heads=_heads(P))
    assert torch.equal(z0[..., :N], z1[..., :N])
    for name in ("log_w", "density", "log_q_f"):
        assert torch.equal(getattr(w0, name), getattr(w1, name)), name
    ref = kernels.eta_posterior_fused(x, z1, _heads(P), gen, center=mu)
    for a, b, name in zip(list(post) + [stats], ref, ("alpha", "beta", "mu", "nu", "stats")):
        assert tuple(a.shape) == tuple(b.shape) and torch.equal(a, b), name
    # and against the oracle directly (posterior parameters of Enc_apg_eta on the one-hot form of the new labels)
    if S <= 100:
        z = torch.nn.functional.one_hot(z1[..., :N].long().cpu(), K).float()
        xs = x.cpu().unsqueeze(0).repeat(S, 1, 1, 1)
        rp = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, z, og.make_prior(K, D))
        _close(post[0], rp[0], what="alpha vs oracle")
        _close(post[3], rp[3], what="nu vs oracle")


def test_objective_with_fused_eta_statistics_is_bit_identical():
    """apg_objective (bench flags): the z step emitting the next eta step's inputs (8 launches of k_gmm_nss_im less per
    10-sweep objective) vs the separate statistics kernel -- same seed, identical traces."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 1000, 10, 5
    torch.manual_seed(0)
    models = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    out, default = {}, O.fuse_eta_stats
    for fuse in (True, False):
        O.fuse_eta_stats = fuse
        try:
            torch.manual_seed(5)
            with torch.no_grad():
                out[fuse] = O.apg_objective(models, x, BENCH_FLAGS, sweeps, "decomposed", Resampler("systematic", S, True, DEV))
        finally:
            O.fuse_eta_stats = default
    for k in ("ess", "density", "log_z"):
        assert torch.equal(out[True][k], out[False][k]), k


def test_z_step_eta_entry_stays_inside_its_buffers():
    """raw C-ABI call with every output embedded in guard bands; ragged tile (S not a multiple of 32), padded label rows."""
    import ctypes as C
    from apgs_b200 import _lib
    from apgs_b200._lib import ptr
    from apgs_b200.gmm.kernels import _prior_struct
    lib = _lib.load()
    S, B, N, K, D, H, LN = 37, 5, 60, 3, 2, 32, 64
    torch.manual_seed(3)
    P = og.make_params(K, D, H)
    G = 2048

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=DEV)
        return buf, buf[G:G + n]
    x, tau, mu, lab_old, u = [t.to(DEV) for t in _latents(S, B, K, D, N, seed=9)]
    lab = _pad_labels(lab_old, LN)
    mlp = [t.contiguous() for t in _mlp(P)]
    heads = [t.contiguous() for t in _heads(P)]
    gen = _Gen(K, D)
    rs, keep = _prior_struct(gen)
    bufs = {"zl": guarded(S * B * LN, torch.uint8, 0xEE), "stats": guarded(S * B * K * (1 + 2 * D), torch.float32, -7.0)}
    for name in ("pa", "pb", "pm", "pn"):
        bufs[name] = guarded(S * B * K * D, torch.float32, -7.0)
    for name in ("log_w", "log_q_f", "density"):
        bufs[name] = guarded(S * B, torch.float32, -7.0)
    o = _lib.GmmOut()
    for name in ("log_w", "log_q_f", "density"):
        setattr(o, name, bufs[name][1].data_ptr())
    po = _lib.GmmPost(*[bufs[n][1].data_ptr() for n in ("pa", "pb", "pm", "pn")])
    rc = lib.apgs_gmm_z_step_fused_labels_eta_f32(ptr(x), ptr(tau), ptr(mu), *[ptr(m) for m in mlp], H, C.byref(rs), ptr(lab), ptr(u),
                                                  S, B, N, K, D, ptr(bufs["zl"][1]), LN, *[ptr(h) for h in heads],
                                                  C.byref(po), ptr(bufs["stats"][1]), C.byref(o), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.apgs_last_error()
    torch.cuda.synchronize()
    for name, (buf, view) in bufs.items():
        fill = 0xEE if buf.dtype == torch.uint8 else -7.0
        assert bool((buf[:G] == fill).all()) and bool((buf[-G:] == fill).all()), name
    zl = bufs["zl"][1].view(S, B, LN)
    assert bool((zl[..., :N] < K).all()) and bool((zl[..., N:] == 0xEE).all())            # pad bytes are never written
    for name in ("pa", "pb", "pm", "pn", "stats", "log_w", "density"):
        assert bool(torch.isfinite(bufs[name][1]).all()) and not bool((bufs[name][1] == -7.0).any()), name
    # error paths: missing outputs / unsupported extras
    rc = lib.apgs_gmm_z_step_fused_labels_eta_f32(ptr(x), ptr(tau), ptr(mu), *[ptr(m) for m in mlp], H, C.byref(rs), ptr(lab), ptr(u),
                                                  S, B, N, K, D, ptr(bufs["zl"][1]), LN, *[ptr(h) for h in heads],
                                                  None, ptr(bufs["stats"][1]), C.byref(o), torch.cuda.current_stream().cuda_stream)
    assert rc != 0


# ------------------------------------------------------------------ eta draw with the variates made in the kernel
def test_in_kernel_eta_draw_distribution_and_weights():
    """apgs_gmm_draw_logw_eta_stats_rng_f32: tau ~ Gamma(alpha, beta) (Marsaglia-Tsang on Philox words, incl. alpha < 1) and
    mu ~ N(q_mu, 1/sqrt(q_nu tau)) (Box-Muller) -- Kolmogorov-Smirnov against the exact laws, moments, reproducibility
    from the seed words, and weights bit-identical to scoring the same draws through apgs_gmm_logw_eta_stats_f32."""
    import scipy.stats as st
    from apgs_b200.gmm import kernels
    S, B, K, D = 4000, 24, 3, 2
    torch.manual_seed(0)
    gen = _Gen(K, D)
    alphas = torch.tensor([0.4, 1.0, 2.5, 7.0, 31.0, 150.0])
    alpha = alphas[torch.arange(K * D) % len(alphas)].view(1, 1, K, D).expand(S, B, K, D).contiguous().to(DEV)
    beta = (torch.rand(1, B, K, D) * 3 + 0.3).expand(S, B, K, D).contiguous().to(DEV)
    q_mu = (torch.randn(1, B, K, D) * 2).expand(S, B, K, D).contiguous().to(DEV)
    q_nu = (torch.rand(1, B, K, D) * 20 + 0.5).expand(S, B, K, D).contiguous().to(DEV)
    post = (alpha, beta, q_mu, q_nu)
    stats = torch.rand(S, B, K, 1 + 2 * D, device=DEV)
    tau_old, mu_old = torch.rand(S, B, K, D, device=DEV) + 0.5, torch.randn(S, B, K, D, device=DEV)
    seed = kernels.philox_seed(DEV)
    w, tau, mu = kernels.draw_logw_eta_stats_rng(stats, mu_old, post, gen, seed, tau_old, mu_old, parts=True)
    w2, tau2, mu2 = kernels.draw_logw_eta_stats_rng(stats, mu_old, post, gen, seed, tau_old, mu_old)
    assert torch.equal(tau, tau2) and torch.equal(mu, mu2) and torch.equal(w.log_w, w2.log_w)       # a function of the seed
    w3, tau3, mu3 = kernels.draw_logw_eta_stats_rng(stats, mu_old, post, gen, kernels.philox_seed(DEV), tau_old, mu_old)
    assert not torch.equal(tau, tau3) and not torch.equal(mu, mu3)
    ref = kernels.logw_eta_stats(stats, mu_old, post, gen, tau, mu, tau_old, mu_old, parts=True)
    for name in ("log_w", "log_q_f", "density", "log_w_f", "log_w_b"):
        assert torch.equal(getattr(w, name), getattr(ref, name)), name
    assert bool(torch.isfinite(tau).all()) and bool((tau > 0).all()) and bool(torch.isfinite(mu).all())
    # laws, per (k, d) element and one instance column (S = 4000 draws each) -- and pooled over instances after scaling
    g = (tau * beta).cpu().double()                                      # standard Gamma(alpha)
    e = ((mu - q_mu) * (q_nu * tau).sqrt()).cpu().double()               # standard normal
    for k in range(K):
        for d in range(D):
            a = float(alpha[0, 0, k, d])
            pooled = g[:, :, k, d].reshape(-1).numpy()
            assert st.kstest(pooled[::3], "gamma", args=(a,)).pvalue > 1e-4, ("gamma KS", a)
            assert abs(pooled.mean() - a) < 6 * (a / pooled.size) ** 0.5 + 1e-9, ("gamma mean", a, pooled.mean())
            assert abs(pooled.var() - a) < 0.05 * a + 12 * a / pooled.size ** 0.5, ("gamma var", a, pooled.var())
            n = e[:, :, k, d].reshape(-1).numpy()
            assert st.kstest(n[::3], "norm").pvalue > 1e-4, "normal KS"
            assert abs(np.corrcoef(n, pooled)[0, 1]) < 0.02, "gamma and normal variates are independent"
    # neighbouring particles / elements are uncorrelated
    flat = e.reshape(S, -1).numpy()
    assert abs(np.corrcoef(flat[:-1].reshape(-1), flat[1:].reshape(-1))[0, 1]) < 0.01
    assert abs(np.corrcoef(flat[:, :-1].reshape(-1), flat[:, 1:].reshape(-1))[0, 1]) < 0.01


def test_objective_with_in_kernel_eta_draw_matches_torch_draws_statistically():
    """rng_in_kernels on / off: same objective, different variate streams -> mean ESS / density / log_z of the last sweep
    agree within Monte-Carlo error; and the CUDA-graph replay of the objective gives fresh draws on every replay."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    from apgs_b200.graph import GraphedObjective
    S, B, sweeps = 100, 400, 4
    torch.manual_seed(0)
    models = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    rs = Resampler("systematic", S, True, DEV)
    out, default = {}, O.rng_in_kernels
    for on in (True, False):
        O.rng_in_kernels = on
        try:
            torch.manual_seed(5)
            with torch.no_grad():
                out[on] = O.apg_objective(models, x, BENCH_FLAGS, sweeps, "decomposed", rs)
        finally:
            O.rng_in_kernels = default
    for k in ("ess", "log_z"):
        a, b = out[True][k][-1].cpu().double(), out[False][k][-1].cpu().double()
        se = (a.var() / B + b.var() / B).sqrt()
        assert abs(float(a.mean() - b.mean())) < 5 * float(se) + 1e-6, (k, float(a.mean()), float(b.mean()), float(se))
    a, b = out[True]["density"][-1].mean(0).cpu().double(), out[False]["density"][-1].mean(0).cpu().double()
    se = (a.var() / B + b.var() / B).sqrt()
    assert abs(float(a.mean() - b.mean())) < 5 * float(se), ("density", float(a.mean()), float(b.mean()), float(se))
    with torch.no_grad():
        gobj = GraphedObjective(lambda xx: O.apg_objective(models, xx, BENCH_FLAGS, sweeps, "decomposed", rs), x, warmup=1)
        t1 = {k: v.clone() for k, v in gobj(x).items() if torch.is_tensor(v)}
        t2 = {k: v.clone() for k, v in gobj(x).items() if torch.is_tensor(v)}
    assert not torch.equal(t1["density"], t2["density"]), "graph replays must draw fresh variates"
    assert bool(torch.isfinite(t2["density"]).all())


def test_objective_on_stream_slices_gives_the_same_trace_shapes_and_statistics():
    """apg_objective_streams: the instance batch in slices on parallel CUDA streams -- instances never interact, so the
    concatenated trace has the shapes of apg_objective and the same statistics (different draws: the slices consume the
    generator in another order)."""
    from apgs_b200.gmm import objectives as O
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 64, 300, 3
    torch.manual_seed(0)
    models = _models_from_params(og.make_params())
    torch.manual_seed(1)
    x = og.simulate(B).to(DEV)
    rs = Resampler("systematic", S, True, DEV)
    flags = dict(BENCH_FLAGS, mode_required=True)
    with torch.no_grad():
        torch.manual_seed(3)
        a = O.apg_objective(models, x, flags, sweeps, "decomposed", rs)
        torch.manual_seed(3)
        b = O.apg_objective_streams(models, x, flags, sweeps, "decomposed", rs, parts=3)
    torch.cuda.synchronize()
    for k in ("ess", "density", "log_z", "E_tau", "E_mu", "E_z"):
        assert tuple(a[k].shape) == tuple(b[k].shape), k
        assert bool(torch.isfinite(b[k]).all()), k
    for k in ("ess", "log_z"):
        u, v = a[k][-1].cpu().double(), b[k][-1].cpu().double()
        se = (u.var() / B + v.var() / B).sqrt()
        assert abs(float(u.mean() - v.mean())) < 5 * float(se) + 1e-6, k
    # the one-shot of the first slice sees the same generator state as the un-sliced run's first instances only in law;
    # E_tau of the one-shot is a function of the data alone -> identical
    assert torch.equal(a["E_tau"][0], b["E_tau"][0])


def test_z_step_in_kernel_uniforms_draw_from_q_and_are_independent():
    """seed= mode of the fused z step (uniforms made inside the kernel by Philox, keyed per launch):
      * deterministic in the seed, different for another seed, and the SAME draw in every variant of the kernel (lean
        labels, non-lean labels, one-hot) -- a draw is a function of (seed, particle, point) only;
      * the draw follows q: with (tau, mu) replicated over S particles the label frequencies over s match q_probs of each
        (b, n) within 4.5 standard errors -- which also rules out uniforms shared between particles;
      * no correlation between neighbouring points / particles."""
    from apgs_b200.gmm import kernels
    S, B, N, K = 512, 6, 60, 3
    P, gen = og.make_params(), _Gen()
    x, tau, mu, lab_old, _ = _latents(S, B, seed=11)
    tau, mu = tau[:1].repeat(S, 1, 1, 1).contiguous(), (mu[:1] * 0.3).repeat(S, 1, 1, 1).contiguous()
    xg, tg, mg, lo = x.to(DEV), tau.to(DEV), mu.to(DEV), _pad_labels(lab_old.to(DEV))
    s1 = torch.tensor([1234567, 0], dtype=torch.int64, device=DEV)
    s2 = torch.tensor([1234568, 0], dtype=torch.int64, device=DEV)
    z_lean, _, w_lean = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=lo, seed=s1)
    z_again, _, w_again = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=lo, seed=s1)
    z_full, q, w_full = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=lo, seed=s1, want_q=True, parts=True, local=True)
    hot_old = torch.nn.functional.one_hot(lab_old.long(), K).float().to(DEV)
    z_hot, _, w_hot = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=hot_old, seed=s1)
    (zs_hot, zs_lab), _, _ = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, also_labels=64, seed=s1)
    z_other, _, _ = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=lo, seed=s2)
    lab = z_lean[..., :N]
    assert torch.equal(z_again[..., :N], lab) and torch.equal(w_again.log_w, w_lean.log_w)   # pad bytes are never written
    assert torch.equal(z_full[..., :N], lab) and torch.equal(w_full.log_w, w_lean.log_w)
    assert torch.equal(z_hot.argmax(-1).to(torch.uint8), lab) and torch.equal(w_hot.log_w, w_lean.log_w)
    assert torch.equal(zs_lab[..., :N], lab) and torch.equal(zs_hot.argmax(-1).to(torch.uint8), lab)
    assert float((z_other[..., :N] != lab).float().mean()) > 0.2
    assert int(lab.max()) < K
    # frequencies over the S particles against q of each point
    hot = torch.nn.functional.one_hot(lab.long(), K).double()
    freq = hot.mean(0)                                           # (B, N, K)
    qn = (q / q.sum(-1, keepdim=True))[0].double()
    se = (qn * (1 - qn) / S).sqrt()
    zscore = ((freq - qn).abs() - 1e-6) / se.clamp_min(1e-4)
    assert float(zscore.max()) < 4.5, float(zscore.max())
    assert float((freq - qn).mean(0).mean(0).abs().max()) < 4.5 * float((qn * (1 - qn)).mean().sqrt()) / (S * B * N) ** 0.5
    # residual correlation between neighbours (points of a particle share a Philox block in fours; particles share a CTA)
    resid = hot - qn.unsqueeze(0)
    for shift_dim in (0, 2):
        a = resid.narrow(shift_dim, 0, resid.shape[shift_dim] - 1)[..., 0]
        b = resid.narrow(shift_dim, 1, resid.shape[shift_dim] - 1)[..., 0]
        corr = float((a * b).mean() / (a.pow(2).mean() * b.pow(2).mean()).sqrt())
        assert abs(corr) < 4.5 / a.numel() ** 0.5, (shift_dim, corr)
    for lag in (8, 32):                                          # lane / pass strides of the point-to-thread mapping
        a, b = resid[:, :, :-lag, 0], resid[:, :, lag:, 0]
        corr = float((a * b).mean() / (a.pow(2).mean() * b.pow(2).mean()).sqrt())
        assert abs(corr) < 4.5 / a.numel() ** 0.5, (lag, corr)


@pytest.mark.parametrize("S,B,N,K,D,H", [(1000, 8, 60, 3, 2, 32), (37, 5, 41, 3, 2, 16), (9, 3, 30, 2, 2, 10), (33, 4, 60, 4, 2, 32),
                                         (21, 3, 17, 3, 3, 7), (40, 2, 100, 5, 2, 32)])
def test_z_step_with_per_instance_tables_is_bit_identical(S, B, N, K, D, H):
    """kernels.z_tables: the data / weight dependent part of Enc_apg_z tabulated once per instance and fetched by the CTAs
    with one bulk copy -- the same numbers as the CTA building them itself, bit for bit, in every variant (lean labels,
    non-lean labels with every optional output, one-hot, sample-only), including instances that take the plain loop."""
    from apgs_b200.gmm import kernels
    torch.manual_seed(5)
    P, gen = og.make_params(K, D, H), _Gen(K, D)
    x, tau, mu, lab_old, _ = _latents(S, B, K, D, N, seed=21)
    if B >= 3:
        mu[:, 1] *= 40.0                                       # saturating pre-activations: these particles take the plain loop
    xg, tg, mg = x.to(DEV), tau.to(DEV), mu.to(DEV)
    lo = _pad_labels(lab_old.to(DEV), LN=(N + 15) // 16 * 16)
    seed = torch.tensor([424242, 0], dtype=torch.int64, device=DEV)
    mlp = _mlp(P)
    tab = kernels.z_tables(xg, mlp, gen)
    assert tab.shape[0] == B and tab.dtype == torch.float32
    for kw in (dict(z_old=lo), dict(z_old=lo, want_q=True, parts=True, local=True),
               dict(z_old=torch.nn.functional.one_hot(lab_old.long(), K).float().to(DEV), want_q=True),
               dict(also_labels=lo.shape[-1])):
        a = kernels.z_step_fused(xg, tg, mg, mlp, gen, seed=seed, **kw)
        b = kernels.z_step_fused(xg, tg, mg, mlp, gen, seed=seed, tables=tab, **kw)
        za, zb = (a[0], b[0]) if not isinstance(a[0], tuple) else (a[0][1], b[0][1])
        assert torch.equal(za[..., :N], zb[..., :N]), kw.keys()
        if a[1] is not None:
            assert torch.equal(a[1], b[1])
        for name in ("log_w", "density", "log_q_f", "log_w_f", "log_w_b", "log_w_local", "log_q_local"):
            va, vb = getattr(a[2], name, None), getattr(b[2], name, None)
            if va is not None:
                assert torch.equal(va, vb), (name, kw.keys())
    # tables built for another shape are refused
    with pytest.raises(ValueError):
        kernels.z_step_fused(xg, tg, mg, mlp, gen, z_old=lo, seed=seed, tables=tab[:, :-4].contiguous())


def test_objective_with_per_instance_tables_is_bit_identical():
    """apg_objective with the bench flags: z tables on / off give the same trace bit for bit (same torch seed)."""
    from apgs_b200.gmm import objectives
    from apgs_b200.resampler import Resampler
    P = og.make_params()
    models = _models_from_params(P)
    x = og.simulate(24, 60, 3, 2).to(DEV)
    out = {}
    for on in (True, False):
        objectives.z_tables_per_objective = on
        try:
            torch.manual_seed(3)
            with torch.no_grad():
                out[on] = objectives.apg_objective(models, x, BENCH_FLAGS, 4, "decomposed", Resampler("systematic", 200, True, DEV))
        finally:
            objectives.z_tables_per_objective = True
    for k in ("ess", "density", "log_z"):
        assert torch.equal(out[True][k], out[False][k]), k


def test_z_loop_probe_runs_the_main_loop_alone():
    """apgs_gmm_z_loop_probe_f32 (bench.py's live ceiling of the MLP part): launches, writes a finite checksum per thread,
    identical on every call; bad arguments are refused before launch."""
    from apgs_b200.gmm import kernels
    a = kernels.z_loop_probe(64, 2, device=DEV)
    b = kernels.z_loop_probe(64, 2, device=DEV)
    assert a.numel() == 64 * 256 and bool(torch.isfinite(a).all()) and torch.equal(a, b)
    assert float(a.abs().max()) > 0
    with pytest.raises(Exception):
        kernels.z_loop_probe(0, 2, device=DEV)


@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_deferred_gather_z_step_through_ancestors_and_composed_gather(strategy):
    """Deferred gather: the z step reading (tau, mu, z_old) through a resampling step's ancestors (src_ancestors) gives the
    same numbers as gathering first, and ops.gather_ancestors(first=...) composes the two index maps -- bit for bit, in the
    lean and the non-lean variants."""
    from apgs_b200 import ops
    from apgs_b200.gmm import kernels
    from apgs_b200.resampler import Resampler
    S, B, N, K = 200, 7, 60, 3
    P, gen = og.make_params(), _Gen()
    x, tau, mu, lab_old, _ = _latents(S, B, seed=31)
    xg, tg, mg, lo = x.to(DEV), tau.to(DEV), mu.to(DEV), _pad_labels(lab_old.to(DEV))
    torch.manual_seed(7)
    rs = Resampler(strategy, S, True, DEV)
    anc1 = rs.ancestors32(torch.randn(S, B, device=DEV) * 2)
    anc2 = rs.ancestors32(torch.randn(S, B, device=DEV) * 2)
    assert anc1.dtype == torch.int32 and int(anc1.min()) >= 0 and int(anc1.max()) < S
    seed = torch.tensor([99, 0], dtype=torch.int64, device=DEV)
    tab = kernels.z_tables(xg, _mlp(P), gen)
    tg1, mg1, lo1 = ops.gather_ancestors([tg, mg, lo], anc1)
    for kw in (dict(), dict(want_q=True, parts=True, local=True)):
        za, qa, wa = kernels.z_step_fused(xg, tg1, mg1, _mlp(P), gen, z_old=lo1, seed=seed, tables=tab, **kw)
        zb, qb, wb = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=lo, seed=seed, tables=tab, src_ancestors=anc1, **kw)
        assert torch.equal(za[..., :N], zb[..., :N])
        assert torch.equal(wa.log_w, wb.log_w) and torch.equal(wa.density, wb.density) and torch.equal(wa.log_q_f, wb.log_q_f)
        if qa is not None:
            assert torch.equal(qa, qb) and torch.equal(wa.log_w_local, wb.log_w_local) and torch.equal(wa.log_w_b, wb.log_w_b)
    # one-hot z_old rows through the ancestors as well
    hot = torch.nn.functional.one_hot(lab_old.long(), K).float().to(DEV)
    (hot1,) = ops.gather_ancestors([hot], anc1)
    zc, _, wc = kernels.z_step_fused(xg, tg1, mg1, _mlp(P), gen, z_old=hot1, seed=seed)
    zd, _, wd = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=hot, seed=seed, src_ancestors=anc1)
    assert torch.equal(zc, zd) and torch.equal(wc.log_w, wd.log_w)
    # composed gather == two gathers
    t2, m2, z2 = ops.gather_ancestors([tg1, mg1, za], anc2)
    t3, m3, z3 = ops.gather_ancestors([tg, mg, za], anc2, first=[anc1, anc1, None])
    assert torch.equal(t2, t3) and torch.equal(m2, m3) and torch.equal(z2, z3)
    with pytest.raises(ValueError):
        ops.gather_ancestors([tg], anc2.long(), first=[anc1])


@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_objective_with_deferred_gather_is_bit_identical(strategy):
    """apg_objective (bench flags): one gather per sweep (deferred) vs two -- the same trace bit for bit under one seed."""
    from apgs_b200.gmm import objectives
    from apgs_b200.resampler import Resampler
    models = _models_from_params(og.make_params())
    x = og.simulate(24, 60, 3, 2).to(DEV)
    out = {}
    for on in (True, False):
        objectives.defer_gather, objectives.gatherless = on, False
        try:
            torch.manual_seed(3)
            with torch.no_grad():
                out[on] = objectives.apg_objective(models, x, BENCH_FLAGS, 4, "decomposed", Resampler(strategy, 200, True, DEV))
        finally:
            objectives.defer_gather, objectives.gatherless = True, True
    for k in ("ess", "density", "log_z"):
        assert torch.equal(out[True][k], out[False][k]), k


def test_eta_step_kernels_through_ancestor_maps_equal_gather_first():
    """apgs_gmm_eta_posterior_stats_labels_via_f32 / apgs_gmm_draw_logw_eta_stats_rng_via_f32 / apgs_compose_ancestors: reading the
    labels, the centre and the old (tau, mu) through ancestor maps gives the numbers of gathering first, bit for bit."""
    from apgs_b200 import ops
    from apgs_b200.gmm import kernels
    from apgs_b200.resampler import Resampler
    S, B, N, K, D = 200, 6, 60, 3, 2
    P, gen = og.make_params(), _Gen()
    x, tau, mu, lab, _ = _latents(S, B, seed=41)
    xg, tg, mg, lg = x.to(DEV), tau.to(DEV), mu.to(DEV), _pad_labels(lab.to(DEV))
    torch.manual_seed(9)
    rs = Resampler("systematic", S, True, DEV)
    zmap = rs.ancestors32(torch.randn(S, B, device=DEV) * 2)
    emap = rs.ancestors32(torch.randn(S, B, device=DEV) * 2)
    comp = ops.compose_ancestors(zmap, emap)
    assert torch.equal(comp.long(), torch.gather(zmap.long(), 0, emap.long()))
    heads = _heads(P)
    (lg1,) = ops.gather_ancestors([lg], zmap)
    tg1, mg1 = ops.gather_ancestors([tg, mg], emap)
    a = kernels.eta_posterior_fused(xg, lg1, heads, gen, center=mg1)
    b = kernels.eta_posterior_fused(xg, lg, heads, gen, center=mg, z_ancestors=zmap, center_ancestors=emap)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    seed = torch.tensor([5, 6], dtype=torch.int64, device=DEV)
    wa, ta, ma = kernels.draw_logw_eta_stats_rng(a[4], mg1, a[:4], gen, seed, tg1, mg1)
    wb, tb, mb = kernels.draw_logw_eta_stats_rng(b[4], mg, b[:4], gen, seed, tg, mg, old_ancestors=emap)
    assert torch.equal(ta, tb) and torch.equal(ma, mb)
    for name in ("log_w", "log_q_f", "density"):
        assert torch.equal(getattr(wa, name), getattr(wb, name)), name
    # z step with separate maps for (tau, mu) and z_old
    zs = torch.tensor([11, 0], dtype=torch.int64, device=DEV)
    z1, _, w1 = kernels.z_step_fused(xg, tg1, mg1, _mlp(P), gen, z_old=lg1, seed=zs)
    z2, _, w2 = kernels.z_step_fused(xg, tg, mg, _mlp(P), gen, z_old=lg, seed=zs, src_ancestors=emap, src_ancestors_z=zmap)
    assert torch.equal(z1[..., :N], z2[..., :N]) and torch.equal(w1.log_w, w2.log_w) and torch.equal(w1.density, w2.density)


@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_gatherless_objective_is_bit_identical(strategy):
    """apg_objective (bench flags): the gatherless sweep, the one-gather-per-sweep sweep and the gather-everything sweep give
    the same trace bit for bit under one seed."""
    from apgs_b200.gmm import objectives
    from apgs_b200.resampler import Resampler
    models = _models_from_params(og.make_params())
    x = og.simulate(24, 60, 3, 2).to(DEV)
    out = {}
    for mode in ("gatherless", "deferred", "plain"):
        objectives.gatherless, objectives.defer_gather = mode == "gatherless", mode != "plain"
        try:
            torch.manual_seed(3)
            with torch.no_grad():
                out[mode] = objectives.apg_objective(models, x, BENCH_FLAGS, 5, "decomposed", Resampler(strategy, 200, True, DEV))
        finally:
            objectives.gatherless, objectives.defer_gather = True, True
    for k in ("ess", "density", "log_z"):
        assert torch.equal(out["gatherless"][k], out["plain"][k]), k
        assert torch.equal(out["deferred"][k], out["plain"][k]), k


def test_r2_entries_stay_inside_their_buffers_and_refuse_bad_arguments():
    """raw C-ABI calls of the round-2 entries (z tables, tab step with ancestor maps, compose, eta `_via` entries) with every
    output embedded in guard bands, ragged shapes (S not a multiple of the tiles, N = 41, H = 16)."""
    import ctypes as C
    from apgs_b200 import _lib
    from apgs_b200._lib import ptr
    from apgs_b200.gmm.kernels import _prior_struct, _post_struct
    lib = _lib.load()
    S, B, N, K, D, H, LN = 37, 5, 41, 3, 2, 16, 48
    torch.manual_seed(13)
    P = og.make_params(K, D, H)
    G = 2048
    st = torch.cuda.current_stream().cuda_stream

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=DEV)
        return buf, buf[G:G + n]

    def intact(bufs):
        torch.cuda.synchronize()
        for name, (buf, view) in bufs.items():
            fill = {torch.uint8: 0xEE, torch.int32: -77}.get(buf.dtype, -7.0)
            assert bool((buf[:G] == fill).all()) and bool((buf[-G:] == fill).all()), name
    x, tau, mu, lab_old, _ = [t.to(DEV) for t in _latents(S, B, K, D, N, seed=19)]
    lab = _pad_labels(lab_old, LN)
    mlp = [t.contiguous() for t in _mlp(P)]
    heads = [t.contiguous() for t in _heads(P)]
    gen = _Gen(K, D)
    rs, keep = _prior_struct(gen)
    amap = torch.randint(0, S, (S, B), dtype=torch.int32, device=DEV)
    bmap = torch.randint(0, S, (S, B), dtype=torch.int32, device=DEV)
    # ---- tables
    nb = lib.apgs_gmm_z_tables_bytes(B, N, H, K, D)
    assert nb > 0 and nb % (16 * B) == 0 and lib.apgs_gmm_z_tables_bytes(0, N, H, K, D) == 0
    bufs = {"tab": guarded(nb // 4, torch.float32, -7.0)}
    assert lib.apgs_gmm_z_tables_f32(ptr(x), *[ptr(m) for m in mlp], H, C.byref(rs), B, N, K, D, ptr(bufs["tab"][1]), st) == 0
    intact(bufs)
    assert bool(torch.isfinite(bufs["tab"][1]).all())
    assert lib.apgs_gmm_z_tables_f32(ptr(x), *[ptr(m) for m in mlp], H, C.byref(rs), B, N, K, D, None, st) < 0
    assert lib.apgs_gmm_z_tables_f32(ptr(x), *[ptr(m) for m in mlp], H, C.byref(rs), B, N, 7, 5, ptr(bufs["tab"][1]), st) < 0
    # ---- compose
    cb = {"comp": guarded(S * B, torch.int32, -77)}
    assert lib.apgs_compose_ancestors(ptr(amap), ptr(bmap), S, B, ptr(cb["comp"][1]), st) == 0
    intact(cb)
    assert torch.equal(cb["comp"][1].view(S, B).long(), torch.gather(amap.long(), 0, bmap.long()))
    assert lib.apgs_compose_ancestors(None, ptr(bmap), S, B, ptr(cb["comp"][1]), st) < 0
    # ---- the tab step with both maps, lean launch
    zb = {"zl": guarded(S * B * LN, torch.uint8, 0xEE)}
    for name in ("log_w", "log_q_f", "density"):
        zb[name] = guarded(S * B, torch.float32, -7.0)
    o = _lib.GmmOut()
    for name in ("log_w", "log_q_f", "density"):
        setattr(o, name, zb[name][1].data_ptr())
    seed = torch.tensor([3, 4], dtype=torch.int64, device=DEV)
    rc = lib.apgs_gmm_z_step_fused_tab_f32(ptr(x), ptr(tau), ptr(mu), *[ptr(m) for m in mlp], H, C.byref(rs), None, ptr(lab), ptr(seed),
                                           S, B, N, K, D, None, ptr(zb["zl"][1]), LN, None, None, None, None, None, None, None,
                                           C.byref(o), ptr(bufs["tab"][1]), ptr(amap), ptr(bmap), st)
    assert rc == 0, lib.apgs_last_error()
    intact(zb)
    zl = zb["zl"][1].view(S, B, LN)
    assert bool((zl[..., :N] < K).all()) and bool((zl[..., N:] == 0xEE).all())
    assert bool(torch.isfinite(zb["log_w"][1]).all()) and not bool((zb["log_w"][1] == -7.0).any())
    rc = lib.apgs_gmm_z_step_fused_tab_f32(ptr(x), ptr(tau), ptr(mu), *[ptr(m) for m in mlp], H, C.byref(rs), None, ptr(lab), None,
                                           S, B, N, K, D, None, ptr(zb["zl"][1]), LN, None, None, None, None, None, None, None,
                                           C.byref(o), ptr(bufs["tab"][1]), ptr(amap), ptr(bmap), st)
    assert rc < 0                                                  # no seed
    # ---- eta entries through maps
    eb = {"stats": guarded(S * B * K * (1 + 2 * D), torch.float32, -7.0)}
    for name in ("pa", "pb", "pm", "pn", "tau", "mu"):
        eb[name] = guarded(S * B * K * D, torch.float32, -7.0)
    for name in ("log_w", "log_q_f", "density"):
        eb[name] = guarded(S * B, torch.float32, -7.0)
    pr = [getattr(gen, "prior_" + n) for n in ("alpha", "beta", "mu", "nu")]
    nbe = lib.apgs_gmm_eta_tables_bytes(B, N, K, D)
    assert nbe > 0 and nbe % (16 * B) == 0 and lib.apgs_gmm_eta_tables_bytes(B, N, 7, 5) == 0
    etab = guarded(nbe // 4, torch.float32, -7.0)
    assert lib.apgs_gmm_eta_tables_f32(ptr(x), *[ptr(h) for h in heads], B, N, K, D, ptr(etab[1]), st) == 0
    intact({"etab": etab})
    assert bool(torch.isfinite(etab[1]).all())
    assert lib.apgs_gmm_eta_tables_f32(ptr(x), *[ptr(h) for h in heads], B, N, K, D, None, st) < 0
    rc = lib.apgs_gmm_eta_posterior_stats_labels_via_f32(ptr(x), ptr(lab), *[ptr(h) for h in heads], S, B, N, K, D,
                                                         *[ptr(v) for v in pr], *[ptr(eb[n][1]) for n in ("pa", "pb", "pm", "pn")],
                                                         ptr(mu), ptr(eb["stats"][1]), LN, ptr(amap), ptr(bmap), ptr(etab[1]), st)
    assert rc == 0, lib.apgs_last_error()
    post = tuple(eb[n][1].view(S, B, K, D) for n in ("pa", "pb", "pm", "pn"))
    ps, keep2 = _post_struct(post)
    o2 = _lib.GmmOut()
    for name in ("log_w", "log_q_f", "density"):
        setattr(o2, name, eb[name][1].data_ptr())
    rc = lib.apgs_gmm_draw_logw_eta_stats_rng_via_f32(ptr(eb["stats"][1]), ptr(mu), C.byref(ps), C.byref(rs), ptr(seed),
                                                      ptr(eb["tau"][1]), ptr(eb["mu"][1]), ptr(tau), ptr(mu), ptr(bmap),
                                                      S, B, K, D, C.byref(o2), st)
    assert rc == 0, lib.apgs_last_error()
    intact(eb)
    for name in ("pa", "pb", "pm", "pn", "stats", "tau", "mu", "log_w", "log_q_f", "density"):
        assert bool(torch.isfinite(eb[name][1]).all()) and not bool((eb[name][1] == -7.0).any()), name


@pytest.mark.parametrize("S,B,N,K,D", [(1000, 6, 60, 3, 2), (37, 5, 41, 3, 2), (70, 3, 17, 2, 2), (33, 4, 60, 4, 2), (21, 3, 30, 3, 1)])
def test_eta_statistics_with_per_instance_tables_are_bit_identical(S, B, N, K, D):
    """kernels.eta_tables: Enc_apg_eta's head factors over (point, label) tabulated once per instance and fetched by the CTAs of
    the statistics kernel with one bulk copy -- same posterior parameters and statistics as the CTA tabulating them, bit for bit;
    through apg_objective the traces are identical as well."""
    from apgs_b200.gmm import kernels, objectives
    from apgs_b200.resampler import Resampler
    torch.manual_seed(S + N)
    P, gen = og.make_params(K, D, 32), _Gen(K, D)
    x, tau, mu, lab, _ = _latents(S, B, K, D, N, seed=51)
    xg, mg = x.to(DEV), mu.to(DEV)
    lg = _pad_labels(lab.to(DEV), LN=(N + 15) // 16 * 16)
    heads = _heads(P)
    tab = kernels.eta_tables(xg, heads)
    assert tab is not None and tab.shape[0] == B
    a = kernels.eta_posterior_fused(xg, lg, heads, gen, center=mg)
    b = kernels.eta_posterior_fused(xg, lg, heads, gen, center=mg, tables=tab)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    with pytest.raises(ValueError):
        kernels.eta_posterior_fused(xg, lg, heads, gen, center=mg, tables=tab[:, :-4].contiguous())
    if (K, D) == (3, 2) and N == 60:
        models = _models_from_params(P)
        xo = og.simulate(16, 60, 3, 2).to(DEV)
        out = {}
        for on in (True, False):
            objectives.z_tables_per_objective = on
            try:
                torch.manual_seed(3)
                with torch.no_grad():
                    out[on] = objectives.apg_objective(models, xo, BENCH_FLAGS, 4, "decomposed", Resampler("systematic", 100, True, DEV))
            finally:
                objectives.z_tables_per_objective = True
        for k in ("ess", "density", "log_z"):
            assert torch.equal(out[True][k], out[False][k]), k
