"""Teacher-forced parity at the FULL BASELINE.json shapes: the kernels run at configs[2] (DMM K=4,
N=200, S=100, B=256) and configs[3] (bmnist K=3, T=10, 96x96, S=10, B=64) and a few instances -- always including b = 0
and b = B-1 -- are checked against the oracle (== the reference bit for bit on CPU).  configs[1] (GMM S=1000, B=1024) is
in tests/test_gmm_lean_gpu.py.  Plus the BPG baselines of the two tasks and the product simulators on the GPU (row N4).
Tolerance: 1e-5 relative to the magnitudes that are summed, as in the small-shape tests."""
import numpy as np
import pytest
import torch

from oracle import bmnist as ob
from oracle import dmm as od

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RT = 1e-5
FLAGS = {"loss_required": False, "ess_required": True, "mode_required": False, "density_required": True}


def _close(a, b, atol, what):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    atol = atol.cpu().double() if torch.is_tensor(atol) else atol
    err = ((a - b).abs() - atol)
    assert float(err.max()) <= 0, "%s: worst excess %g (max abs diff %g)" % (what, float(err.max()), float((a - b).abs().max()))


def _dmm_mods(K, D, H=32, NSS=8, seed=0):
    from apgs_b200.dmm import models as M
    torch.manual_seed(seed)
    P = od.make_params(K, D, H, NSS, H, H)
    ms = M.init_apg_models(K, D, H, NSS, H, H, 0.5, True, DEV)
    for m, key in zip(ms, ("enc_rws_mu", "enc_apg_local", "enc_apg_mu", "dec")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})
    return P, ms


def test_full_shape_C3_dmm_slices_vs_oracle():
    """one local step + one mu step at S=100, B=256, N=200, K=4 through the fused encoders and the weight kernels."""
    from apgs_b200.dmm import kernels
    K, D, N, S, B = 4, 2, 200, 100, 256
    P, (enc_rws_mu, enc_local, enc_apg_mu, dec) = _dmm_mods(K, D, seed=3)
    prior = od.make_prior(K, D)
    torch.manual_seed(4)
    x = od.simulate(B, N, K, D)
    g = torch.Generator(device=DEV).manual_seed(5)
    xd = x.to(DEV)
    mu = torch.randn(S, B, K, D, device=DEV, generator=g) * 3
    z_old = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N), device=DEV, generator=g), K).float()
    beta_old = torch.rand(S, B, N, 1, device=DEV, generator=g) * 0.98 + 0.01
    u = torch.rand(S, B, N, device=DEV, generator=g)
    sel = [0, 1, 127, 254, 255]
    with torch.no_grad():
        q, z, c1, c0, c1b, c0b = kernels.local_encoder_fused(xd, mu, enc_local, uniforms=u, z_old=z_old)
        beta = (torch.rand(S, B, N, 1, device=DEV, generator=g) * 0.98 + 0.01)
        h2, h2_old = dec.hidden2(beta), dec.hidden2(beta_old)
        w = kernels.logw_local(xd, mu, q, z, beta, c1, c0, h2, z_old, beta_old, c1b, c0b, h2_old, dec, parts=True, local=True)
    # ---- oracle on the selected instances
    c = lambda t: t[:, sel].cpu()                                             # noqa: E731
    xs = x[sel].unsqueeze(0).repeat(S, 1, 1, 1)
    p = P["enc_apg_local"]
    q_ref = od.enc_local_probs(p, xs, c(mu), K)
    _close(c(q), q_ref, 2e-5 * q_ref + 1e-7, "q_probs (C3)")
    cdf = torch.cumsum(q_ref, -1)
    lab_ref = (c(u).unsqueeze(-1) * cdf[..., -1:] >= cdf[..., :-1]).sum(-1)
    bad = c(z).argmax(-1) != lab_ref
    assert float(bad.float().mean()) <= 2e-3
    gap = (c(u).unsqueeze(-1) * cdf[..., -1:] - cdf[..., :-1]).abs().min(-1).values
    assert bool((gap[bad] < 1e-5).all())
    c1_ref, c0_ref = od.enc_local_cons(p, xs, c(mu), c(z))
    c1b_ref, c0b_ref = od.enc_local_cons(p, xs, c(mu), c(z_old))
    for a, b, nm in ((c1, c1_ref, "con1"), (c0, c0_ref, "con0"), (c1b, c1b_ref, "con1_old"), (c0b, c0b_ref, "con0_old")):
        _close(c(a), b, 2e-5 * b.abs(), nm + " (C3)")
    h2_ref, h2o_ref = od.dec_hidden2(P["dec"], c(beta)), od.dec_hidden2(P["dec"], c(beta_old))
    _close(c(h2), h2_ref, 1e-5, "hidden2 (C3): a unit vector, 1e-5 of its norm")
    # weights on the KERNEL's own proposal parameters (teacher-forced)
    ref = od.local_step_weights(prior, P["dec"], xs, c(mu), c(q), c(z), c(beta), c(c1), c(c0), c(h2), c(z_old), c(beta_old),
                                c(c1b), c(c0b), c(h2_old), parts=True)
    mag = ref[5].abs() + ref[6].abs() + (ref[2].abs() + ref[4].abs()).sum(-1)
    _close(c(w.log_w_f), ref[5], RT * mag, "local sum log_w_f (C3)")
    _close(c(w.density), ref[4].sum(-1), RT * ref[4].abs().sum(-1), "local density (C3)")
    _close(c(w.log_w), ref[1], RT * 2 * mag, "local log_w (C3)")
    _close(c(w.log_w_local), ref[0], RT * (ref[2].abs() + ref[4].abs()) * 2 + 1e-4, "log_w_local (C3)")
    # ---- mu step at the full shape
    with torch.no_grad():
        q_mu, q_sigma = enc_apg_mu.params(xd, z, beta, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
        mu_new = q_mu + q_sigma * torch.randn(S, B, K, D, device=DEV, generator=g)
        wm = kernels.logw_mu(xd, z, h2, q_mu, q_sigma, mu_new, mu, dec, parts=True)
    inp = torch.cat((xs, c(z)), -1)
    q_mu_ref, q_sigma_ref = od.enc_mu_params(P["enc_apg_mu"], inp, prior)
    _close(c(q_mu), q_mu_ref, 2e-5 * q_mu_ref.abs() + 2e-6, "q_mu (C3)")
    _close(c(q_sigma), q_sigma_ref, 2e-5 * q_sigma_ref.abs() + 1e-7, "q_sigma (C3)")
    ref = od.mu_step_weights(prior, P["dec"], xs, c(z), c(h2), c(q_mu), c(q_sigma), c(mu_new), c(mu), parts=True)
    mag_f = ref[3].abs() + ref[2].abs() + ref[1].abs()
    _close(c(wm.ll_f), ref[3], RT * ref[3].abs(), "mu ll_f (C3)")
    _close(c(wm.log_w_f), ref[4], RT * mag_f, "mu log_w_f (C3)")
    _close(c(wm.log_w), ref[0], RT * (2 * mag_f + ref[5].abs() + ref[4].abs()), "mu log_w (C3)")
    # ---- resample of the DMM latents at the full shape
    from apgs_b200 import ops
    from oracle import resampler as orc
    ub = torch.rand(B, device=DEV, generator=g)
    anc, wn, ess, lse = ops.sample_ancestors(log_w=wm.log_w, strategy="systematic", uniforms=ub, want_stats=True)
    a_ref = orc.systematic_ancestors(c(wn).numpy(), ub[sel].cpu().numpy())
    assert np.array_equal(c(anc).numpy(), a_ref)
    outs = ops.gather_ancestors([mu_new, z, beta, h2], anc)
    for o, v in zip(outs, (mu_new, z, beta, h2)):
        assert np.array_equal(c(o).numpy(), orc.gather_ancestors(c(v).numpy(), a_ref))


def _bm_models(seed=0):
    from apgs_b200.bmnist import models as M
    torch.manual_seed(seed)
    P = ob.make_params()
    ms = M.init_models(96, 28, 400, 400, 2, 10, True, DEV)
    for m, key in ((ms[0], "enc_coor"), (ms[2], "enc_digit"), (ms[3], "dec_digit")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})
    return P, ms


def test_full_shape_C4_bmnist_slices_vs_oracle():
    """crop, paste + Bernoulli likelihood, template matching and the `what` weights at S=10, B=64, T=10, K=3, 96x96."""
    from apgs_b200.bmnist import Affine_Transformer, kernels
    from apgs_b200.bmnist import sim
    K, T, S, B, FP, DP, ZD = 3, 10, 10, 64, 96, 28, 10
    P, (enc_coor, dec_coor, enc_digit, dec_digit) = _bm_models()
    AT = Affine_Transformer(FP, DP, True, DEV)
    torch.manual_seed(7)
    frames = sim.simulate(B, T, K)                                          # (B,T,96,96) CPU
    z_where = (torch.rand(S, B, T, K, 2) * 2 - 1) * 1.1                     # some digits partly outside the frame
    z_what, z_what_old = torch.randn(S, B, K, ZD), torch.randn(S, B, K, ZD)
    fd, zw = frames.to(DEV), z_where.to(DEV)
    sel = [0, 31, 63]
    c = lambda t: t[:, sel].cpu()                                             # noqa: E731
    frs = frames[sel].unsqueeze(0).repeat(S, 1, 1, 1, 1)
    with torch.no_grad():
        cropped = AT.frame_to_digit(fd, zw)
        assert torch.equal(c(cropped), ob.frame_to_digit(frs, z_where[:, sel])), "crop (C4): bit-exact"
        q_mu, q_std = enc_digit.params(cropped.view(S, B, T, K, DP * DP))
        q_mu_ref, q_std_ref = ob.enc_digit_params(P["enc_digit"], c(cropped).view(S, len(sel), T, K, DP * DP))
        _close(c(q_mu), q_mu_ref, 2e-4 * q_mu_ref.abs() + 2e-5, "Enc_digit mean (C4; cuBLAS vs CPU GEMM)")
        dig_f, dig_b = dec_digit.templates(z_what.to(DEV)), dec_digit.templates(z_what_old.to(DEV))
        ll_f, recon = AT.reconstruct_ll(dig_f, zw, fd, want_recon=True)
        ll_b, _ = AT.reconstruct_ll(dig_b, zw, fd)
        # paste + sum_k + clamp + Bernoulli ll on the KERNEL's digit templates (teacher-forced)
        recon_ref, ll_ref = ob.recon_ll(c(dig_f), z_where[:, sel], frs)
        _close(c(ll_f), ll_ref, RT * ll_ref.abs() + 1e-3, "ll_f (C4)")
        pasted = ob.digit_to_frame(c(dig_f), z_where[:, sel])
        assert torch.equal(c(AT.digit_to_frame(dig_f, zw)), pasted), "paste (C4): bit-exact"
        assert float((c(recon) - recon_ref).abs().max()) < 1e-6, "recon (C4)"
        log_w, log_q_f, dens, ll_sum = kernels.logw_what(q_mu, q_std, z_what.to(DEV), z_what_old.to(DEV), ll_f, ll_b)
        lw_ref = ob.what_step_weight(c(q_mu), c(q_std), z_what[:, sel], z_what_old[:, sel], c(ll_f), c(ll_b))
        ref_w = lw_ref[0] if isinstance(lw_ref, (tuple, list)) else lw_ref
        mag = c(ll_f).abs().sum(-1) + c(ll_b).abs().sum(-1) + 100.0
        _close(c(log_w), ref_w, RT * mag, "what log_w (C4)")
        # template matching of frame 0 against digit templates (N3)
        tmpl = dig_f[:, :, 0]
        conv = kernels.template_match(fd[:, 0].unsqueeze(0).expand(S, -1, -1, -1), tmpl, softmax=False)
        ref_conv = torch.nn.functional.conv2d(frs[:, :, 0].reshape(1, S * len(sel), FP, FP), c(tmpl).reshape(S * len(sel), 1, DP, DP),
                                              groups=S * len(sel)).view(S, len(sel), -1)
        _close(c(conv).reshape(S, len(sel), -1), ref_conv, RT * 800.0, "template match scores (C4)")


def test_dmm_and_bmnist_bpg_objectives_run_on_the_kernels():
    """N4: bpg_objective of both tasks (apgs/dmm/objectives.py:209-236, apgs/bmnist/objectives.py:272-302): shapes,
    finiteness, and the prior-proposal weight really is the likelihood ratio (log p - log q cancels in the kernel)."""
    from apgs_b200.dmm import objectives as DO, kernels as dk
    from apgs_b200.bmnist import objectives as BO, Affine_Transformer, sim, kernels as bk
    from apgs_b200.resampler import Resampler
    flags = {"loss_required": False, "ess_required": False, "mode_required": False, "density_required": True}
    K, D, N, S, B = 4, 2, 200, 30, 6
    P, models = _dmm_mods(K, D, seed=1)
    dec = models[3]
    torch.manual_seed(2)
    x = od.simulate(B, N, K, D).to(DEV)
    with torch.no_grad():
        tr = DO.bpg_objective(models, x, K, flags, 3, Resampler("systematic", S, True, DEV))
    assert tuple(tr["density"].shape) == (3, S, B) and bool(torch.isfinite(tr["density"]).all())
    with torch.no_grad():
        z = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N), device=DEV), K).float()
        beta = torch.rand(S, B, N, 1, device=DEV) * 0.98 + 0.01
        mu_old = torch.randn(S, B, K, D, device=DEV) * 3
        t2 = {"density": []}
        torch.manual_seed(3)
        log_w, mu, t2 = DO.bpg_update_mu(models[2], dec, x, z, beta, mu_old, K, t2)
        h2 = dec.hidden2(beta)
    prior = od.make_prior(K, D)
    xs = x.cpu().unsqueeze(0).repeat(S, 1, 1, 1)
    ll_f = od.likelihood(prior, P["dec"], xs, mu.cpu(), z.cpu(), h2.cpu()).sum(-1).sum(-1)
    ll_b = od.likelihood(prior, P["dec"], xs, mu_old.cpu(), z.cpu(), h2.cpu()).sum(-1).sum(-1)
    _close(log_w, ll_f - ll_b, RT * (ll_f.abs() + ll_b.abs()) + 1e-3, "DMM bpg log_w = ll_f - ll_b")
    _close(t2["density"][0][0], od.mu_log_prior(prior, mu.cpu()).sum(-1).sum(-1), 1e-4, "DMM bpg density = log p(mu)")
    # ---- bmnist
    K, T, S, B = 3, 4, 6, 3
    Pb, bm = _bm_models()
    AT = Affine_Transformer(96, 28, True, DEV)
    torch.manual_seed(4)
    frames = sim.simulate(B, T, K, device=DEV)
    mm = sim.blobs(1)[0].to(DEV).repeat(S, B, K, 1, 1)
    with torch.no_grad():
        tr = BO.bpg_objective(bm, AT, frames, flags, 2, Resampler("systematic", S, True, DEV), mm)
    assert tuple(tr["density"].shape) == (2, S, B) and bool(torch.isfinite(tr["density"]).all())
    with torch.no_grad():
        z_where = (torch.rand(S, B, T, K, 2, device=DEV) * 2 - 1)
        z_old = torch.randn(S, B, K, 10, device=DEV)
        t3 = {"density": [torch.zeros(1, S, B, device=DEV)]}
        fr5 = frames.unsqueeze(0).expand(S, -1, -1, -1, -1)
        log_w, z_new, t3 = BO.bpg_what(bm[3], AT, fr5, z_where, z_old, t3)
        ll_f, _ = AT.reconstruct_ll(bm[3].templates(z_new), z_where, fr5)
        ll_b, _ = AT.reconstruct_ll(bm[3].templates(z_old), z_where, fr5)
    _close(log_w, (ll_f.sum(-1) - ll_b.sum(-1)), RT * (ll_f.abs().sum(-1) + ll_b.abs().sum(-1)) + 1e-3, "bmnist bpg log_w")
    lp = torch.distributions.Normal(0., 1.).log_prob(z_new.cpu()).sum(-1).sum(-1)
    _close(t3["density"][0][0], ll_f.sum(-1).cpu() + lp, RT * (ll_f.abs().sum(-1).cpu() + lp.abs()) + 1e-3, "bmnist bpg density")


def test_bmnist_simulator_renders_with_the_paste_kernel():
    """sim.simulate(device=cuda) == the CPU formula (frames rendered by k_stn_paste)."""
    from apgs_b200.bmnist import sim
    torch.manual_seed(11)
    a = sim.simulate(5, device=DEV)
    torch.manual_seed(11)
    b = sim.simulate(5)
    assert a.is_cuda and torch.equal(a.cpu(), b)
