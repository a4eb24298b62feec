"""GPU parity: DMM kernels (D1 NSS pooling, D2 mu step, D3 local step, one-shot) and the sweep, through the C ABI,
against the reference-generated golden fixtures and the oracle.  Tolerance: 1e-5 relative to the magnitudes that are
summed, as in tests/test_gmm_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import dists
from oracle import dmm as od

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


def _models(P, K, D):
    from apgs_b200.dmm import models as M
    ms = M.init_apg_models(K, D, 32, 8, 32, 32, 0.5, True, DEV)
    for m, key in zip(ms, ("enc_rws_mu", "enc_apg_local", "enc_apg_mu", "dec")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})       # the reference's state_dict keys
    return ms


def _close(a, b, atol, what):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    atol = atol.cpu().double() if torch.is_tensor(atol) else atol
    err = ((a - b).abs() - atol)
    assert float(err.max()) <= 0, "%s: worst excess %g (max abs diff %g)" % (what, float(err.max()), float((a - b).abs().max()))


def _d(*ts):
    return [t.to(DEV) for t in ts]


@pytest.mark.parametrize("S,B,N,K,J,D", [(6, 5, 60, 4, 8, 2), (3, 2, 201, 3, 5, 2), (1, 1, 7, 16, 3, 8), (33, 4, 200, 4, 8, 2)])
def test_D1_nss_pool_vs_oracle(S, B, N, K, J, D):
    from apgs_b200.dmm import kernels
    torch.manual_seed(S + N)
    nss1 = torch.randn(S, B, N, J)
    nss2 = torch.softmax(torch.randn(S, B, N, K) * 2, -1)
    prior = {"mu_mu": torch.randn(D), "mu_sigma": torch.rand(D) + 1}
    ref = od.nss_pool(nss1, nss2, prior)
    out = kernels.nss_pool(nss1.to(DEV), nss2.to(DEV), prior["mu_mu"].to(DEV), prior["mu_sigma"].to(DEV))
    assert tuple(out.shape) == tuple(ref.shape)
    mag = (nss1.abs().unsqueeze(-2) * nss2.unsqueeze(-1)).sum(2) / nss2.sum(2).unsqueeze(-1)      # sum of |terms| / denominator
    _close(out[..., :J], ref[..., :J], RT * mag + 1e-7, "pooled nss")
    assert torch.equal(out[..., J:].cpu(), ref[..., J:])


def test_D2_D3_oneshot_vs_reference_golden(golden):
    from apgs_b200.dmm import kernels
    g = golden("dmm")
    S, B, N, K, D, sweeps = [int(v) for v in g["dims"]]
    P, prior = _params(g), od.make_prior(K, D)
    dec = _models(P, K, D)[3]
    x = _t(g, "x")
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    # ---- D2
    ref = od.mu_step_weights(prior, P["dec"], xs, _t(g, "mu_z"), _t(g, "mu_hidden2"), _t(g, "mu_q_mu"), _t(g, "mu_q_sigma"),
                             _t(g, "mu_mu"), _t(g, "mu_mu_old"), parts=True)
    assert torch.equal(ref[0], _t(g, "mu_ref_log_w"))
    w = kernels.logw_mu(x.to(DEV), *_d(_t(g, "mu_z"), _t(g, "mu_hidden2"), _t(g, "mu_q_mu"), _t(g, "mu_q_sigma"), _t(g, "mu_mu"),
                                       _t(g, "mu_mu_old")), dec, parts=True)
    mag_f = ref[3].abs() + ref[2].abs() + ref[1].abs()
    _close(w.ll_f, ref[3], RT * ref[3].abs(), "mu ll_f")
    _close(w.log_q_f, ref[1], RT * ref[1].abs() + 1e-5, "mu log_q_f")
    _close(w.density, ref[2], RT * ref[2].abs(), "mu log_prior_f")
    _close(w.log_w_f, ref[4], RT * mag_f, "mu log_w_f")
    _close(w.log_w, _t(g, "mu_ref_log_w"), RT * (2 * mag_f + ref[5].abs()), "mu log_w")
    # ---- D3
    args = [_t(g, "lo_" + n) for n in ("mu", "q_probs", "z", "beta", "con1", "con0", "hidden2", "z_old", "beta_old", "con1_old",
                                        "con0_old", "hidden2_old")]
    ref = od.local_step_weights(prior, P["dec"], xs, *args, parts=True)
    assert torch.equal(ref[1], _t(g, "lo_ref_log_w"))
    w = kernels.logw_local(x.to(DEV), *_d(*args), dec, parts=True, local=True)
    # per point: three O(1..10) terms forward and backward
    _close(w.log_q_local, ref[2], RT * ref[2].abs() + 2e-5, "local log_q_f")
    _close(w.ll_local, ref[3], RT * ref[3].abs() + 1e-5, "local ll_f")
    _close(w.log_w_local, ref[0], RT * (ref[2].abs() + ref[4].abs()) * 2 + 1e-4, "log_w_local")
    mag = ref[5].abs() + ref[6].abs() + (ref[2].abs() + ref[4].abs()).sum(-1)
    _close(w.log_w_f, ref[5], RT * mag, "local sum log_w_f")
    _close(w.density, ref[4].sum(-1), RT * ref[4].abs().sum(-1), "local density")
    _close(w.log_w, _t(g, "lo_ref_log_w"), RT * 2 * mag, "local log_w")
    # ---- one-shot
    args = [_t(g, "os_" + n) for n in ("q_mu", "q_sigma", "mu", "q_probs", "z", "beta", "con1", "con0", "hidden2")]
    lw, lq, lp, ll = od.oneshot_weights(prior, P["dec"], xs, *args)
    assert torch.equal(lw, _t(g, "os_ref_log_w"))
    w = kernels.logw_oneshot(x.to(DEV), *_d(*args), dec)
    _close(w.density, lp, RT * (lp.abs() + ll.abs()), "oneshot log_p")
    _close(w.log_q_f, lq, RT * lq.abs() + 1e-4, "oneshot log_q")
    _close(w.ll_f, ll, RT * ll.abs(), "oneshot ll")
    _close(w.log_w, lw, RT * (lp.abs() + lq.abs() + ll.abs()), "oneshot log_w")


def test_dmm_encoder_parameters_match_oracle(golden):
    g = golden("dmm")
    S, B, N, K, D, sweeps = [int(v) for v in g["dims"]]
    P, prior = _params(g), od.make_prior(K, D)
    enc_rws_mu, enc_apg_local, enc_apg_mu, dec = _models(P, K, D)
    x = _t(g, "x")
    with torch.no_grad():
        q_mu, q_sigma = enc_apg_mu.params(x.to(DEV), _t(g, "mu_z").to(DEV), None, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
        np.testing.assert_allclose(q_mu.cpu().numpy(), g["mu_q_mu"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(q_sigma.cpu().numpy(), g["mu_q_sigma"], rtol=2e-5, atol=1e-7)
        q_mu, q_sigma = enc_rws_mu.params(x.to(DEV), K, (dec.prior_mu_mu, dec.prior_mu_sigma), S=S)
        np.testing.assert_allclose(q_mu.cpu().numpy(), g["os_q_mu"], rtol=2e-5, atol=2e-6)
        qp = enc_apg_local.probs(x.to(DEV), _t(g, "lo_mu").to(DEV), K)
        np.testing.assert_allclose(qp.cpu().numpy(), g["lo_q_probs"], rtol=2e-5, atol=1e-7)
        c1, c0 = enc_apg_local.cons(x.to(DEV), _t(g, "lo_mu").to(DEV), _t(g, "lo_z").to(DEV))
        np.testing.assert_allclose(c1.cpu().numpy(), g["lo_con1"], rtol=2e-5)
        np.testing.assert_allclose(dec.hidden2(_t(g, "lo_beta").to(DEV)).cpu().numpy(), g["lo_hidden2"], rtol=2e-5, atol=1e-6)
        # reference forward() contract
        p = dec(x.to(DEV).unsqueeze(0).expand(S, -1, -1, -1), mu=_t(g, "lo_mu").to(DEV), z=_t(g, "lo_z").to(DEV), beta=_t(g, "lo_beta").to(DEV))
        assert set(p.keys()) == {"means", "states", "angles", "likelihood"} and p["likelihood"].log_prob.shape == (S, B, N, D)


# ------------------------------------------------------------------------ N1: fused encoders / decoder
def _mods(K, D, H, NSS=8, seed=0):
    from apgs_b200.dmm import models as M
    torch.manual_seed(seed)
    P = od.make_params(K, D, H, NSS, H, H)
    ms = M.init_apg_models(K, D, H, NSS, H, H, 0.5, True, DEV)
    for m, key in zip(ms, ("enc_rws_mu", "enc_apg_local", "enc_apg_mu", "dec")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})
    return P, ms


@pytest.mark.parametrize("S,B,N,K,D,H", [(5, 3, 200, 4, 2, 32), (19, 2, 37, 3, 2, 16), (9, 1, 65, 4, 3, 33), (2, 2, 5, 5, 2, 8)])
def test_N1_enc_apg_mu_label_tables_vs_oracle(S, B, N, K, D, H):
    """Enc_apg_mu through per-instance label tables + the one-hot pooling kernel == the oracle's per-particle MLPs."""
    P, (_, _, enc_apg_mu, dec) = _mods(K, D, H, seed=S + N)
    prior = od.make_prior(K, D)
    x = torch.randn(B, N, D) * 2
    z = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    if S > 2:
        z[1, 0] = torch.nn.functional.one_hot(torch.zeros(N, dtype=torch.long), K).float()      # empty clusters
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    inp = torch.cat((xs, z), -1)
    p = P["enc_apg_mu"]
    nss_ref = od.nss_pool(od._mlp2(p, "nss1", inp, torch.tanh), torch.softmax(od._mlp2(p, "nss2", inp, torch.tanh), -1), prior)
    q_mu_ref, q_sigma_ref = od.enc_mu_params(p, inp, prior)
    from apgs_b200.dmm import kernels
    with torch.no_grad():
        tab1, tab2 = enc_apg_mu.label_tables(x.to(DEV))
        nss = kernels.nss_pool_onehot(tab1, tab2, z.to(DEV), dec.prior_mu_mu, dec.prior_mu_sigma)
        q_mu, q_sigma = enc_apg_mu.params(x.to(DEV), z.to(DEV), None, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
    _close(nss, nss_ref, 2e-5 * nss_ref.abs() + 2e-6, "pooled nss")
    _close(q_mu, q_mu_ref, 2e-5 * q_mu_ref.abs() + 2e-6, "q_mu")
    _close(q_sigma, q_sigma_ref, 2e-5 * q_sigma_ref.abs() + 1e-7, "q_sigma")
    enc_apg_mu.fused = False                                   # the PyTorch per-particle path agrees too
    try:
        with torch.no_grad():
            q_mu2, _ = enc_apg_mu.params(x.to(DEV), z.to(DEV), None, K, (dec.prior_mu_mu, dec.prior_mu_sigma))
    finally:
        type(enc_apg_mu).fused = True
        del enc_apg_mu.fused
    _close(q_mu2, q_mu_ref, 2e-5 * q_mu_ref.abs() + 2e-6, "q_mu (unfused)")


@pytest.mark.parametrize("S,B,N,K,D,H", [(5, 3, 200, 4, 2, 32), (19, 2, 37, 3, 2, 16), (9, 1, 65, 4, 3, 33), (2, 2, 5, 5, 2, 8)])
def test_N1_enc_apg_local_fused_vs_oracle(S, B, N, K, D, H):
    from apgs_b200.dmm import kernels
    P, (_, enc_local, _, _) = _mods(K, D, H, seed=S + N + 1)
    p = P["enc_apg_local"]
    x = torch.randn(B, N, D) * 2
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    mu = torch.randn(S, B, K, D) * 3
    one_hot = lambda: torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    z_f, z_old = one_hot(), one_hot()
    q_ref = od.enc_local_probs(p, xs, mu, K)
    c1_ref, c0_ref = od.enc_local_cons(p, xs, mu, z_f)
    c1b_ref, c0b_ref = od.enc_local_cons(p, xs, mu, z_old)
    # teacher-forced: the given assignment, concentrations for it and for z_old
    q, z, c1, c0, c1b, c0b = kernels.local_encoder_fused(x.to(DEV), mu.to(DEV), enc_local, z_forced=z_f.to(DEV), z_old=z_old.to(DEV))
    assert torch.equal(z.cpu(), z_f)
    _close(q, q_ref, 2e-5 * q_ref + 1e-7, "q_probs")
    for a, b, w in ((c1, c1_ref, "con1"), (c0, c0_ref, "con0"), (c1b, c1b_ref, "con1_old"), (c0b, c0b_ref, "con0_old")):
        assert tuple(a.shape) == (S, B, N, 1)
        _close(a, b, 2e-5 * b.abs(), w)
    # sampling: inverse CDF of the oracle's probabilities at the supplied uniforms
    u = torch.rand(S, B, N)
    q2, z2, c1s, c0s, n1, n0 = kernels.local_encoder_fused(x.to(DEV), mu.to(DEV), enc_local, uniforms=u.to(DEV))
    assert n1 is None and n0 is None
    z2 = z2.cpu()
    assert bool(((z2 == 0) | (z2 == 1)).all()) and bool((z2.sum(-1) == 1).all())
    cdf = torch.cumsum(q_ref, -1)
    lab_ref = (u.unsqueeze(-1) * cdf[..., -1:] >= cdf[..., :-1]).sum(-1)
    bad = z2.argmax(-1) != lab_ref
    assert float(bad.float().mean()) <= 2e-3
    gap = (u.unsqueeze(-1) * cdf[..., -1:] - cdf[..., :-1]).abs().min(-1).values
    assert bool((gap[bad] < 1e-5).all())
    c1s_ref, c0s_ref = od.enc_local_cons(p, xs, mu, z2)            # concentrations belong to the DRAWN assignment
    _close(c1s, c1s_ref, 2e-5 * c1s_ref.abs(), "con1 at the draw")
    _close(c0s, c0s_ref, 2e-5 * c0s_ref.abs(), "con0 at the draw")


def test_N1_local_encoder_saturating_inputs():
    """|x - mu| far beyond the tanh saturation point: finite outputs that still match the oracle."""
    from apgs_b200.dmm import kernels
    S, B, N, K, D, H = 4, 2, 50, 4, 2, 32
    P, (_, enc_local, _, _) = _mods(K, D, H, seed=7)
    p = P["enc_apg_local"]
    x = torch.randn(B, N, D) * 300
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    mu = torch.randn(S, B, K, D) * 300
    z_f = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    q, z, c1, c0, _, _ = kernels.local_encoder_fused(x.to(DEV), mu.to(DEV), enc_local, z_forced=z_f.to(DEV))
    c1_ref, c0_ref = od.enc_local_cons(p, xs, mu, z_f)
    assert bool(torch.isfinite(q).all()) and bool(torch.isfinite(c1).all())
    _close(c1, c1_ref, 3e-5 * c1_ref.abs(), "con1")
    _close(c0, c0_ref, 3e-5 * c0_ref.abs(), "con0")
    q_ref = od.enc_local_probs(p, xs, mu, K)
    _close(q, q_ref, 1e-4 * q_ref + 1e-6, "q_probs")          # logits of O(100): fp32 rounding of the split first layer


@pytest.mark.parametrize("shape,D,H", [((7, 3, 200, 1), 2, 32), ((1000,), 2, 16), ((3, 5, 1), 3, 40)])
def test_N1_decoder_hidden2_vs_oracle(shape, D, H):
    from apgs_b200.dmm import kernels
    P, (_, _, _, dec) = _mods(4, D, H, seed=H)
    beta = torch.rand(*shape)
    ref = od.dec_hidden2(P["dec"], beta if shape[-1] == 1 else beta.unsqueeze(-1))
    out = kernels.decoder_hidden2(beta.to(DEV), dec.recon_mu)
    assert tuple(out.shape) == tuple(ref.shape)
    _close(out, ref, 2e-5 * ref.abs() + 2e-6, "hidden2")
    assert float(((out ** 2).sum(-1) - 1).abs().max()) < 1e-5


def test_N1_dmm_golden_through_fused_encoders(golden):
    g = golden("dmm")
    S, B, N, K, D, sweeps = [int(v) for v in g["dims"]]
    P = _params(g)
    _, enc_apg_local, _, _ = _models(P, K, D)
    x = _t(g, "x")
    q, z, c1, c0, c1b, c0b = enc_apg_local.sample_fused(x.to(DEV), _t(g, "lo_mu").to(DEV), z_forced=_t(g, "lo_z").to(DEV),
                                                        z_old=_t(g, "lo_z_old").to(DEV))
    np.testing.assert_allclose(q.cpu().numpy(), g["lo_q_probs"], rtol=2e-5, atol=1e-7)
    for a, k in ((c1, "lo_con1"), (c0, "lo_con0"), (c1b, "lo_con1_old"), (c0b, "lo_con0_old")):
        np.testing.assert_allclose(a.cpu().numpy(), g[k], rtol=2e-5)


def test_dmm_sweep_statistics_match_oracle():
    """BASELINE.json configs[2] structure (K=4 rings, N=200) at reduced S,B: end to end, statistical (RNG differs)."""
    from apgs_b200.dmm import objectives as O
    from apgs_b200.resampler import Resampler
    K, D, N = 4, 2, 200
    S, B, sweeps = 30, 120, 4
    torch.manual_seed(0)
    P = od.make_params(K, D)
    torch.manual_seed(1)
    x = od.simulate(B, N, K, D)
    flags = {"loss_required": True, "ess_required": True, "mode_required": True, "density_required": True}
    torch.manual_seed(2)
    with torch.no_grad():
        ref = od.apg_objective(P, od.make_prior(K, D), x.repeat(S, 1, 1, 1), K, flags, sweeps)
        models = _models(P, K, D)
        torch.manual_seed(3)
        tr = O.apg_objective(models, x.to(DEV), K, flags, sweeps, Resampler("systematic", S, True, DEV))
    for k in ("loss_phi", "loss_theta", "ess", "density"):
        assert tuple(tr[k].shape) == tuple(ref[k].shape), k
        assert bool(torch.isfinite(tr[k]).all()), k
    assert tuple(tr["E_recon"].shape) == (sweeps, B, N, D)
    for k in ("ess", "density"):
        a = tr[k][-1].cpu().reshape(-1, B).mean(0)
        b = ref[k][-1].reshape(-1, B).mean(0)
        se = (a.var() / B + b.var() / B).sqrt()
        assert abs(float(a.mean() - b.mean())) < 5 * float(se), (k, float(a.mean()), float(b.mean()), float(se))


def test_dmm_full_size_runs():
    """BASELINE.json configs[2]: K=4, N=200, S=100, B=256."""
    from apgs_b200.dmm import objectives as O, models as M
    from apgs_b200.resampler import Resampler
    K, D, N, S, B = 4, 2, 200, 100, 256
    torch.manual_seed(0)
    models = M.init_apg_models(K, D, 32, 8, 32, 32, 0.5, True, DEV)
    x = od.simulate(B, N, K, D).to(DEV)
    flags = {"loss_required": False, "ess_required": True, "mode_required": False, "density_required": True}
    with torch.no_grad():
        tr = O.apg_objective(models, x, K, flags, 3, Resampler("systematic", S, True, DEV))
    assert tuple(tr["ess"].shape) == (3, B) and bool(torch.isfinite(tr["density"]).all())


def test_beta_sample_matches_torch_distribution():
    """beta_sample draws Beta(con1, con0) like torch.distributions (moments + support), from adjacent or separate buffers."""
    from apgs_b200.dmm import kernels
    torch.manual_seed(0)
    buf = torch.empty(2, 400000, 1, device=DEV)
    buf[0], buf[1] = 2.5, 0.7
    for c1, c0 in ((buf[0], buf[1]), (buf[0].clone(), buf[1].clone())):
        b = kernels.beta_sample(c1, c0)
        assert tuple(b.shape) == tuple(c1.shape) and float(b.min()) > 0 and float(b.max()) < 1
        m, v = 2.5 / 3.2, 2.5 * 0.7 / (3.2 ** 2 * 4.2)
        assert abs(float(b.mean()) - m) < 5 * (v / b.numel()) ** 0.5
        assert abs(float(b.var()) - v) < 0.02 * v


# ------------------------------------------------------------------------------ N2: training mode
def _train_mods(K=4, D=2, H=32):
    from apgs_b200.dmm import models as M
    torch.manual_seed(0)
    P = od.make_params(K, D, H, 8, H, H)
    ms, opt = M.init_apg_models(K, D, H, 8, H, H, 0.5, True, DEV, lr=1e-3)
    for m, key in zip(ms, ("enc_rws_mu", "enc_apg_local", "enc_apg_mu", "dec")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})
    return P, ms, opt


def _grad_close(module, Pg, rtol=3e-3):
    scale = max(float(v.grad.abs().max()) for v in Pg.values() if v.grad is not None)
    for name, p in module.named_parameters():
        ref = Pg[name].grad
        assert p.grad is not None, name
        np.testing.assert_allclose(p.grad.cpu().numpy(), ref.numpy(), rtol=rtol, atol=1e-4 * scale, err_msg=name)


def test_N2_dmm_step_gradients_match_oracle_autograd():
    """loss_phi / loss_theta .backward() through the drop-in mu and local steps == autograd through the oracle
    (= the reference's computation) on the same draws."""
    from apgs_b200.dmm import objectives as O
    K, D, N, S, B = 4, 2, 50, 7, 5
    P, (enc_rws_mu, enc_apg_local, enc_apg_mu, dec), _ = _train_mods(K, D)
    prior = od.make_prior(K, D)
    torch.manual_seed(1)
    x = od.simulate(B, N, K, D)
    xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
    z = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
    beta, mu_old = torch.rand(S, B, N, 1), torch.randn(S, B, K, D) * 2
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    new = lambda: {"loss_phi": [], "loss_theta": [], "ess": [], "E_mu": [], "E_z": [], "E_recon": [], "density": []}
    leaf = lambda d: {k: v.clone().requires_grad_(True) for k, v in d.items()}
    # ---- mu step
    trace = new()
    log_w, mu, trace = O.apg_update_mu(enc_apg_mu, dec, x.to(DEV), z.to(DEV), beta.to(DEV), mu_old.to(DEV), K, trace, flags)
    (trace["loss_phi"][0] + trace["loss_theta"][0]).sum().backward()
    Pm, Pd = leaf(P["enc_apg_mu"]), leaf(P["dec"])
    q_mu, q_sigma = od.enc_mu_params(Pm, torch.cat((xs, z), -1), prior)
    lw, lq, _, ll = od.mu_step_weights(prior, Pd, xs, z, od.dec_hidden2(Pd, beta), q_mu, q_sigma, mu.cpu(), mu_old)
    w = torch.softmax(lw.detach(), 0)
    ((w * (-lq)).sum(0).mean() + (w * (-ll)).sum(0).mean()).backward()
    np.testing.assert_allclose(float(trace["loss_phi"][0].detach()), float(((w * (-lq)).sum(0).mean()).detach()), rtol=3e-4)
    np.testing.assert_allclose(float(trace["loss_theta"][0].detach()), float(((w * (-ll)).sum(0).mean()).detach()), rtol=3e-4)
    _grad_close(enc_apg_mu, Pm)
    _grad_close(dec, Pd)
    # ---- local step
    for m in (enc_apg_mu, dec):
        m.zero_grad()
    trace = new()
    trace["loss_phi"].append(torch.zeros(1, device=DEV)); trace["loss_theta"].append(torch.zeros(1, device=DEV))
    trace["density"].append(torch.zeros(1, S, B, device=DEV))
    log_w, z_new, beta_new, trace = O.apg_update_local(enc_apg_local, dec, x.to(DEV), mu, z.to(DEV), beta.to(DEV), K, trace, flags)
    (trace["loss_phi"][0] + trace["loss_theta"][0]).sum().backward()
    Pl, Pd = leaf(P["enc_apg_local"]), leaf(P["dec"])
    muc, zc, bc = mu.cpu(), z_new.cpu(), beta_new.cpu()
    q_probs = od.enc_local_probs(Pl, xs, muc, K)
    c1, c0 = od.enc_local_cons(Pl, xs, muc, zc)
    c1b, c0b = od.enc_local_cons(Pl, xs, muc, z)
    lwl, _, lqf, llf, _ = od.local_step_weights(prior, Pd, xs, muc, q_probs, zc, bc, c1, c0, od.dec_hidden2(Pd, bc),
                                                z, beta, c1b, c0b, od.dec_hidden2(Pd, beta))
    wl = torch.softmax(lwl.detach(), 0)
    ref_phi, ref_theta = (wl * (-lqf)).sum(0).sum(-1).mean(), (wl * (-llf)).sum(0).sum(-1).mean()
    (ref_phi + ref_theta).backward()
    np.testing.assert_allclose(float(trace["loss_phi"][0].detach()), float((ref_phi).detach()), rtol=3e-4)
    np.testing.assert_allclose(float(trace["loss_theta"][0].detach()), float((ref_theta).detach()), rtol=3e-4)
    _grad_close(enc_apg_local, Pl)
    _grad_close(dec, Pd)


def test_N2_dmm_training_loop_runs_like_reference_train():
    """the reference's train() inner loop (apgs/dmm/apg_training.py) on the drop-in: finite losses, every parameter gets
    a gradient, Adam steps change the weights."""
    from apgs_b200.dmm import objectives as O
    from apgs_b200.resampler import Resampler
    K, D, N, S, B, sweeps = 4, 2, 200, 10, 20, 4
    P, models, opt = _train_mods(K, D)
    torch.manual_seed(1)
    data = od.simulate(B, N, K, D).to(DEV)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    rs = Resampler("systematic", S, True, DEV)
    before = [p.detach().clone() for m in models for p in m.parameters()]
    for it in range(3):
        opt.zero_grad()
        trace = O.apg_objective(models, data.repeat(S, 1, 1, 1), K, flags, sweeps, rs)
        loss = trace["loss_phi"].sum() + trace["loss_theta"].sum()
        loss.backward()
        assert bool(torch.isfinite(loss))
        for m in models:
            for n_, p in m.named_parameters():
                assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n_
        opt.step()
    after = [p.detach() for m in models for p in m.parameters()]
    assert all(not torch.equal(a, b) for a, b in zip(before, after))


def test_local_encoder_draws_in_the_kernel_beta_distribution_and_decoder():
    """apgs_dmm_local_encoder_draw_f32: (i) with explicit uniforms the encoder outputs are those of the plain entry bit for
    bit; (ii) beta ~ Beta(con1, con0): with (mu, z) replicated over S the draws of a point are iid over s -- mean / variance
    against the closed forms within 4.5 standard errors, a Kolmogorov-Smirnov distance against torch's own Beta draws, no
    correlation between neighbouring points or particles, different seeds differ; (iii) hidden2 == the decoder kernel on the
    drawn beta, bit for bit; (iv) the in-kernel categorical uniform draws labels with the frequencies of q_probs."""
    from apgs_b200.dmm import kernels
    S, B, N, K, D, H = 768, 3, 64, 4, 2, 32
    P, (_, enc_local, _, dec) = _mods(K, D, H, seed=5)
    x = torch.randn(B, N, D) * 2
    mu = (torch.randn(1, B, K, D) * 3).repeat(S, 1, 1, 1).contiguous()
    z_f = torch.nn.functional.one_hot(torch.randint(0, K, (1, B, N)), K).float().repeat(S, 1, 1, 1).contiguous()
    xg, mg, zg = x.to(DEV), mu.to(DEV), z_f.to(DEV)
    s1 = torch.tensor([77, 1], dtype=torch.int64, device=DEV)
    s2 = torch.tensor([78, 1], dtype=torch.int64, device=DEV)
    u = torch.rand(S, B, N, device=DEV)
    plain = kernels.local_encoder_fused(xg, mg, enc_local, uniforms=u, z_old=zg)
    drawn = kernels.local_encoder_fused(xg, mg, enc_local, uniforms=u, z_old=zg, seed=s1, decoder=dec.recon_mu)
    for a, b in zip(plain, drawn[:6]):
        assert torch.equal(a, b)
    # teacher-forced assignment: the concentrations of a point are the same for every particle
    q, z, c1, c0, _, _, beta, h2 = kernels.local_encoder_fused(xg, mg, enc_local, z_forced=zg, seed=s1, decoder=dec.recon_mu)
    again = kernels.local_encoder_fused(xg, mg, enc_local, z_forced=zg, seed=s1, decoder=dec.recon_mu)
    other = kernels.local_encoder_fused(xg, mg, enc_local, z_forced=zg, seed=s2, decoder=dec.recon_mu)
    assert torch.equal(beta, again[6]) and torch.equal(h2, again[7])
    assert float((beta != other[6]).float().mean()) > 0.99
    assert tuple(beta.shape) == (S, B, N, 1) and tuple(h2.shape) == (S, B, N, D)
    fi = torch.finfo(torch.float32)
    assert float(beta.min()) >= fi.tiny and float(beta.max()) <= 1.0 - fi.eps
    assert torch.equal(h2, kernels.decoder_hidden2(beta, dec.recon_mu))
    a, b = c1[0, :, :, 0].double(), c0[0, :, :, 0].double()
    bt = beta[..., 0].double()
    mean_ref, var_ref = a / (a + b), a * b / ((a + b) ** 2 * (a + b + 1))
    zm = (bt.mean(0) - mean_ref).abs() / (var_ref / S).sqrt()
    assert float(zm.max()) < 4.5, float(zm.max())
    m4 = ((bt - mean_ref) ** 4).mean(0)
    zv = (bt.var(0, unbiased=False) - var_ref).abs() / ((m4 - var_ref ** 2).clamp_min(1e-30) / S).sqrt()
    assert float(zv.max()) < 5.0, float(zv.max())
    # KS distance against torch's Beta for a few points (two-sample, n = m = S: D_crit(1e-4) ~ 2.2 / sqrt(S/2) ... use 0.12)
    torch.manual_seed(0)
    for (bb, nn) in ((0, 0), (1, 5), (2, 63)):
        ref = torch.distributions.Beta(c1[0, bb, nn, 0], c0[0, bb, nn, 0]).sample((S,)).double().cpu().sort().values
        got = bt[:, bb, nn].cpu().sort().values
        allv = torch.cat((ref, got)).sort().values
        d = (torch.searchsorted(ref, allv, right=True).double() / S - torch.searchsorted(got, allv, right=True).double() / S).abs().max()
        assert float(d) < 0.12, (bb, nn, float(d))
    resid = (bt - mean_ref) / var_ref.sqrt()
    for dim in (0, 2):
        p, q_ = resid.narrow(dim, 0, resid.shape[dim] - 1), resid.narrow(dim, 1, resid.shape[dim] - 1)
        corr = float((p * q_).mean())
        assert abs(corr) < 4.5 / p.numel() ** 0.5, (dim, corr)
    # (iv) the categorical draw from the in-kernel uniform
    qs, zs, *_ = kernels.local_encoder_fused(xg, mg, enc_local, seed=s1)
    freq = zs.double().mean(0)
    qn = (qs[0] / qs[0].sum(-1, keepdim=True)).double()
    se = (qn * (1 - qn) / S).sqrt().clamp_min(1e-4)
    assert float(((freq - qn).abs() / se).max()) < 4.5


def test_dmm_objective_with_in_kernel_draws_matches_torch_draws_statistically():
    """apg_objective with the draws made in the kernels vs torch draws: mean ESS / density / log Z agree within 5 standard errors."""
    from apgs_b200.dmm import objectives, apg_objective, init_apg_models
    from apgs_b200.resampler import Resampler
    K, D, N, S, B, sweeps = 4, 2, 200, 50, 48, 4
    torch.manual_seed(0)
    models = init_apg_models(K, D, 32, 8, 32, 32, 0.5, True, DEV)
    x = od.simulate(B, N, K, D).to(DEV)
    flags = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}
    out = {}
    for on in (True, False):
        objectives.rng_in_kernels = on
        try:
            torch.manual_seed(1)
            with torch.no_grad():
                trs = [apg_objective(models, x, K, flags, sweeps, Resampler('systematic', S, True, DEV)) for _ in range(6)]
        finally:
            objectives.rng_in_kernels = True
        out[on] = {k: torch.stack([t[k].double().reshape(sweeps, -1).mean(-1) for t in trs]) for k in ('ess', 'density', 'log_z')}
    for k in ('ess', 'density', 'log_z'):
        a, b = out[True][k], out[False][k]
        se = (a.var(0) / a.shape[0] + b.var(0) / b.shape[0]).sqrt().clamp_min(1e-9)
        zmax = float(((a.mean(0) - b.mean(0)).abs() / se).max())
        assert zmax < 5.0, (k, zmax)


@pytest.mark.parametrize("M_shape,H,H2,Z", [((100, 256, 4), 12, 32, 2), ((3, 5, 4), 14, 32, 3), ((7,), 5, 9, 1), ((2, 300), 32, 40, 4)])
def test_mu_heads_in_one_launch_match_the_modules(M_shape, H, H2, Z):
    """apgs_dmm_mu_heads_f32 against the reference's nn.Sequential heads (apgs/dmm/models.py: mean_mu / mean_log_sigma,
    forward :43-44) in fp64 and torch fp32: W2 tanh(W1 v + b1) + b2 and exp(...), widths that are not multiples of 4, row counts
    that are not multiples of the CTA."""
    import copy
    import torch.nn as nn
    from apgs_b200.dmm import kernels
    torch.manual_seed(H + H2)
    mean = nn.Sequential(nn.Linear(H, H2), nn.Tanh(), nn.Linear(H2, Z)).to(DEV)
    lsig = nn.Sequential(nn.Linear(H, H2), nn.Tanh(), nn.Linear(H2, Z)).to(DEV)
    v = torch.randn(*M_shape, H, device=DEV) * 2
    assert kernels.mu_heads_supported(mean, lsig)
    with torch.no_grad():
        qm, qs = kernels.mu_heads(v, mean, lsig)
        rm32, rs32 = mean(v), lsig(v).exp()
        rm, rs = copy.deepcopy(mean).double()(v.double()), copy.deepcopy(lsig).double()(v.double()).exp()
    assert tuple(qm.shape) == tuple(M_shape) + (Z,) and tuple(qs.shape) == tuple(qm.shape)
    em, es = float((qm.double() - rm).abs().max()), float(((qs.double() - rs).abs() / rs).max())
    em32, es32 = float((rm32.double() - rm).abs().max()), float(((rs32.double() - rs).abs() / rs).max())
    assert em <= 2e-6 and es <= 2e-6, (em, es)
    assert em <= 4 * em32 + 3e-7 and es <= 4 * es32 + 3e-7, (em, em32, es, es32)
    assert not kernels.mu_heads_supported(nn.Sequential(nn.Linear(40, H2), nn.Tanh(), nn.Linear(H2, Z)),
                                          nn.Sequential(nn.Linear(40, H2), nn.Tanh(), nn.Linear(H2, Z)))


def test_enc_mu_fused_heads_switch():
    """Enc_rws_mu / Enc_apg_mu with and without the one-launch heads (module switch `fused_heads`); training mode keeps the
    differentiable modules."""
    from apgs_b200.dmm import models as M
    torch.manual_seed(4)
    K, D, N, S, B = 4, 2, 50, 6, 5
    enc = M.Enc_rws_mu(K, D, 32, 8).to(DEV)
    x = torch.randn(S, B, N, D, device=DEV)
    priors = (torch.zeros(D, device=DEV), torch.ones(D, device=DEV) * 4)
    with torch.no_grad():
        m1, s1 = enc.params(x, K, priors)
        M._EncMu.fused_heads = False
        try:
            m0, s0 = enc.params(x, K, priors)
        finally:
            M._EncMu.fused_heads = True
    assert float((m1 - m0).abs().max()) <= 2e-6 and float(((s1 - s0).abs() / s0).max()) <= 2e-6
    q = enc(x, K, priors)                                            # parameters require grad: the graph survives
    assert q['means'].log_prob.requires_grad
