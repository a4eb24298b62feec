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
