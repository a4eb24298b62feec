"""CPU: the oracle (oracle/) against the golden fixtures generated from the reference."""
import numpy as np
import pytest
import torch

from oracle import resampler as orc
from oracle import gmm as og


def _cases(g):
    return [(i, int(S), int(B)) for i, (S, B, _) in enumerate(g["cases"])]


def test_systematic_ancestors_bit_exact(golden):
    g = golden("resampler")
    for ci, S, B in _cases(g):
        a = orc.systematic_ancestors(g["c%d_w" % ci], g["c%d_u_sys" % ci])
        assert a.shape == (S, B)
        assert np.array_equal(a, g["c%d_anc_sys" % ci]), "case %d" % ci


def test_multinomial_ancestors_bit_exact(golden):
    g = golden("resampler")
    for ci, S, B in _cases(g):
        a = orc.multinomial_ancestors(g["c%d_p_mul" % ci], g["c%d_u_mul" % ci])
        assert np.array_equal(a, g["c%d_anc_mul" % ci]), "case %d" % ci
        p = orc.categorical_probs(torch.from_numpy(g["c%d_w" % ci])).numpy()
        np.testing.assert_allclose(p, g["c%d_p_mul" % ci], rtol=3e-7)


def test_gather_and_stats(golden):
    g = golden("resampler")
    for ci, S, B in _cases(g):
        if "c%d_var4" % ci in g:
            assert np.array_equal(orc.gather_ancestors(g["c%d_var4" % ci], g["c%d_anc_sys" % ci]), g["c%d_res4" % ci])
            assert np.array_equal(orc.gather_ancestors(g["c%d_var5" % ci], g["c%d_anc_mul" % ci]), g["c%d_res5" % ci])
        w, ess, lse = orc.weight_stats(torch.from_numpy(g["c%d_log_w" % ci]))
        np.testing.assert_allclose(w.numpy(), g["c%d_w" % ci], rtol=2e-6, atol=1e-30)
        np.testing.assert_allclose(ess.numpy(), g["c%d_ess" % ci], rtol=1e-5)
        np.testing.assert_allclose(lse.numpy(), g["c%d_lse" % ci], rtol=1e-6, atol=1e-6)


def test_minus_inf_weights_never_selected(golden):
    g = golden("resampler")
    lw, a_s, a_m = g["c2_log_w"], g["c2_anc_sys"], g["c2_anc_mul"]
    dead = np.where(np.isinf(lw[:, 1]))[0]
    assert len(dead) > 0
    assert not np.isin(a_s[:, 1], dead).any() and not np.isin(a_m[:, 1], dead).any()


# ------------------------------------------------------------------------------------ GMM
def _params(g):
    P = {}
    for k in g.files:
        if k.startswith("P/"):
            _, mod, name = k.split("/", 2)
            P.setdefault(mod, {})[name] = torch.from_numpy(g[k])
    return P


def _t(g, k):
    return torch.from_numpy(g[k])


def test_gmm_nss_posterior(golden):
    g = golden("gmm")
    prior = og.make_prior()
    for tag in ("soft", "hard"):
        a, b, m, n = og.nss_posterior(_t(g, "g1_xp"), _t(g, "g1_%s_gamma" % tag), prior)
        for name, v in (("alpha", a), ("beta", b), ("mu", m), ("nu", n)):
            assert torch.equal(v, _t(g, "g1_%s_%s" % (tag, name))), (tag, name)


def test_gmm_eta_step_teacher_forced(golden):
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["eta_z"].shape[0]
    x = _t(g, "x").unsqueeze(0).repeat(S, 1, 1, 1)
    post = og.enc_apg_eta_posterior(P["enc_apg_eta"], x, _t(g, "eta_z"), prior)
    assert torch.equal(post[0], _t(g, "eta_alpha")) and torch.equal(post[1], _t(g, "eta_beta"))
    assert torch.equal(post[2], _t(g, "eta_mu_q"))
    log_w, log_q_f, log_p_f, ll_f = og.eta_step_weights(x, _t(g, "eta_z"), post, prior, _t(g, "eta_tau"), _t(g, "eta_mu"),
                                                        _t(g, "eta_tau_old"), _t(g, "eta_mu_old"))
    assert torch.equal(log_w, _t(g, "eta_log_w"))
    assert torch.equal(log_p_f, _t(g, "eta_density")[0])
    w = torch.softmax(log_w, 0)
    assert torch.equal(1. / (w ** 2).sum(0), _t(g, "eta_ess")[0])
    assert torch.equal((w * (-log_q_f)).sum(0).mean(), _t(g, "eta_loss")[0])


def test_gmm_z_step_teacher_forced(golden):
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["z_old"].shape[0]
    x = _t(g, "x").unsqueeze(0).repeat(S, 1, 1, 1)
    q_probs = og.enc_apg_z_probs(P["enc_apg_z"], x, _t(g, "z_tau"), _t(g, "z_mu"))
    # the fixture holds Categorical's renormalised copy (categorical.py:70)
    assert torch.equal(q_probs / q_probs.sum(-1, keepdim=True), _t(g, "z_q_probs"))
    lwl, log_w, log_q_f, dens = og.z_step_weights(x, _t(g, "z_tau"), _t(g, "z_mu"), q_probs, prior, _t(g, "z_new"), _t(g, "z_old"))
    assert torch.equal(log_w, _t(g, "z_log_w"))
    assert torch.equal(dens.sum(-1), _t(g, "z_density")[0])
    w_local = torch.softmax(lwl, 0)
    assert torch.equal((w_local * (-log_q_f)).sum(0).sum(-1).mean(), _t(g, "z_loss")[0])


def test_gmm_oneshot_and_joint_teacher_forced(golden):
    g = golden("gmm")
    P, prior = _params(g), og.make_prior()
    S = g["os_z"].shape[0]
    x = _t(g, "x").unsqueeze(0).repeat(S, 1, 1, 1)
    post = og.enc_rws_eta_posterior(P["enc_rws_eta"], x, prior)
    assert torch.equal(post[0], _t(g, "os_alpha")) and torch.equal(post[2], _t(g, "os_mu_q"))
    q_probs = og.enc_apg_z_probs(P["enc_apg_z"], x, _t(g, "os_tau"), _t(g, "os_mu"))
    log_w, log_q, log_p = og.oneshot_weights(x, post, prior, _t(g, "os_tau"), _t(g, "os_mu"), q_probs, _t(g, "os_z"))
    assert torch.equal(log_w, _t(g, "os_log_w")) and torch.equal(log_p, _t(g, "os_density")[0])
    # joint
    post = og.enc_apg_eta_posterior(P["enc_apg_eta"], x, _t(g, "j_z_old"), prior)
    q_probs = og.enc_apg_z_probs(P["enc_apg_z"], x, _t(g, "j_tau"), _t(g, "j_mu"))
    log_w, log_q_f, dens = og.joint_step_weights(x, post, prior, _t(g, "j_tau"), _t(g, "j_mu"), _t(g, "j_tau_old"),
                                                 _t(g, "j_mu_old"), q_probs, _t(g, "j_z"), _t(g, "j_z_old"))
    assert torch.equal(log_w, _t(g, "j_log_w")) and torch.equal(dens, _t(g, "j_density")[0])


@pytest.mark.parametrize("block", ["decomposed", "joint"])
@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_gmm_full_objective_reproduces_reference_trace(golden, block, strategy):
    """Same seed, same torch RNG calls -> the whole reference trace (configs[0]) bit for bit."""
    g = golden("gmm")
    S, sweeps = 10, 5
    torch.manual_seed(0)
    P = og.make_params()
    for mod in P:
        for k in P[mod]:
            assert torch.equal(P[mod][k], _params(g)[mod][k])
    x = _t(g, "full_x").repeat(S, 1, 1, 1)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    torch.manual_seed(7)
    tr = og.apg_objective(P, og.make_prior(), x, flags, sweeps, block, strategy)
    for k in ("loss", "ess", "density"):
        ref = _t(g, "full_%s_%s_%s" % (block, strategy, k))
        assert tr[k].shape == ref.shape
        torch.testing.assert_close(tr[k], ref, rtol=1e-6, atol=1e-6)


def test_gmm_gibbs_and_bpg_baselines_reproduce_reference_traces(golden):
    """gibbs_objective / bpg_objective (apgs/gmm/objectives.py:236-330) under the generating seeds."""
    g = golden("gmm")
    S, sweeps = 10, 5
    torch.manual_seed(0)
    P, prior = og.make_params(), og.make_prior()
    x = _t(g, "full_x")
    with torch.no_grad():
        torch.manual_seed(8)
        tr = og.gibbs_objective(P, prior, x.repeat(S, 1, 1, 1), sweeps)
        torch.testing.assert_close(tr["density"], _t(g, "gibbs_density"), rtol=1e-6, atol=1e-5)
        torch.manual_seed(9)
        tr = og.bpg_objective(P, prior, x.repeat(4 * S, 1, 1, 1), sweeps, "systematic")
        torch.testing.assert_close(tr["density"], _t(g, "bpg_density"), rtol=1e-6, atol=1e-5)


# ------------------------------------------------------------------------ STN / bmnist
def test_stn_closed_form_bit_exact(golden):
    from oracle import stn
    g = golden("stn")
    assert np.array_equal(stn.digit_to_frame(g["digit"], g["z_where"]), g["pasted"])
    assert np.array_equal(stn.frame_to_digit(g["frames"], g["z_where"]), g["cropped"])
    recon, ll = stn.recon_and_ll(g["digit"], g["z_where"], g["frames"])
    assert np.array_equal(recon.numpy(), g["recon"]) and np.array_equal(ll.numpy(), g["ll"])
    # the torch-vectorised restatement used by the sweep oracle agrees bit for bit
    from oracle import bmnist as ob
    assert torch.equal(ob.digit_to_frame(_t(g, "digit"), _t(g, "z_where")), _t(g, "pasted"))
    assert torch.equal(ob.frame_to_digit(_t(g, "frames"), _t(g, "z_where")), _t(g, "cropped"))


def test_bmnist_step_weights(golden):
    from oracle import bmnist as ob
    g = golden("bmnist")
    prior = ob.make_prior()
    for tag in ("w0", "w1", "w5"):
        pm = {"log_q_f": _t(g, tag + "_log_q_f"), "log_p_f": _t(g, tag + "_log_p_f")}
        z_old = _t(g, tag + "_z_old")
        has_prev = (tag + "_z_old_prev") in g.files
        lqb = sum(ob.dists.normal_log_prob(z_old[:, :, k], _t(g, tag + "_q_mean")[:, :, k], _t(g, tag + "_q_std")[:, :, k]).sum(-1).unsqueeze(-1)
                  for k in range(3))
        lpb = torch.cat([ob.coor_log_prior(prior, z_old[:, :, k], _t(g, tag + "_z_old_prev")[:, :, k] if has_prev else None).unsqueeze(-1)
                         for k in range(3)], -1).sum(-1)
        lqb = torch.cat([ob.dists.normal_log_prob(z_old[:, :, k], _t(g, tag + "_q_mean")[:, :, k], _t(g, tag + "_q_std")[:, :, k]).sum(-1).unsqueeze(-1)
                         for k in range(3)], -1).sum(-1)
        pm["log_q_b"], pm["log_p_b"] = lqb, lpb
        lw = ob.where_step_weight(pm, _t(g, tag + "_ll_f").squeeze(-1), _t(g, tag + "_ll_b").squeeze(-1))
        assert torch.equal(lw, _t(g, tag + "_log_w")), tag
    lw, lq, dens = ob.what_step_weight(_t(g, "wh_q_mu"), _t(g, "wh_q_std"), _t(g, "wh_z_what"), _t(g, "wh_z_what_old"),
                                       _t(g, "wh_ll_f"), _t(g, "wh_ll_b"))
    assert torch.equal(lw, _t(g, "wh_log_w")) and torch.equal(dens, _t(g, "wh_dens"))


@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_bmnist_full_objective_reproduces_reference_trace(golden, strategy):
    from oracle import bmnist as ob
    g = golden("bmnist")
    S, B, T, K, sweeps, FP, DP = [int(v) for v in g["dims"]]
    torch.manual_seed(0)
    P = ob.make_params()
    frames = _t(g, "frames").repeat(S, 1, 1, 1, 1)
    mm = _t(g, "template").repeat(S, B, K, 1, 1)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    torch.manual_seed(2)
    with torch.no_grad():
        tr = ob.apg_objective(P, ob.make_prior(), frames, K, flags, sweeps, mm, strategy)
    for k in ("loss_phi", "loss_theta", "ess", "density"):
        torch.testing.assert_close(tr[k], _t(g, "full_%s_%s" % (strategy, k)), rtol=1e-5, atol=1e-3)


# ---------------------------------------------------------------------------------- DMM
def _dmm_params(g):
    P = {}
    for k in g.files:
        if k.startswith("P/"):
            _, mod, name = k.split("/", 2)
            P.setdefault(mod, {})[name] = torch.from_numpy(g[k])
    return P


def test_dmm_step_weights_teacher_forced(golden):
    from oracle import dmm as od
    g = golden("dmm")
    S, B, N, K, D, sweeps = [int(v) for v in g["dims"]]
    P, prior = _dmm_params(g), od.make_prior(K, D)
    x = _t(g, "x").unsqueeze(0).repeat(S, 1, 1, 1)
    lw = od.mu_step_weights(prior, P["dec"], x, _t(g, "mu_z"), _t(g, "mu_hidden2"), _t(g, "mu_q_mu"), _t(g, "mu_q_sigma"),
                            _t(g, "mu_mu"), _t(g, "mu_mu_old"))[0]
    assert torch.equal(lw, _t(g, "mu_ref_log_w"))
    lw = od.local_step_weights(prior, P["dec"], x, _t(g, "lo_mu"), _t(g, "lo_q_probs"), _t(g, "lo_z"), _t(g, "lo_beta"),
                               _t(g, "lo_con1"), _t(g, "lo_con0"), _t(g, "lo_hidden2"), _t(g, "lo_z_old"), _t(g, "lo_beta_old"),
                               _t(g, "lo_con1_old"), _t(g, "lo_con0_old"), _t(g, "lo_hidden2_old"))[1]
    assert torch.equal(lw, _t(g, "lo_ref_log_w"))
    lw = od.oneshot_weights(prior, P["dec"], x, _t(g, "os_q_mu"), _t(g, "os_q_sigma"), _t(g, "os_mu"), _t(g, "os_q_probs"),
                            _t(g, "os_z"), _t(g, "os_beta"), _t(g, "os_con1"), _t(g, "os_con0"), _t(g, "os_hidden2"))[0]
    assert torch.equal(lw, _t(g, "os_ref_log_w"))
    # encoders reproduce the recorded proposal parameters from the stored weights
    q_mu, q_sigma = od.enc_mu_params(P["enc_apg_mu"], torch.cat((x, _t(g, "mu_z")), -1), prior)
    assert torch.equal(q_mu, _t(g, "mu_q_mu")) and torch.equal(q_sigma, _t(g, "mu_q_sigma"))
    assert torch.equal(od.enc_local_probs(P["enc_apg_local"], x, _t(g, "lo_mu"), K), _t(g, "lo_q_probs"))


@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_dmm_full_objective_reproduces_reference_trace(golden, strategy):
    from oracle import dmm as od
    g = golden("dmm")
    S, B, N, K, D, sweeps = [int(v) for v in g["dims"]]
    torch.manual_seed(0)
    P = od.make_params(K, D)
    for mod in P:
        for k in P[mod]:
            assert torch.equal(P[mod][k], _dmm_params(g)[mod][k])
    x = _t(g, "x").repeat(S, 1, 1, 1)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    torch.manual_seed(2)
    with torch.no_grad():
        tr = od.apg_objective(P, od.make_prior(K, D), x, K, flags, sweeps, strategy)
    for k in ("loss_phi", "loss_theta", "ess", "density"):
        torch.testing.assert_close(tr[k], _t(g, "full_%s_%s" % (strategy, k)), rtol=1e-6, atol=1e-5)
