"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, CPU, fixed seeds) through tests/golden/ref_harness.py.

The reference ships no tests or vectors of its own, so these fixtures ARE the
pin for the oracle: tests/test_oracle_golden.py checks oracle/ against them on CPU, and the
-m gpu tests check the CUDA path against the same files.

    python tests/golden/make_golden.py [section ...]     sections: resampler gmm dmm bmnist stn

Runs only where /root/reference exists (the build container); the fixtures are committed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness  # noqa: E402

ref_harness.install()


def _np(t):
    if isinstance(t, (tuple, list)):
        return [_np(v) for v in t]
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: _np(v) for k, v in arrays.items()})
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


# ------------------------------------------------------------------------------- resampler
def section_resampler():
    from apgs.resampler import Resampler
    out = {}
    cases = [(10, 20, 1.0), (1, 5, 1.0), (64, 3, 8.0), (257, 33, 0.1), (1000, 16, 2.0), (3000, 2, 12.0)]
    out["cases"] = np.array(cases, dtype=np.float64)
    for ci, (S, B, sig) in enumerate(cases):
        torch.manual_seed(100 + ci)
        log_w = torch.randn(S, B) * sig
        if ci == 2:
            log_w[::3, 1] = -float("inf")           # particles that must never be selected
        w = torch.softmax(log_w, 0)
        # systematic: apgs/resampler.py:26-37
        torch.manual_seed(200 + ci)
        anc_sys = Resampler("systematic", S, False, None).sample_ancestral_index(log_w)
        torch.manual_seed(200 + ci)
        u_sys = torch.rand(B, 1)
        # multinomial: apgs/resampler.py:39-41 ; uniforms re-drawn from the same generator state
        torch.manual_seed(300 + ci)
        anc_mul = Resampler("multinomial", S, False, None).sample_ancestral_index(log_w)
        torch.manual_seed(300 + ci)
        u_mul = torch.rand(B, S, dtype=torch.float64)
        # gather: apgs/resampler.py:46-54
        out.update({
            "c%d_log_w" % ci: log_w, "c%d_w" % ci: w, "c%d_u_sys" % ci: u_sys.reshape(-1),
            "c%d_anc_sys" % ci: anc_sys.to(torch.int32), "c%d_u_mul" % ci: u_mul,
            "c%d_anc_mul" % ci: anc_mul.to(torch.int32),
            # what Categorical.__init__ hands to torch.multinomial (categorical.py:70)
            "c%d_p_mul" % ci: (w.transpose(0, 1) / w.transpose(0, 1).sum(-1, keepdim=True)).transpose(0, 1).contiguous(),
            "c%d_ess" % ci: 1. / (w ** 2).sum(0), "c%d_lse" % ci: torch.logsumexp(log_w, 0),
        })
        if S * B <= 2000:        # gather fixtures only for the small cases (keeps the file small)
            var4 = torch.randn(S, B, 3, 2)
            var5 = torch.randn(S, B, 2, 3, 2)
            r = Resampler("systematic", S, False, None)
            out.update({"c%d_var4" % ci: var4, "c%d_var5" % ci: var5,
                        "c%d_res4" % ci: r.resample_4dims(var4, anc_sys),
                        "c%d_res5" % ci: r.resample_5dims(var5, anc_mul)})
    save("resampler", **out)


# ------------------------------------------------------------------------------------- GMM
def section_gmm():
    from apgs.resampler import Resampler
    from apgs.gmm.apg_training import init_apg_models
    from apgs.gmm import objectives as O
    from apgs.gmm.kls_gmm import posterior_eta
    from oracle import gmm as og
    K, D, H, N = 3, 2, 32, 60
    out = {}

    # G1: posterior_eta on random statistics, incl. an empty cluster (quirk Q1)
    torch.manual_seed(1)
    S, B = 4, 3
    xp = torch.randn(S, B, N, D)
    gam = torch.softmax(torch.randn(S, B, N, K) * 2, -1)
    gam_hard = torch.nn.functional.one_hot(torch.randint(0, K - 1, (S, B, N)), K).float()  # cluster K-1 empty
    gen = og.make_prior(K, D)
    for tag, g in (("soft", gam), ("hard", gam_hard)):
        a, b, m, n = posterior_eta(xp, g, gen["alpha"], gen["beta"], gen["mu"], gen["nu"])
        out.update({"g1_%s_gamma" % tag: g, "g1_%s_alpha" % tag: a, "g1_%s_beta" % tag: b,
                    "g1_%s_mu" % tag: m, "g1_%s_nu" % tag: n})
    out["g1_xp"] = xp

    # teacher-forced steps of the reference sweep
    S, B = 6, 5
    torch.manual_seed(0)
    models = init_apg_models(K, D, H, False, None)
    enc_rws_eta, enc_apg_z, enc_apg_eta, generative = models
    for n_, m_ in (("enc_rws_eta", enc_rws_eta), ("enc_apg_z", enc_apg_z), ("enc_apg_eta", enc_apg_eta)):
        for k, v in m_.state_dict().items():
            out["P/%s/%s" % (n_, k)] = v
    torch.manual_seed(1)
    x = og.simulate(B, N, K, D).repeat(S, 1, 1, 1)
    out["x"] = x[0]
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}

    def fresh():
        return {"loss": [], "ess": [], "E_tau": [], "E_mu": [], "E_z": [], "density": []}

    with torch.no_grad():
        torch.manual_seed(2)
        tr = fresh()
        log_w, tau, mu, z, tr = O.oneshot(enc_rws_eta, enc_apg_z, generative, x, tr, flags)
        q_eta = enc_rws_eta(x, prior_ng=generative.prior_ng, sampled=False, tau_old=tau, mu_old=mu)
        q_z = enc_apg_z(x, tau=tau, mu=mu, sampled=False, z_old=z)
        out.update({"os_log_w": log_w, "os_tau": tau, "os_mu": mu, "os_z": z,
                    "os_alpha": q_eta["precisions"].dist.concentration, "os_beta": q_eta["precisions"].dist.rate,
                    "os_mu_q": q_eta["means"].dist.loc, "os_q_probs": q_z["states"].dist.probs,
                    "os_loss": tr["loss"][0], "os_ess": tr["ess"][0], "os_density": tr["density"][0]})
        tau, mu, z = O.resample_variables(Resampler("systematic", S, False, None), tau, mu, z, log_weights=log_w)

        # eta step (apgs/gmm/objectives.py:161-195)
        torch.manual_seed(3)
        tr = fresh()
        log_w_eta, tau_new, mu_new, tr = O.apg_update_eta(enc_apg_eta, generative, x, z, tau, mu, tr, flags)
        q = enc_apg_eta(x, z=z, prior_ng=generative.prior_ng, sampled=False, tau_old=tau_new, mu_old=mu_new)
        out.update({"eta_z": z, "eta_tau_old": tau, "eta_mu_old": mu, "eta_tau": tau_new, "eta_mu": mu_new,
                    "eta_log_w": log_w_eta, "eta_alpha": q["precisions"].dist.concentration,
                    "eta_beta": q["precisions"].dist.rate, "eta_mu_q": q["means"].dist.loc,
                    "eta_scale_q": q["means"].dist.scale,
                    "eta_loss": tr["loss"][0], "eta_ess": tr["ess"][0], "eta_density": tr["density"][0]})

        # z step (apgs/gmm/objectives.py:197-227)
        torch.manual_seed(4)
        tr2 = {"loss": [torch.zeros(1)], "ess": [], "E_tau": [], "E_mu": [], "E_z": [],
               "density": [torch.zeros(1, S, B)]}
        log_w_z, z_new, tr2 = O.apg_update_z(enc_apg_z, generative, x, tau_new, mu_new, z, tr2, flags)
        qz = enc_apg_z(x, tau=tau_new, mu=mu_new, sampled=False, z_old=z)
        out.update({"z_tau": tau_new, "z_mu": mu_new, "z_old": z, "z_new": z_new, "z_log_w": log_w_z,
                    "z_q_probs": qz["states"].dist.probs, "z_loss": tr2["loss"][0], "z_density": tr2["density"][0]})

        # joint step (apgs/gmm/objectives.py:112-158); mode flag off (the reference's mode branch is broken)
        torch.manual_seed(5)
        tr3 = fresh()
        log_w_j, tau_j, mu_j, z_j, tr3 = O.apg_update_joint(enc_apg_z, enc_apg_eta, generative, x, z, tau, mu, tr3, flags)
        out.update({"j_z_old": z, "j_tau_old": tau, "j_mu_old": mu, "j_tau": tau_j, "j_mu": mu_j, "j_z": z_j,
                    "j_log_w": log_w_j, "j_loss": tr3["loss"][0], "j_ess": tr3["ess"][0],
                    "j_density": tr3["density"][0]})

    # whole-objective traces under one seed (BASELINE.json configs[0]: S=10, B=20, 5 sweeps)
    S, B, sweeps = 10, 20, 5
    torch.manual_seed(1)
    xx = og.simulate(B, N, K, D)
    out["full_x"] = xx
    for block in ("decomposed", "joint"):
        for strategy in ("systematic", "multinomial"):
            torch.manual_seed(7)
            tr = O.apg_objective(models, xx.repeat(S, 1, 1, 1), flags, sweeps, block, Resampler(strategy, S, False, None))
            for k in ("loss", "ess", "density"):
                out["full_%s_%s_%s" % (block, strategy, k)] = tr[k]

    # baselines around the path: Gibbs sweeps and BPG (prior proposals, S * bpg_factor particles)
    dflags = {"loss_required": False, "ess_required": False, "mode_required": False, "density_required": True}
    P = {n_: {k: v for k, v in m_.state_dict().items()} for n_, m_ in
         (("enc_rws_eta", enc_rws_eta), ("enc_apg_z", enc_apg_z), ("enc_apg_eta", enc_apg_eta))}
    with torch.no_grad():
        torch.manual_seed(8)
        tr = O.gibbs_objective(models, xx.repeat(S, 1, 1, 1), dflags, sweeps)
        torch.manual_seed(8)
        mine = og.gibbs_objective(P, gen, xx.repeat(S, 1, 1, 1), sweeps)
        assert torch.equal(tr["density"], mine["density"]), "oracle gibbs_objective != reference"
        out["gibbs_density"] = tr["density"]
        S_bpg = S * 4
        torch.manual_seed(9)
        tr = O.bpg_objective(models, xx.repeat(S_bpg, 1, 1, 1), dflags, sweeps, Resampler("systematic", S_bpg, False, None))
        torch.manual_seed(9)
        mine = og.bpg_objective(P, gen, xx.repeat(S_bpg, 1, 1, 1), sweeps, "systematic")
        assert torch.equal(tr["density"], mine["density"]), "oracle bpg_objective != reference"
        out["bpg_density"] = tr["density"]
    save("gmm", **out)


# ------------------------------------------------------------------------------------- STN
def section_stn():
    from apgs.bmnist.affine_transformer import Affine_Transformer
    from apgs.bmnist.models import MBern_log_prob
    FP, DP = 96, 28
    AT = Affine_Transformer(FP, DP, False, None)
    torch.manual_seed(11)
    S, B, T, K = 2, 2, 3, 3
    z = torch.randn(S, B, T, K, 2) * 0.6
    z[0, 0, 0, 0] = torch.tensor([1.4, -1.3])       # partly outside the frame
    z[1, 1, 2, 2] = torch.tensor([-2.5, 0.2])       # completely outside along x
    z[0, 1, 1, 1] = torch.tensor([0.0, 0.0])
    digit = torch.rand(S, B, K, DP, DP)
    frames = torch.rand(S, B, T, FP, FP)
    z_in = z.clone()
    pasted = AT.digit_to_frame(digit, z)
    cropped = AT.frame_to_digit(frames, z)
    assert torch.equal(z, z_in)                     # the reference does not mutate z_where
    recon = torch.clamp(pasted.sum(-3), min=0.0, max=1.0)
    ll = MBern_log_prob(recon, frames)
    save("stn", z_where=z, digit=digit, frames=frames, pasted=pasted, cropped=cropped, recon=recon, ll=ll)


# ---------------------------------------------------------------------------------- bmnist
def section_bmnist():
    """Reference apg_objective (bmnist) under fixed seeds: whole traces + every per-step log-weight (captured by
    wrapping the reference's resample_variables) + the oracle's step records, which are asserted to reproduce them."""
    from apgs.resampler import Resampler
    from apgs.bmnist.apg_training import init_models
    import apgs.bmnist.objectives as O
    from apgs.bmnist.affine_transformer import Affine_Transformer
    from oracle import bmnist as ob
    FP, DP, K = 96, 28, 3
    S, B, T, sweeps = 4, 3, 3, 3
    out = {"dims": np.array([S, B, T, K, sweeps, FP, DP])}
    torch.manual_seed(0)
    models = init_models(FP, DP, 400, 400, 2, 10, False, None, None, None)
    torch.manual_seed(0)
    P = ob.make_params()
    torch.manual_seed(1)
    frames1 = ob.simulate(B, T, K)
    frames = frames1.repeat(S, 1, 1, 1, 1)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, DP), torch.linspace(-1, 1, DP), indexing="ij")
    tmpl = torch.exp(-(xx ** 2 + yy ** 2) * 3).clamp(0, 1)
    mm = tmpl.repeat(S, B, K, 1, 1)
    out["frames"], out["template"] = frames1, tmpl
    AT = Affine_Transformer(FP, DP, False, None)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    ref_lw = []
    orig = O.resample_variables

    def rec(resampler, z_where, z_what, log_weights):
        ref_lw.append(log_weights.clone())
        return orig(resampler, z_where, z_what, log_weights)
    for strategy in ("systematic", "multinomial"):
        del ref_lw[:]
        O.resample_variables = rec
        with torch.no_grad():
            torch.manual_seed(2)
            tr = O.apg_objective(models, AT, frames, K, flags, sweeps, Resampler(strategy, S, False, None), mm)
            O.resample_variables = orig
            torch.manual_seed(2)
            recd = []
            mine = ob.apg_objective(P, ob.make_prior(), frames, K, flags, sweeps, mm, strategy, record=recd)
        for k in ("loss_phi", "loss_theta", "ess", "density"):
            assert torch.equal(tr[k], mine[k]), (strategy, k)
            out["full_%s_%s" % (strategy, k)] = tr[k]
        assert len(ref_lw) == len(recd)
        for i, (a, r) in enumerate(zip(ref_lw, recd)):
            assert torch.equal(a, r["log_w"]), (strategy, i)
        if strategy != "systematic":
            continue
        # step fixtures (inputs from the oracle's record, log-weights from the reference itself)
        ws = [i for i, r in enumerate(recd) if r["step"] == "where"]
        for tag, i in (("w0", ws[0]), ("w1", ws[1]), ("w5", ws[-1])):
            r = recd[i]
            pm = r["pm"]
            out.update({tag + "_q_mean": pm["q_mean"], tag + "_q_std": pm["q_std"], tag + "_z_new": pm["z_where"],
                        tag + "_z_old": r["z_old_t"], tag + "_ll_f": r["ll_f"], tag + "_ll_b": r["ll_b"],
                        tag + "_log_w": ref_lw[i], tag + "_log_q_f": pm["log_q_f"], tag + "_log_p_f": pm["log_p_f"],
                        tag + "_z_what": r["z_what"]})
            if r["z_old_t_1"] is not None:
                out[tag + "_z_old_prev"] = r["z_old_t_1"]
                prev = recd[i - 1]["pm"]["z_where"]
                out[tag + "_z_prev"] = prev
        i = [i for i, r in enumerate(recd) if r["step"] == "what"][0]
        r = recd[i]
        out.update({"wh_q_mu": r["q_mu"], "wh_q_std": r["q_std"], "wh_z_what": r["z_what"], "wh_z_what_old": r["z_what_old"],
                    "wh_ll_f": r["ll_f"], "wh_ll_b": r["ll_b"], "wh_log_w": ref_lw[i], "wh_log_q_f": r["log_q_f"],
                    "wh_dens": r["dens"], "wh_z_where": r["z_where"]})
        r = recd[0]
        out.update({"os_z_where": r["z_where"], "os_z_what": r["z_what"], "os_q_mu": r["q_mu"], "os_q_std": r["q_std"],
                    "os_ll": r["ll"], "os_log_p_where": r["log_p_where"], "os_log_q_where": r["log_q_where"],
                    "os_log_w": ref_lw[0], "os_log_p": r["log_p"], "os_log_q": r["log_q"]})
    save("bmnist", **out)


# ------------------------------------------------------------------------------------- DMM
def section_dmm():
    """Reference DMM apg_objective under fixed seeds: whole traces, every per-step log-weight (captured by wrapping the
    reference's resample_variables) and the oracle's step records, asserted to reproduce them bit for bit."""
    from apgs.resampler import Resampler
    from apgs.dmm.apg_training import init_apg_models
    import apgs.dmm.objectives as O
    from oracle import dmm as od
    K, D, N = 4, 2, 60
    S, B, sweeps = 6, 5, 3
    out = {"dims": np.array([S, B, N, K, D, sweeps])}
    torch.manual_seed(0)
    models = init_apg_models(K, D, 32, 8, 32, 32, 0.5, False, None)
    torch.manual_seed(0)
    P = od.make_params(K, D)
    for mod in P:
        for k, v in P[mod].items():
            out["P/%s/%s" % (mod, k)] = v
    torch.manual_seed(1)
    x1 = od.simulate(B, N, K, D)
    x = x1.repeat(S, 1, 1, 1)
    out["x"] = x1
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    ref_lw = []
    orig = O.resample_variables

    def rec(resampler, mu, z, beta, log_weights):
        ref_lw.append(log_weights.clone())
        return orig(resampler, mu, z, beta, log_weights)
    for strategy in ("systematic", "multinomial"):
        del ref_lw[:]
        O.resample_variables = rec
        with torch.no_grad():
            torch.manual_seed(2)
            tr = O.apg_objective(models, x, K, flags, sweeps, Resampler(strategy, S, False, None))
            O.resample_variables = orig
            torch.manual_seed(2)
            recd = []
            mine = od.apg_objective(P, od.make_prior(K, D), x, K, flags, sweeps, strategy, record=recd)
        for k in ("loss_phi", "loss_theta", "ess", "density"):
            assert torch.equal(tr[k], mine[k]), (strategy, k)
            out["full_%s_%s" % (strategy, k)] = tr[k]
        assert len(ref_lw) == len(recd)
        for i, (a, r) in enumerate(zip(ref_lw, recd)):
            assert torch.equal(a, r["log_w"]), (strategy, i)
        if strategy != "systematic":
            continue
        for tag, step in (("os", "oneshot"), ("mu", "mu"), ("lo", "local")):
            i = [j for j, r in enumerate(recd) if r["step"] == step][0]
            for k, v in recd[i].items():
                if torch.is_tensor(v):
                    out["%s_%s" % (tag, k)] = v
            out["%s_ref_log_w" % tag] = ref_lw[i]
    save("dmm", **out)


SECTIONS = {"resampler": section_resampler, "gmm": section_gmm, "stn": section_stn, "bmnist": section_bmnist,
            "dmm": section_dmm}

if __name__ == "__main__":
    todo = sys.argv[1:] or list(SECTIONS)
    for s in todo:
        SECTIONS[s]()
