"""Maximum OBSERVED errors of the kernels against the oracle (== the reference on CPU) for every quantity whose test
tolerance is wider than 1e-5, and the worst log-weight error as a fraction of its 1e-5 bound (VERDICT r1 "weak" #2).
Writes a markdown table to stdout: python benchmarks/parity_errors.py > profiles/r2_parity_errors.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gmm as og, dmm as od  # noqa: E402
from apgs_b200.gmm import kernels as gk  # noqa: E402

DEV = "cuda:0"
rows = []


def rel(a, b, floor=0.0):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float(((a - b).abs() / (b.abs() + floor)).max())


def absmax(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


class Gen:
    def __init__(self, K, D):
        pr = og.make_prior(K, D)
        self.prior_alpha, self.prior_beta = pr["alpha"].to(DEV), pr["beta"].to(DEV)
        self.prior_mu, self.prior_nu, self.prior_pi = pr["mu"].to(DEV), pr["nu"].to(DEV), pr["pi"].to(DEV)


# ------------------------------------------------------------------ GMM z step (S = 1000, 16 instances, N = 60, H = 32)
S, B, N, K, D = 1000, 16, 60, 3, 2
torch.manual_seed(0)
P, prior = og.make_params(K, D, 32), og.make_prior(K, D)
x = og.simulate(B, N, K, D)
tau, mu = torch.rand(S, B, K, D) * 3 + 0.2, torch.randn(S, B, K, D) * 2
lab = torch.randint(0, K, (S, B, N))
z_old = torch.nn.functional.one_hot(lab, K).float()
u = torch.rand(S, B, N)
mlp = [P["enc_apg_z"][k].to(DEV) for k in ("pi_log_prob.0.weight", "pi_log_prob.0.bias", "pi_log_prob.2.weight", "pi_log_prob.2.bias")]
z_new, q, w = gk.z_step_fused(x.to(DEV), tau.to(DEV), mu.to(DEV), mlp, Gen(K, D), z_old=z_old.to(DEV), uniforms=u.to(DEV), want_q=True, parts=True, local=True)
xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
q_ref = og.enc_apg_z_probs(P["enc_apg_z"], xs, tau, mu)
ref = og.z_step_weights(xs, tau, mu, q_ref, prior, z_new.cpu(), z_old, parts=True)
rows.append(("GMM `q_probs` (fused Enc_apg_z, SFU ex2 / rcp)", "2e-5 rel", "%.2e rel" % rel(q, q_ref, 1e-7)))
rows.append(("GMM `log_q_local` per point", "2.5e-5 abs", "%.2e abs" % absmax(w.log_q_local, ref[2])))
rows.append(("GMM z step `log_w` / (1e-5 (|f| + |b|))", "<= 1", "%.3f" % float(((w.log_w.cpu() - ref[1]).abs() / (1e-5 * (ref[4].abs() + ref[5].abs()))).max())))
rows.append(("GMM z step density", "1e-5 rel", "%.2e rel" % rel(w.density, ref[3].sum(-1))))
# ------------------------------------------------------------------ GMM eta posterior / statistics / weights
heads = [P["enc_apg_eta"][k].to(DEV) for k in ("gamma.0.weight", "gamma.0.bias", "ob.0.weight", "ob.0.bias")]
*post, stats = gk.eta_posterior_fused(x.to(DEV), z_old.to(DEV), heads, Gen(K, D), center=mu.to(DEV))
rp = og.enc_apg_eta_posterior(P["enc_apg_eta"], xs, z_old, prior)
for name, a, b in zip(("alpha", "beta", "mu", "nu"), post, rp):
    rows.append(("GMM posterior `%s` (fused heads, label table)" % name, "2e-5 rel (+ cancellation term for beta / mu)", "%.2e rel" % rel(a, b, 1e-3)))
g_std, e_std = torch._standard_gamma(post[0]), torch.randn_like(post[2])
we, tau2, mu2 = gk.draw_logw_eta_stats(stats, mu.to(DEV), tuple(post), Gen(K, D), g_std, e_std, tau.to(DEV), mu.to(DEV), parts=True)
pc = [p.cpu() for p in post]
re_ = og.eta_step_weights(xs, z_old, pc, prior, tau2.cpu(), mu2.cpu(), tau, mu, parts=True)
mag = sum((og.log_likelihood(xs, z_old, t, m, True).abs() + og.eta_log_prior(t, m, prior).abs() + og.eta_log_q(t, m, pc).abs())
          for t, m in ((tau2.cpu(), mu2.cpu()), (tau, mu)))
rows.append(("GMM eta step `log_w` / (1e-5 x summed magnitudes)", "<= 1", "%.3f" % float(((we.log_w.cpu() - re_[0]).abs() / (1e-5 * mag)).max())))
# ------------------------------------------------------------------ DMM local encoder (C3 shape, 3 instances)
from apgs_b200.dmm import kernels as dk, models as DM  # noqa: E402
K, D, N, S, B = 4, 2, 200, 100, 3
torch.manual_seed(1)
Pd = od.make_params(K, D, 32, 8, 32, 32)
ms = DM.init_apg_models(K, D, 32, 8, 32, 32, 0.5, True, DEV)
for m, key in zip(ms, ("enc_rws_mu", "enc_apg_local", "enc_apg_mu", "dec")):
    m.load_state_dict({k: v.to(DEV) for k, v in Pd[key].items()})
x = od.simulate(B, N, K, D)
xs = x.unsqueeze(0).repeat(S, 1, 1, 1)
mu = torch.randn(S, B, K, D) * 3
zf = torch.nn.functional.one_hot(torch.randint(0, K, (S, B, N)), K).float()
with torch.no_grad():
    q, z, c1, c0, _, _ = dk.local_encoder_fused(x.to(DEV), mu.to(DEV), ms[1], z_forced=zf.to(DEV))
    beta = torch.rand(S, B, N, 1)
    h2 = ms[3].hidden2(beta.to(DEV))
q_ref = od.enc_local_probs(Pd["enc_apg_local"], xs, mu, K)
c1r, c0r = od.enc_local_cons(Pd["enc_apg_local"], xs, mu, zf)
rows.append(("DMM `q_probs` (fused Enc_apg_local)", "2e-5 rel", "%.2e rel" % rel(q, q_ref, 1e-7)))
rows.append(("DMM Beta concentrations", "2e-5 rel", "%.2e rel" % max(rel(c1, c1r), rel(c0, c0r))))
rows.append(("DMM decoder `hidden2` (unit vector)", "1e-5 abs", "%.2e abs" % absmax(h2, od.dec_hidden2(Pd["dec"], beta))))
# ------------------------------------------------------------------ one-launch heads and the packed-operand tensor-core layer
import copy  # noqa: E402
import torch.nn as nn  # noqa: E402
from apgs_b200.bmnist import kernels as bk  # noqa: E402
torch.manual_seed(2)
with torch.no_grad():
    v = torch.randn(100, 256, 4, 12, device=DEV) * 2
    hm = nn.Sequential(nn.Linear(12, 32), nn.Tanh(), nn.Linear(32, 2)).to(DEV)
    hs = nn.Sequential(nn.Linear(12, 32), nn.Tanh(), nn.Linear(32, 2)).to(DEV)
    qm, qs = dk.mu_heads(v, hm, hs)
    rm, rs = copy.deepcopy(hm).double()(v.double()), copy.deepcopy(hs).double()(v.double()).exp()
    rows.append(("DMM `q_mu` / `q_sigma` (one-launch heads) vs fp64 modules", "2e-6 abs / rel",
                 "%.2e abs / %.2e rel (torch fp32 modules: %.2e / %.2e)" % (absmax(qm, rm), rel(qs, rs), absmax(hm(v), rm), rel(hs(v).exp(), rs))))
    xc = torch.softmax(torch.randn(640, 4761, device=DEV) * 3, -1)
    wc, bc = torch.randn(400, 4761, device=DEV) / 69.0, torch.randn(400, device=DEV)
    ref = xc.double() @ wc.double().t() + bc.double()
    mag = xc.double().abs() @ wc.double().abs().t() + bc.double().abs()
    o = bk.linear_tc5_packed(bk.pack_rows(xc, 0), bk.pack_rows(wc, 1), bc)
    rows.append(("bmnist Enc_coor layer 4761 -> 400 (3xTF32 `tcgen05` from packed operands) / sum |x w|", "4e-6",
                 "%.2e (torch fp32 GEMM: %.2e)" % (float(((o.double() - ref).abs() / mag).max()),
                                                    float(((torch.nn.functional.linear(xc, wc, bc).double() - ref).abs() / mag).max()))))
    hid = torch.relu(o)
    cm = nn.Sequential(nn.Linear(400, 200), nn.ReLU(), nn.Linear(200, 2), nn.Tanh()).to(DEV)
    cs = nn.Sequential(nn.Linear(400, 200), nn.ReLU(), nn.Linear(200, 2)).to(DEV)
    w1t, b1 = bk.coor_heads_weights(cm[0].weight, cm[0].bias, cs[0].weight, cs[0].bias)
    qm, qs = bk.coor_heads(hid, w1t, b1, cm[2].weight, cm[2].bias, cs[2].weight, cs[2].bias)
    rm, rs = copy.deepcopy(cm).double()(hid.double()), copy.deepcopy(cs).double()(hid.double()).exp()
    rows.append(("bmnist `q_mean` / `q_std` of z_where (one-launch heads) vs fp64 modules", "2e-6 abs / rel",
                 "%.2e abs / %.2e rel (torch fp32 modules: %.2e / %.2e)" % (absmax(qm, rm), rel(qs, rs), absmax(cm(hid), rm), rel(cs(hid).exp(), rs))))
print("# r2: maximum observed errors against the oracle (`python benchmarks/parity_errors.py` on a B200)\n")
print("Every log-WEIGHT is tested to the north_star's 1e-5 of the magnitudes summed (SURVEY.md H2); the intermediates below have wider test "
      "tolerances because they go through the SFU approximations (`ex2.approx`, `rcp.approx`, `lg2.approx`, ~2^-22 each) of the fused MLPs.  "
      "Observed maxima over 16 instances x 1000 particles (GMM) / 3 instances x 100 particles (DMM):\n")
print("| quantity | test tolerance | max observed |\n|---|---|---|")
for r in rows:
    print("| %s | %s | %s |" % r)
