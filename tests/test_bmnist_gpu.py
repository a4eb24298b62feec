"""GPU parity: spatial-transformer crop / paste (bit-exact: pure gathers) and the fused paste + Bernoulli
log-likelihood, z_where / z_what step weights, against the reference-generated golden fixtures and the oracle."""
import numpy as np
import pytest
import torch

from oracle import bmnist as ob
from oracle import stn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(g, k):
    return torch.from_numpy(g[k])


def _AT():
    from apgs_b200.bmnist.affine_transformer import Affine_Transformer
    return Affine_Transformer(96, 28, True, DEV)


def test_B1_B2_crop_paste_bit_exact_vs_reference_golden(golden):
    g = golden("stn")
    AT = _AT()
    z = _t(g, "z_where").to(DEV)
    z_in = z.clone()
    pasted = AT.digit_to_frame(_t(g, "digit").to(DEV), z)
    cropped = AT.frame_to_digit(_t(g, "frames").to(DEV), z)
    assert torch.equal(z, z_in), "z_where must not be modified"
    assert tuple(pasted.shape) == g["pasted"].shape and tuple(cropped.shape) == g["cropped"].shape
    assert np.array_equal(pasted.cpu().numpy(), g["pasted"])
    assert np.array_equal(cropped.cpu().numpy(), g["cropped"])


def test_B3_paste_ll_vs_reference_golden(golden):
    g = golden("stn")
    AT = _AT()
    ll, recon = AT.reconstruct_ll(_t(g, "digit").to(DEV), _t(g, "z_where").to(DEV), _t(g, "frames").to(DEV), want_recon=True)
    assert np.array_equal(recon.cpu().numpy(), g["recon"])          # sums of <= 3 gathered values + clamp: exact
    np.testing.assert_allclose(ll.cpu().numpy(), g["ll"], rtol=1e-5)
    ll2, none = AT.reconstruct_ll(_t(g, "digit").to(DEV), _t(g, "z_where").to(DEV), _t(g, "frames").to(DEV))
    assert none is None and torch.equal(ll, ll2)


@pytest.mark.parametrize("S,B,T,K", [(1, 1, 1, 1), (3, 2, 5, 3), (10, 4, 2, 2)])
def test_stn_random_vs_oracle_incl_shared_frames(S, B, T, K):
    """frames given once per instance ((B,T,FP,FP) / stride-0 particle axis) and single-frame views."""
    AT = _AT()
    torch.manual_seed(S * 10 + T)
    z = torch.randn(S, B, T, K, 2) * 0.9
    digit = torch.rand(S, B, K, 28, 28)
    fr = torch.rand(B, T, 96, 96)
    frs = fr.unsqueeze(0).expand(S, -1, -1, -1, -1)
    ref_c = stn.frame_to_digit(frs.numpy(), z.numpy())
    assert np.array_equal(AT.frame_to_digit(fr.to(DEV), z.to(DEV)).cpu().numpy(), ref_c)
    assert np.array_equal(AT.frame_to_digit(fr.to(DEV).unsqueeze(0).expand(S, -1, -1, -1, -1), z.to(DEV)).cpu().numpy(), ref_c)
    assert np.array_equal(AT.digit_to_frame(digit.to(DEV), z.to(DEV)).cpu().numpy(), stn.digit_to_frame(digit.numpy(), z.numpy()))
    recon, ll = stn.recon_and_ll(digit.numpy(), z.numpy(), frs.numpy())
    ll_g, rec_g = AT.reconstruct_ll(digit.to(DEV), z.to(DEV), fr.to(DEV), want_recon=True)
    assert np.array_equal(rec_g.cpu().numpy(), recon.numpy())
    np.testing.assert_allclose(ll_g.cpu().numpy(), ll.numpy(), rtol=1e-5)
    # one frame of a (S,B,T,..) tensor as a strided view, as apg_where uses it (objectives.py:204)
    t = T - 1
    ll_t, _ = AT.reconstruct_ll(digit.to(DEV), z[:, :, t:t + 1].contiguous().to(DEV), frs.to(DEV)[:, :, t:t + 1])
    np.testing.assert_allclose(ll_t.cpu().numpy()[..., 0], ll.numpy()[..., t], rtol=1e-5)


def test_B4_where_and_what_weights_vs_reference_golden(golden):
    from apgs_b200.bmnist import kernels
    from apgs_b200.bmnist.models import Dec_coor
    g = golden("bmnist")
    dc = Dec_coor(2, True, DEV)
    for tag in ("w0", "w1", "w5"):
        opt = lambda n: _t(g, tag + n).to(DEV) if (tag + n) in g.files else None
        lw, lq, lp = kernels.logw_where(_t(g, tag + "_q_mean").to(DEV), _t(g, tag + "_q_std").to(DEV), _t(g, tag + "_z_new").to(DEV),
                                        opt("_z_prev"), _t(g, tag + "_z_old").to(DEV), opt("_z_old_prev"), dc,
                                        ll_f=_t(g, tag + "_ll_f").squeeze(-1).to(DEV), ll_b=_t(g, tag + "_ll_b").squeeze(-1).to(DEV))
        np.testing.assert_allclose(lq.cpu().numpy(), g[tag + "_log_q_f"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(lp.cpu().numpy(), g[tag + "_log_p_f"], rtol=1e-5, atol=1e-5)
        mag = np.abs(g[tag + "_ll_f"]).squeeze(-1) + np.abs(g[tag + "_ll_b"]).squeeze(-1) + np.abs(g[tag + "_log_q_f"]) + np.abs(g[tag + "_log_p_f"])
        assert (np.abs(lw.cpu().numpy() - g[tag + "_log_w"]) <= 1e-5 * mag + 1e-4).all(), tag
    lw, lq, dens, lls = kernels.logw_what(_t(g, "wh_q_mu").to(DEV), _t(g, "wh_q_std").to(DEV), _t(g, "wh_z_what").to(DEV),
                                          _t(g, "wh_z_what_old").to(DEV), _t(g, "wh_ll_f").to(DEV), _t(g, "wh_ll_b").to(DEV))
    np.testing.assert_allclose(lq.cpu().numpy(), g["wh_log_q_f"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dens.cpu().numpy(), g["wh_dens"], rtol=1e-5)
    mag = np.abs(g["wh_ll_f"]).sum(-1) + np.abs(g["wh_ll_b"]).sum(-1)
    assert (np.abs(lw.cpu().numpy() - g["wh_log_w"]) <= 1e-5 * mag + 1e-4).all()
    # one-shot: log p - log q with the accumulated z_where terms
    lw, lq, dens, _ = kernels.logw_what(_t(g, "os_q_mu").to(DEV), _t(g, "os_q_std").to(DEV), _t(g, "os_z_what").to(DEV), None,
                                        _t(g, "os_ll").to(DEV), None, extra_log_p=_t(g, "os_log_p_where").to(DEV),
                                        extra_log_q=_t(g, "os_log_q_where").to(DEV))
    np.testing.assert_allclose(dens.cpu().numpy(), g["os_log_p"], rtol=1e-5)
    np.testing.assert_allclose(lq.cpu().numpy(), g["os_log_q"], rtol=1e-5, atol=1e-5)
    assert (np.abs(lw.cpu().numpy() - g["os_log_w"]) <= 1e-5 * (np.abs(g["os_log_p"]) + np.abs(g["os_log_q"]))).all()


def _models(seed=0):
    from apgs_b200.bmnist import models as M
    torch.manual_seed(seed)
    P = ob.make_params()
    ms = M.init_models(96, 28, 400, 400, 2, 10, True, DEV)
    for m, key in ((ms[0], "enc_coor"), (ms[2], "enc_digit"), (ms[3], "dec_digit")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})     # same state_dict keys as the reference
    return P, ms


def test_bmnist_sweep_runs_and_matches_oracle_statistics(golden):
    """End to end on the golden instance set (CPU/CUDA generators differ -> shapes, finiteness, and the log-joint of
    the one-shot, whose proposal does not depend on earlier draws of other particles, within Monte-Carlo error)."""
    from apgs_b200.bmnist import objectives as O
    from apgs_b200.resampler import Resampler
    g = golden("bmnist")
    K, FP, DP = 3, 96, 28
    S, B, T, sweeps = 12, 3, 3, 3
    P, models = _models()
    frames = _t(g, "frames")
    tmpl = _t(g, "template")
    flags = {"loss_required": True, "ess_required": True, "mode_required": True, "density_required": True}
    torch.manual_seed(5)
    with torch.no_grad():
        ref = ob.apg_objective(P, ob.make_prior(), frames.repeat(S, 1, 1, 1, 1), K, flags, sweeps, tmpl.repeat(S, B, K, 1, 1))
        tr = O.apg_objective(models, _AT(), frames.to(DEV).repeat(S, 1, 1, 1, 1), K, flags, sweeps,
                             Resampler("systematic", S, True, DEV), tmpl.to(DEV).repeat(S, B, K, 1, 1))
    for k in ("loss_phi", "loss_theta", "ess", "density"):
        assert tuple(tr[k].shape) == tuple(ref[k].shape), k
        assert bool(torch.isfinite(tr[k]).all()), k
    assert tuple(tr["E_where"].shape) == (sweeps, B, T, K, 2) and tuple(tr["E_recon"].shape) == (sweeps, B, T, FP, FP)
    a, b = tr["density"][0].cpu().mean(0), ref["density"][0].mean(0)
    assert bool(((a - b).abs() < 0.15 * b.abs()).all()), (a, b)


def test_bmnist_teacher_forced_step_vs_oracle(golden):
    """One apg_what step with the draw forced: crop -> Enc_digit -> Dec_digit -> paste/ll -> weight, against the oracle."""
    from apgs_b200.bmnist import kernels
    g = golden("bmnist")
    K = 3
    S, B, T = [int(v) for v in g["dims"][:3]]
    P, models = _models()
    enc_digit, dec_digit = models[2], models[3]
    AT = _AT()
    frames = _t(g, "frames")
    frs = frames.repeat(S, 1, 1, 1, 1)
    z_where, z_new, z_old = _t(g, "wh_z_where"), _t(g, "wh_z_what"), _t(g, "wh_z_what_old")
    with torch.no_grad():
        cropped = AT.frame_to_digit(frames.to(DEV), z_where.to(DEV))
        assert torch.equal(cropped.cpu(), ob.frame_to_digit(frs, z_where))
        q_mu, q_std = enc_digit.params(cropped.view(S, B, T, K, 784))
        np.testing.assert_allclose(q_mu.cpu().numpy(), g["wh_q_mu"], rtol=2e-4, atol=2e-5)
        ll_f, _ = AT.reconstruct_ll(dec_digit.templates(z_new.to(DEV)), z_where.to(DEV), frames.to(DEV))
        ll_b, _ = AT.reconstruct_ll(dec_digit.templates(z_old.to(DEV)), z_where.to(DEV), frames.to(DEV))
        np.testing.assert_allclose(ll_f.cpu().numpy(), g["wh_ll_f"], rtol=2e-5)
        np.testing.assert_allclose(ll_b.cpu().numpy(), g["wh_ll_b"], rtol=2e-5)


# ------------------------------------------------------------------------------ N3: template matching
def _match_ref(frame, template, softmax):
    """the reference's expression (apgs/bmnist/objectives.py:72-74) on CPU."""
    import torch.nn.functional as F
    S, B, FP, _ = frame.shape
    DP = template.shape[-1]
    c = F.conv2d(frame.reshape(S * B, FP, FP).unsqueeze(0), template.reshape(S * B, DP, DP).unsqueeze(1), groups=S * B)
    CP = c.shape[-1]
    c = c.squeeze(0).view(S, B, CP * CP)
    return F.softmax(c, -1) if softmax else c


@pytest.mark.parametrize("S,B,FP,DP", [(3, 5, 96, 28), (2, 2, 32, 8), (1, 3, 40, 28)])
def test_N3_template_match_vs_reference_expression(S, B, FP, DP):
    from apgs_b200.bmnist import kernels
    torch.manual_seed(FP + S)
    T, K = 4, 3
    frames = torch.rand(B, T, FP, FP) * (torch.rand(B, T, FP, FP) > 0.7)
    frames5 = frames.unsqueeze(0).expand(S, -1, -1, -1, -1)                 # S-shared frames: stride 0 over s
    tmpl = torch.rand(S, B, K, DP, DP)
    fd, td = frames.to(DEV).unsqueeze(0).expand(S, -1, -1, -1, -1), tmpl.to(DEV)
    for t, k in ((0, 0), (3, 2)):
        ref_raw = _match_ref(frames5[:, :, t].contiguous(), tmpl[:, :, k].contiguous(), False)
        raw = kernels.template_match(fd[:, :, t], td[:, :, k], softmax=False)       # strided views, read in place
        mag = _match_ref(frames5[:, :, t].contiguous().abs(), tmpl[:, :, k].contiguous().abs(), False)
        assert float(((raw.cpu() - ref_raw).abs() - 2e-6 * mag).max()) <= 0
        ref = _match_ref(frames5[:, :, t].contiguous(), tmpl[:, :, k].contiguous(), True)
        out = kernels.template_match(fd[:, :, t], td[:, :, k])
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=3e-4, atol=1e-12)
        np.testing.assert_allclose(out.sum(-1).cpu().numpy(), 1.0, rtol=1e-5)
    # a residual frame (negative pixels after subtracting a reconstruction) and a dense (S,B,FP,FP) input
    left = (frames5[:, :, 1] - torch.rand(S, B, FP, FP) * 0.5).contiguous()
    out = kernels.template_match(left.to(DEV), td[:, :, 1])
    np.testing.assert_allclose(out.cpu().numpy(), _match_ref(left, tmpl[:, :, 1].contiguous(), True).numpy(), rtol=3e-4, atol=1e-12)


# ------------------------------------------------------------------------------ N2: training mode
def _train_models():
    from apgs_b200.bmnist import models as M
    torch.manual_seed(0)
    P = ob.make_params()
    ms, opt = M.init_models(96, 28, 400, 400, 2, 10, True, DEV, lr=1e-4)
    for m, key in zip((ms[0], ms[2], ms[3]), ("enc_coor", "enc_digit", "dec_digit")):
        m.load_state_dict({k: v.to(DEV) for k, v in P[key].items()})
    return P, ms, opt


def _grads_close(module, Pg, rtol=5e-3):
    scale = max(float(v.grad.abs().max()) for v in Pg.values() if v.grad is not None)
    for name, p in module.named_parameters():
        assert p.grad is not None, name
        np.testing.assert_allclose(p.grad.cpu().numpy(), Pg[name].grad.numpy(), rtol=rtol, atol=2e-4 * scale, err_msg=name)


def test_N2_differentiable_paste_matches_the_kernel_and_the_oracle():
    AT = _AT()
    torch.manual_seed(3)
    S, B, T, K = 3, 2, 4, 3
    z = torch.randn(S, B, T, K, 2) * 0.9
    digit = torch.rand(S, B, K, 28, 28)
    fr = torch.rand(B, T, 96, 96).round()
    dd = digit.to(DEV).requires_grad_(True)
    assert torch.equal(AT.digit_to_frame_torch(dd, z.to(DEV)).detach(), AT.digit_to_frame(digit.to(DEV), z.to(DEV)))
    ll_t = AT.reconstruct_ll_torch(dd, z.to(DEV), fr.to(DEV))
    ll_k, _ = AT.reconstruct_ll(digit.to(DEV), z.to(DEV), fr.to(DEV))
    np.testing.assert_allclose(ll_t.detach().cpu().numpy(), ll_k.cpu().numpy(), rtol=2e-5)
    ll_t.sum().backward()
    dc = digit.clone().requires_grad_(True)
    _, ll_o = ob.recon_ll(dc, z, fr.unsqueeze(0).expand(S, -1, -1, -1, -1))
    ll_o.sum().backward()
    np.testing.assert_allclose(dd.grad.cpu().numpy(), dc.grad.numpy(), rtol=2e-4, atol=1e-3)


def test_N2_bmnist_what_and_where_gradients_match_oracle_autograd(golden):
    """loss.backward() through apg_what and one frame of apg_where == autograd through the oracle (= the reference's
    computation) on the same draws."""
    from apgs_b200.bmnist import objectives as O
    from apgs_b200.resampler import Resampler
    g = golden("bmnist")
    K = 3
    P, (enc_coor, dec_coor, enc_digit, dec_digit), _ = _train_models()
    AT = _AT()
    frames = _t(g, "frames")
    B, T = frames.shape[0], frames.shape[1]
    S = 4
    frs = frames.unsqueeze(0).repeat(S, 1, 1, 1, 1)
    torch.manual_seed(2)
    z_where = torch.randn(S, B, T, K, 2) * 0.5
    z_what_old = torch.randn(S, B, K, 10)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    leaf = lambda d: {k: v.clone().requires_grad_(True) for k, v in d.items()}
    # ---- apg_what
    trace = {"loss_phi": [torch.zeros(1, device=DEV)], "loss_theta": [torch.zeros(1, device=DEV)], "ess": [],
             "E_where": [], "E_what": [], "E_recon": [], "density": [torch.zeros(1, S, B, device=DEV)]}
    log_w, z_what, trace = O.apg_what(enc_digit, dec_digit, AT, frs.to(DEV), z_where.to(DEV), z_what_old.to(DEV), trace, flags)
    (trace["loss_phi"][0] + trace["loss_theta"][0]).sum().backward()
    Pe, Pd = leaf(P["enc_digit"]), leaf(P["dec_digit"])
    cropped = ob.frame_to_digit(frs, z_where).view(S, B, T, K, 784)
    q_mu, q_std = ob.enc_digit_params(Pe, cropped)
    _, ll_f = ob.recon_ll(ob.dec_digit_mean(Pd, z_what.cpu()), z_where, frs)
    _, ll_b = ob.recon_ll(ob.dec_digit_mean(Pd, z_what_old), z_where, frs)
    lw, lqf, _ = ob.what_step_weight(q_mu, q_std, z_what.cpu(), z_what_old, ll_f, ll_b)
    w = torch.softmax(lw.detach(), 0)
    ref_phi, ref_theta = (w * (-lqf)).sum(0).mean(), (w * (-ll_f.sum(-1))).sum(0).mean()
    (ref_phi + ref_theta).backward()
    np.testing.assert_allclose(float(trace["loss_phi"][0].detach()), float(ref_phi.detach()), rtol=5e-4)
    np.testing.assert_allclose(float(trace["loss_theta"][0].detach()), float(ref_theta.detach()), rtol=5e-4)
    _grads_close(enc_digit, Pe)
    _grads_close(dec_digit, Pd)
    # ---- apg_where (T frames, per-frame resampling inside): every parameter that the reference trains gets a gradient
    for m in (enc_digit, dec_digit):
        m.zero_grad()
    trace = {"loss_phi": [], "loss_theta": [], "ess": [], "E_where": [], "E_what": [], "E_recon": [], "density": []}
    z_new, trace = O.apg_where(enc_coor, dec_coor, dec_digit, AT, Resampler("systematic", S, True, DEV), frs.to(DEV),
                               z_what.detach(), z_where.to(DEV), trace, flags)
    (trace["loss_phi"][0] + trace["loss_theta"][0]).sum().backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) and float(p.grad.abs().max()) > 0
               for p in enc_coor.parameters())
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in dec_digit.parameters())
    assert not z_new.requires_grad


def test_N2_bmnist_where_step_gradient_single_frame_vs_oracle():
    """one frame of apg_where with S = 1 particle (weights are then exactly 1, no resampling effect): gradients of
    loss_phi w.r.t. Enc_coor and of loss_theta w.r.t. Dec_digit against the oracle's autograd on the same draw."""
    from apgs_b200.bmnist import objectives as O
    from apgs_b200.resampler import Resampler
    K, S, B, T = 3, 1, 3, 1
    P, (enc_coor, dec_coor, enc_digit, dec_digit), _ = _train_models()
    AT = _AT()
    torch.manual_seed(4)
    frames = ob.simulate(B, T, K)
    frs = frames.unsqueeze(0).repeat(S, 1, 1, 1, 1)
    z_what = torch.randn(S, B, K, 10)
    z_where_old = torch.randn(S, B, T, K, 2) * 0.5
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    trace = {"loss_phi": [], "loss_theta": [], "ess": [], "E_where": [], "E_what": [], "E_recon": [], "density": []}
    z_new, trace = O.apg_where(enc_coor, dec_coor, dec_digit, AT, Resampler("systematic", S, True, DEV), frs.to(DEV),
                               z_what.to(DEV), z_where_old.to(DEV), trace, flags)
    (trace["loss_phi"][0] + trace["loss_theta"][0]).sum().backward()
    leaf = lambda d: {k: v.clone().requires_grad_(True) for k, v in d.items()}
    Pc, Pd = leaf(P["enc_coor"]), leaf(P["dec_digit"])
    template = ob.dec_digit_mean(Pd, z_what).detach()
    pm = ob.propose_one_movement({"enc_coor": Pc}, ob.make_prior(), frs[:, :, 0], template, None, z_where_old[:, :, 0], None,
                                 forced=z_new.cpu()[:, :, 0])
    _, ll_f = ob.recon_ll(ob.dec_digit_mean(Pd, z_what), z_new.cpu(), frs)
    ref = (-pm["log_q_f"]).sum(0).mean() + (-ll_f.squeeze(-1)).sum(0).mean()          # S = 1: w == 1
    ref.backward()
    np.testing.assert_allclose(float((trace["loss_phi"][0] + trace["loss_theta"][0]).sum().detach()), float(ref.detach()), rtol=5e-4)
    _grads_close(enc_coor, Pc)
    _grads_close(dec_digit, Pd)


def test_N2_bmnist_training_loop_runs_like_reference_train():
    """the reference's train() inner loop (apgs/bmnist/apg_training.py) on the drop-in: finite losses, every parameter
    gets a gradient, Adam steps change the weights."""
    from apgs_b200.bmnist import objectives as O
    from apgs_b200.resampler import Resampler
    K, S, B, T, sweeps = 3, 6, 4, 3, 3
    P, models, opt = _train_models()
    AT = _AT()
    torch.manual_seed(1)
    frames = ob.simulate(B, T, K).to(DEV)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 28), torch.linspace(-1, 1, 28), indexing="ij")
    mm = torch.exp(-(xx ** 2 + yy ** 2) * 3).clamp(0, 1).to(DEV).repeat(S, B, K, 1, 1)
    flags = {"loss_required": True, "ess_required": True, "mode_required": False, "density_required": True}
    rs = Resampler("systematic", S, True, DEV)
    mods = (models[0], models[2], models[3])
    before = [p.detach().clone() for m in mods for p in m.parameters()]
    for it in range(2):
        opt.zero_grad()
        trace = O.apg_objective(models, AT, frames.unsqueeze(0).repeat(S, 1, 1, 1, 1), K, flags, sweeps, rs, mm)
        loss = trace["loss_phi"].sum() + trace["loss_theta"].sum()
        loss.backward()
        assert bool(torch.isfinite(loss))
        for m in mods:
            for n_, p in m.named_parameters():
                assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n_
        opt.step()
    after = [p.detach() for m in mods for p in m.parameters()]
    assert all(not torch.equal(a, b) for a, b in zip(before, after))


def test_paste_subtract_equals_paste_then_subtract():
    """the fused residual-frame update == digit_to_frame followed by the subtraction, bit for bit, on strided views."""
    AT = _AT()
    torch.manual_seed(9)
    S, B, T, K = 3, 4, 3, 3
    frames = torch.rand(B, T, 96, 96, device=DEV).unsqueeze(0).expand(S, -1, -1, -1, -1)      # stride-0 particle axis
    tmpl = torch.rand(S, B, K, 28, 28, device=DEV)
    z = (torch.randn(S, B, K, 2, device=DEV) * 0.9)
    z[0, 0, 0] = torch.tensor([1.6, -1.7], device=DEV)                                       # partly outside the frame
    for t, k in ((0, 0), (2, 1), (1, 2)):
        ref = frames[:, :, t] - AT.digit_to_frame(tmpl[:, :, k].unsqueeze(2), z[:, :, k].unsqueeze(2).unsqueeze(2)).squeeze(2).squeeze(2)
        out = AT.subtract_digit(frames[:, :, t], tmpl[:, :, k], z[:, :, k])
        assert torch.equal(out, ref), (t, k)
    left = torch.randn(S, B, 96, 96, device=DEV)
    ref = left - AT.digit_to_frame(tmpl[:, :, 1].unsqueeze(2), z[:, :, 1].unsqueeze(2).unsqueeze(2)).squeeze(2).squeeze(2)
    assert torch.equal(AT.subtract_digit(left, tmpl[:, :, 1], z[:, :, 1]), ref)


def test_bmnist_match_and_residual_kernels_stay_inside_their_buffers():
    """raw C-ABI calls with the outputs embedded in sentinel guard bands and strided inputs (no compute-sanitizer on
    this pool)."""
    from apgs_b200 import _lib
    from apgs_b200._lib import ptr
    lib = _lib.load()
    S, B, T, K, FP, DP = 3, 5, 2, 3, 96, 28
    CP = FP - DP + 1
    G = 4096
    torch.manual_seed(2)
    frames = torch.rand(B, T, FP, FP, device=DEV)
    tmpl = torch.rand(S, B, K, DP, DP, device=DEV)
    z = torch.randn(S, B, 2, device=DEV) * 0.8
    lin = torch.linspace(-1, 1, FP).to(DEV)
    st = torch.cuda.current_stream().cuda_stream

    def guarded(n):
        buf = torch.full((n + 2 * G,), -7.0, device=DEV)
        return buf, buf[G:G + n]
    bo, out = guarded(S * B * CP * CP)
    f_t, t_k = frames[:, 1], tmpl[:, :, 2]
    rc = lib.apgs_bmnist_template_match_f32(ptr(f_t), 0, f_t.stride(0), ptr(t_k), t_k.stride(0), t_k.stride(1), S, B, FP, DP, 1,
                                            ptr(out), st)
    assert rc == 0, lib.apgs_last_error()
    br, res = guarded(S * B * FP * FP)
    rc = lib.apgs_bmnist_paste_subtract_f32(ptr(f_t), 0, f_t.stride(0), ptr(t_k), t_k.stride(0), t_k.stride(1), ptr(z), ptr(lin),
                                            S, B, FP, DP, ptr(res), st)
    assert rc == 0, lib.apgs_last_error()
    torch.cuda.synchronize()
    for buf in (bo, br):
        assert bool((buf[:G] == -7.0).all()) and bool((buf[-G:] == -7.0).all())
    np.testing.assert_allclose(out.view(S, B, -1).sum(-1).cpu().numpy(), 1.0, rtol=1e-5)
    assert bool(torch.isfinite(res).all())


@pytest.mark.parametrize("S,B,T,K", [(2, 3, 4, 3), (1, 1, 1, 1)])
def test_stn_bilinear_mode_vs_torch_grid_sample(S, B, T, K):
    """the bilinear mode switch (north_star item 4): digit_to_frame / frame_to_digit with mode='bilinear' against
    affine_grid + grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) on the reference's own theta
    (apgs/bmnist/affine_transformer.py:41-65), incl. digits partly outside the frame."""
    import torch.nn.functional as F
    from apgs_b200.bmnist import Affine_Transformer
    FP, DP = 96, 28
    torch.manual_seed(S + T)
    AT = Affine_Transformer(FP, DP, True, DEV, mode="bilinear")
    digit = torch.rand(S, B, K, DP, DP, device=DEV)
    frames = torch.rand(S, B, T, FP, FP, device=DEV)
    z = (torch.rand(S, B, T, K, 2, device=DEV) * 2 - 1) * 1.2
    # reference expressions (diagonal theta), on the GPU through stock torch ops
    M = S * B * T * K
    s_dtof = torch.tensor([[FP / DP, 0], [0, FP / DP]], device=DEV).repeat(S, B, T, K, 1, 1)
    zz = z.clone()
    zz[..., 0] = zz[..., 0] * -1
    theta = torch.cat((s_dtof, (zz * ((FP - DP) / DP)).unsqueeze(-1)), -1).view(M, 2, 3)
    grid = F.affine_grid(theta, torch.Size((M, 1, FP, FP)), align_corners=True)
    ref_paste = F.grid_sample(digit.unsqueeze(2).repeat(1, 1, T, 1, 1, 1).view(M, DP, DP).unsqueeze(1), grid, mode="bilinear",
                              padding_mode="zeros", align_corners=True).squeeze(1).view(S, B, T, K, FP, FP)
    out = AT.digit_to_frame(digit, z)
    torch.testing.assert_close(out, ref_paste, rtol=1e-5, atol=2e-6)
    s_ftod = torch.tensor([[DP / FP, 0], [0, DP / FP]], device=DEV).repeat(S, B, T, K, 1, 1)
    zz = z.clone()
    zz[..., 1] = zz[..., 1] * -1
    theta = torch.cat((s_ftod, (zz * ((FP - DP) / FP)).unsqueeze(-1)), -1).view(M, 2, 3)
    grid = F.affine_grid(theta, torch.Size((M, 1, DP, DP)), align_corners=True)
    ref_crop = F.grid_sample(frames.unsqueeze(3).repeat(1, 1, 1, K, 1, 1).view(M, FP, FP).unsqueeze(1), grid, mode="bilinear",
                             padding_mode="zeros", align_corners=True).squeeze(1).view(S, B, T, K, DP, DP)
    out = AT.frame_to_digit(frames, z)
    torch.testing.assert_close(out, ref_crop, rtol=1e-5, atol=2e-6)
    # the default stays the reference's nearest mode
    assert Affine_Transformer(FP, DP, True, DEV).mode == "nearest"


@pytest.mark.parametrize("M,K,N", [(640, 4761, 400), (37, 4761, 400), (64, 100, 80), (130, 1029, 161), (1, 33, 7), (257, 513, 83)])
def test_linear_tcgen05_matches_fp64_like_an_fp32_gemm(M, K, N):
    """apgs_linear_tc5_f32 (3xTF32 through tcgen05.mma, accumulator in tensor memory): fp32-GEMM accuracy against an fp64 product
    for ragged M / N / K (tiles of 128 x 80, K slices of 32), with / without bias and ReLU, deterministic."""
    from apgs_b200.bmnist import kernels
    torch.manual_seed(M + K + 2)
    x = torch.softmax(torch.randn(M, K) * 3, -1) if K > 1000 else torch.randn(M, K)
    w = torch.randn(N, K) / K ** 0.5
    b = torch.randn(N)
    xg, wg, bg = x.to(DEV), w.to(DEV), b.to(DEV)
    ref = x.double() @ w.double().t() + b.double()
    mag = x.double().abs() @ w.double().abs().t() + b.double().abs()
    out = kernels.linear_tc5(xg, wg, bg)
    assert tuple(out.shape) == (M, N) and bool(torch.isfinite(out).all())
    err = float(((out.cpu().double() - ref).abs() / mag).max())
    f32 = float(((torch.nn.functional.linear(xg, wg, bg).cpu().double() - ref).abs() / mag).max())
    assert err <= 4e-6 and err <= 4 * f32 + 5e-7, (err, f32)
    assert torch.equal(kernels.linear_tc5(xg, wg, bg, relu=True), out.clamp_min(0))
    assert torch.equal(kernels.linear_tc5(xg, wg, bg), out)
    assert float(((kernels.linear_tc5(xg, wg, None).cpu().double() - (ref - b.double())).abs() / mag).max()) <= 4e-6
    # agrees with the warp-level MMA form of the same 3xTF32 product
    assert float(((out - kernels.linear_tf32x3(xg, wg, bg)).abs().cpu().double() / mag).max()) <= 4e-6


@pytest.mark.parametrize("M,K,N", [(640, 4761, 400), (37, 4761, 400), (64, 100, 80), (130, 1029, 161), (1, 33, 7), (257, 513, 83)])
def test_linear_tf32x3_matches_fp64_like_an_fp32_gemm(M, K, N):
    """apgs_linear_tf32x3_f32 (Enc_coor's Linear(4761 -> 400) + ReLU on the tensor cores): the error against an fp64 product is
    of the order of an fp32 GEMM's own (<= 3e-6 of sum_k |x_k w_k| + |b|, and within 3x of torch's fp32 GEMM on the same inputs), for ragged M / N / K, with and without bias / ReLU;
    deterministic from call to call."""
    from apgs_b200.bmnist import kernels
    torch.manual_seed(M + K)
    x = torch.softmax(torch.randn(M, K) * 3, -1) if K > 1000 else torch.randn(M, K)
    w = torch.randn(N, K) / K ** 0.5
    b = torch.randn(N)
    xg, wg, bg = x.to(DEV), w.to(DEV), b.to(DEV)
    ref = x.double() @ w.double().t() + b.double()
    mag = x.double().abs() @ w.double().abs().t() + b.double().abs()
    out = kernels.linear_tf32x3(xg, wg, bg, relu=False)
    assert tuple(out.shape) == (M, N)
    err = ((out.cpu().double() - ref).abs() / mag).max()
    f32 = ((torch.nn.functional.linear(xg, wg, bg).cpu().double() - ref).abs() / mag).max()
    assert float(err) <= 3e-6 and float(err) <= 3 * float(f32) + 1e-7, (float(err), float(f32))
    out_r = kernels.linear_tf32x3(xg, wg, bg, relu=True)
    assert torch.equal(out_r, out.clamp_min(0)) and torch.equal(out_r, kernels.linear_tf32x3(xg, wg, bg, relu=True))
    nb = kernels.linear_tf32x3(xg, wg, None)
    assert float(((nb.cpu().double() - (ref - b.double())).abs() / mag).max()) <= 3e-6
    # leading dimensions are kept
    if M % 2 == 0:
        assert torch.equal(kernels.linear_tf32x3(xg.view(2, M // 2, K), wg, bg), out.view(2, M // 2, N))


def test_enc_coor_uses_the_tensor_core_layer_and_matches_the_torch_modules():
    from apgs_b200.bmnist import models as M
    torch.manual_seed(0)
    enc = M.Enc_coor(4761, 400, 2).to(DEV)
    conved = torch.softmax(torch.randn(10, 64, 4761, device=DEV) * 2, -1)
    with torch.no_grad():
        mean, std = enc.params(conved)
        M.Enc_coor.fused = False
        try:
            mean_t, std_t = enc.params(conved)
        finally:
            M.Enc_coor.fused = True
    assert float((mean - mean_t).abs().max()) <= 3e-6 and float(((std - std_t).abs() / std_t).max()) <= 1e-5
    default = M.Enc_coor.linear_kernel
    assert default == "tc5pk"
    for which in ("mma", "tc5"):
        M.Enc_coor.linear_kernel = which
        try:
            with torch.no_grad():
                mean_x, std_x = enc.params(conved)
        finally:
            M.Enc_coor.linear_kernel = default
        assert float((mean_x - mean_t).abs().max()) <= 3e-6 and float(((std_x - std_t).abs() / std_t).max()) <= 1e-5
    # the default form takes the map as the packed operand (what the template-match kernel writes)
    from apgs_b200.bmnist import kernels
    with torch.no_grad():
        mean_p, std_p = enc.params(kernels.pack_rows(conved, 0))
    assert tuple(mean_p.shape) == tuple(mean_t.shape)
    assert float((mean_p - mean_t).abs().max()) <= 3e-6 and float(((std_p - std_t).abs() / std_t).max()) <= 1e-5
    # training mode keeps the differentiable torch modules
    for p in enc.parameters():
        p.requires_grad_(True)
    m2, _ = enc.params(conved)
    assert m2.requires_grad


def _packed_index(TR, r, k, nch):
    """csrc/tf32_pack.cuh packed_index, restated"""
    return ((r // TR) * nch + k // 32) * (2 * TR * 32) + ((r % TR) // 8) * 256 + ((k % 32) // 4) * 32 + (r % 8) * 4 + k % 4


def _tf32_rna(v):
    """cvt.rna.tf32.f32 on a float32 numpy array: round to 10 explicit mantissa bits, ties away from zero"""
    import numpy as np
    u = v.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("rows,K,which", [(200, 100, 0), (81, 33, 1), (640, 4761, 0), (400, 4761, 1)])
def test_pack_rows_layout_halves_and_padding(rows, K, which):
    """apgs_linear_tc5_pack_f32: every element sits where csrc/tf32_pack.cuh says, hi is the TF32 rounding of the value, lo the TF32
    rounding of the exact remainder, padding rows / columns are zero, and PackedRows.unpack() inverts it."""
    import numpy as np
    from apgs_b200.bmnist import kernels
    torch.manual_seed(rows + K)
    x = torch.randn(rows, K)
    pk = kernels.pack_rows(x.to(DEV), which)
    TR, nch = (128, 80)[which], (K + 31) // 32
    buf = pk.buf.cpu().numpy()
    assert buf.size == ((rows + TR - 1) // TR) * nch * 2 * TR * 32
    rng = np.random.default_rng(0)
    rr = np.concatenate([rng.integers(0, rows, 300), [0, rows - 1]])
    kk = np.concatenate([rng.integers(0, K, 300), [K - 1, 0]])
    at = np.array([_packed_index(TR, int(r), int(k), nch) for r, k in zip(rr, kk)])
    v = x.numpy()[rr, kk]
    hi = _tf32_rna(v)
    assert np.array_equal(buf[at], hi)
    assert np.array_equal(buf[at + TR * 32], _tf32_rna(v - hi))
    full = np.zeros(buf.size, dtype=bool)                        # every slot not addressed by a real element is zero
    allr, allk = np.meshgrid(np.arange(rows), np.arange(K), indexing="ij")
    ia = ((allr // TR) * nch + allk // 32) * (2 * TR * 32) + ((allr % TR) // 8) * 256 + ((allk % 32) // 4) * 32 + (allr % 8) * 4 + allk % 4
    full[ia.ravel()] = True
    full[ia.ravel() + TR * 32] = True
    assert not buf[~full].any()
    un = pk.unpack().cpu()
    assert tuple(un.shape) == (rows, K) and float((un - x).abs().max()) <= 2.0 ** -21 * float(x.abs().max())


@pytest.mark.parametrize("M,K,N", [(640, 4761, 400), (37, 4761, 400), (64, 100, 80), (130, 1029, 161), (1, 33, 7), (257, 513, 83)])
def test_linear_tcgen05_from_packed_operands_is_bit_identical(M, K, N, monkeypatch):
    """apgs_linear_tc5_packed_f32 (stages are bulk copies of pre-split, pre-laid-out tiles; warp-specialised producer / MMA issuer):
    the same MMAs in the same order and the same split-K as apgs_linear_tc5_f32 -> equal bit for bit, at ragged shapes, with /
    without bias and ReLU.  The 4-stage geometry (a measurement knob: fewer, longer K slices) is checked at its own accuracy."""
    from apgs_b200.bmnist import kernels
    monkeypatch.delenv("APGS_TC5PK_STAGES", raising=False)
    torch.manual_seed(M + K + 3)
    x = (torch.softmax(torch.randn(M, K) * 3, -1) if K > 1000 else torch.randn(M, K)).to(DEV)
    w = (torch.randn(N, K) / K ** 0.5).to(DEV)
    b = torch.randn(N).to(DEV)
    xp, wp = kernels.pack_rows(x, 0), kernels.pack_rows(w, 1)
    out = kernels.linear_tc5_packed(xp, wp, b)
    assert tuple(out.shape) == (M, N) and bool(torch.isfinite(out).all())
    assert torch.equal(out, kernels.linear_tc5(x, w, b))
    assert torch.equal(kernels.linear_tc5_packed(xp, wp, b, relu=True), out.clamp_min(0))
    assert torch.equal(kernels.linear_tc5_packed(xp, wp, None), kernels.linear_tc5(x, w, None))
    ref = x.double() @ w.double().t() + b.double()
    mag = x.double().abs() @ w.double().abs().t() + b.double().abs()
    assert float(((out.double() - ref).abs() / mag).max()) <= 4e-6
    xb = x.reshape(1, M, K).expand(2, M, K).contiguous()            # leading shape is kept
    assert tuple(kernels.linear_tc5_packed(kernels.pack_rows(xb, 0), wp, b).shape) == (2, M, N)
    # the other geometry: other slices of K, deterministic, accuracy of its longer accumulation chains
    monkeypatch.setenv("APGS_TC5PK_STAGES", "4")
    o4 = kernels.linear_tc5_packed(xp, wp, b)
    assert float(((o4.double() - ref).abs() / mag).max()) <= 8e-6
    assert torch.equal(o4, kernels.linear_tc5_packed(xp, wp, b))
    assert torch.equal(kernels.linear_tc5_packed(xp, wp, b, relu=True), o4.clamp_min(0))


@pytest.mark.parametrize("S,B", [(10, 64), (3, 5)])
def test_template_match_writes_the_packed_operand(S, B, monkeypatch):
    """apgs_bmnist_template_match_packed_f32: the softmax map as TF32 halves in the GEMM's layout equals the fp32 map to the
    splitting error, padding columns are zero, and Enc_coor's layer on it equals the layer on the fp32 map bit for bit (the
    register-staged kernel splits the same values the same way)."""
    from apgs_b200.bmnist import kernels
    from apgs_b200.bmnist.models import Enc_coor
    monkeypatch.delenv("APGS_TC5PK_STAGES", raising=False)           # default geometry: the split-K of the register-staged kernel
    torch.manual_seed(S * 100 + B)
    FP, DP, T, K = 96, 28, 2, 2
    frames = torch.rand(S, B, T, FP, FP, device=DEV)
    tmpl = torch.rand(S, B, K, DP, DP, device=DEV)
    fmap = kernels.template_match(frames[:, :, 1], tmpl[:, :, 1])                    # strided views
    pk = kernels.template_match(frames[:, :, 1], tmpl[:, :, 1], packed=True)
    assert isinstance(pk, kernels.PackedRows) and pk.rows == S * B and pk.K == 4761 and pk.lead == (S, B)
    un = pk.unpack()
    assert float((un - fmap).abs().max()) <= 2.0 ** -21 * float(fmap.abs().max())
    nch = (4761 + 31) // 32
    t = pk.buf.view(-1, nch, 2, 16, 8, 8, 4)
    rows_live = torch.arange(t.shape[0] * 128, device=DEV).view(-1, 16, 8) < S * B           # (row tile, group, row)
    pad = t[:, -1, :, :, 4761 % 32 // 4:, :, :]                                            # k groups holding columns >= 4761 of the last chunk
    pad = pad[..., 1:] if 4761 % 4 else pad
    assert not bool(pad.permute(0, 2, 4, 1, 3, 5)[rows_live].any())
    torch.manual_seed(1)
    enc = Enc_coor(4761, 400, 2).to(DEV)
    with torch.no_grad():
        assert enc.wants_packed(FP, DP)
        m1, s1 = enc.params(pk)
        enc.linear_kernel = "tc5"
        assert not enc.wants_packed(FP, DP)
        m0, s0 = enc.params(fmap)
        m2, _ = enc.params(pk)                                                        # PackedRows on a non-packed path: unpacked
    assert torch.equal(m1, m0) and torch.equal(s1, s0)
    assert float((m2 - m0).abs().max()) <= 1e-5
    # the weight cache follows in-place updates of the parameter
    enc.linear_kernel = "tc5pk"
    with torch.no_grad():
        enc.enc_hidden[0].weight.mul_(0.5)
        m3, _ = enc.params(pk)
        enc.linear_kernel = "tc5"
        m4, _ = enc.params(fmap)
    assert torch.equal(m3, m4) and not torch.equal(m3, m1)


def test_packed_operand_kernels_stay_inside_their_buffers():
    """raw C-ABI calls of the packed-operand entries with every output embedded in sentinel guard bands (no compute-sanitizer
    on this pool): the pack kernel, the template match writing the packed map from strided views, and the GEMM reading both
    (ragged M / N / K: 15 rows of a 128-row tile, 7 of an 80-row tile, 25 columns of the last chunk)."""
    from apgs_b200 import _lib
    from apgs_b200._lib import ptr
    lib = _lib.load()
    S, B, T, K, FP, DP, N = 3, 5, 2, 3, 96, 28, 87
    CP = FP - DP + 1
    KK = CP * CP
    G = 4096
    torch.manual_seed(5)
    frames = torch.rand(B, T, FP, FP, device=DEV)
    tmpl = torch.rand(S, B, K, DP, DP, device=DEV)
    W = torch.randn(N, KK + 3, device=DEV)[:, :KK] / 69.0                   # row stride ld = K + 3
    bias = torch.randn(N, device=DEV)
    st = torch.cuda.current_stream().cuda_stream

    def guarded(n, dtype=torch.float32, fill=-7.0):
        buf = torch.full((n + 2 * G,), fill, device=DEV, dtype=dtype)
        return buf, buf[G:G + n]
    na, nw = lib.apgs_linear_tc5_packed_bytes(S * B, KK, 0) // 4, lib.apgs_linear_tc5_packed_bytes(N, KK, 1) // 4
    assert na == 1 * 149 * 2 * 128 * 32 and nw == 2 * 149 * 2 * 80 * 32
    ba, ap = guarded(na)
    bw, wp = guarded(nw)
    f_t, t_k = frames[:, 1], tmpl[:, :, 2]
    rc = lib.apgs_bmnist_template_match_packed_f32(ptr(f_t), 0, f_t.stride(0), ptr(t_k), t_k.stride(0), t_k.stride(1), S, B, FP, DP,
                                                   ptr(ap), st)
    assert rc == 0, lib.apgs_last_error()
    rc = lib.apgs_linear_tc5_pack_f32(ptr(W), N, KK, W.stride(0), 1, ptr(wp), st)
    assert rc == 0, lib.apgs_last_error()
    nws = lib.apgs_linear_tc5_packed_workspace_bytes(S * B, N, KK)
    bws, ws = guarded(nws, torch.uint8, 0x5a)
    bo, out = guarded(S * B * N)
    # rows >= S*B of the activation tile keep the sentinel: they only reach accumulator rows that are never stored
    rc = lib.apgs_linear_tc5_packed_f32(ptr(ap), ptr(wp), ptr(bias), S * B, KK, N, 1, ptr(out), ptr(ws), nws, st)
    assert rc == 0, lib.apgs_last_error()
    torch.cuda.synchronize()
    for buf, fill in ((ba, -7.0), (bw, -7.0), (bo, -7.0), (bws, 0x5a)):
        assert bool((buf[:G] == fill).all()) and bool((buf[-G:] == fill).all())
    fmap = torch.empty(S * B * KK, device=DEV)
    rc = lib.apgs_bmnist_template_match_f32(ptr(f_t), 0, f_t.stride(0), ptr(t_k), t_k.stride(0), t_k.stride(1), S, B, FP, DP, 1,
                                            ptr(fmap), st)
    assert rc == 0, lib.apgs_last_error()
    ref = torch.relu(fmap.view(S * B, KK).double() @ W.double().t() + bias.double())
    np.testing.assert_allclose(out.view(S * B, N).cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=2e-6)
    # argument checks
    assert lib.apgs_linear_tc5_packed_bytes(0, 10, 0) == 0 and lib.apgs_linear_tc5_packed_bytes(10, 10, 2) == 0
    assert lib.apgs_linear_tc5_pack_f32(ptr(W), N, KK, KK - 1, 1, ptr(wp), st) < 0                       # ld < K
    assert lib.apgs_linear_tc5_packed_f32(ptr(ap), ptr(wp), ptr(bias), S * B, KK, N, 1, ptr(out), ptr(ws), 16, st) < 0
    assert lib.apgs_bmnist_template_match_packed_f32(ptr(f_t), 0, f_t.stride(0), ptr(t_k), t_k.stride(0), t_k.stride(1), S, B, 64, 28,
                                                     ptr(ap), st) == -5                                    # APGS_E_UNSUPPORTED


@pytest.mark.parametrize("M,H,H2,Z", [(640, 400, 200, 2), (13, 400, 200, 2), (1, 37, 9, 3), (77, 130, 65, 1)])
def test_enc_coor_heads_in_one_launch_match_the_modules(M, H, H2, Z):
    """apgs_bmnist_coor_heads_f32 against the reference's nn.Sequential heads (apgs/bmnist/models.py:16-26,31-32) in fp64 and in
    torch fp32: tanh(W2 relu(W1 h + b1) + b2) and exp(...), ragged shapes (rows not a multiple of the 8-row tile, H not a multiple
    of 4, H2 not a multiple of 32), outputs inside their buffers."""
    import copy
    import torch.nn as nn
    from apgs_b200.bmnist import kernels
    torch.manual_seed(M + H)
    mean = nn.Sequential(nn.Linear(H, H2), nn.ReLU(), nn.Linear(H2, Z), nn.Tanh()).to(DEV)
    lstd = nn.Sequential(nn.Linear(H, H2), nn.ReLU(), nn.Linear(H2, Z)).to(DEV)
    hidden = torch.relu(torch.randn(M, H, device=DEV))
    w1t, b1 = kernels.coor_heads_weights(mean[0].weight, mean[0].bias, lstd[0].weight, lstd[0].bias)
    assert tuple(w1t.shape) == ((H + 3) // 4, 2 * H2, 4)
    with torch.no_grad():
        qm, qs = kernels.coor_heads(hidden, w1t, b1, mean[2].weight, mean[2].bias, lstd[2].weight, lstd[2].bias)
        rm32, rs32 = mean(hidden), lstd(hidden).exp()
        mean64, lstd64 = copy.deepcopy(mean).double(), copy.deepcopy(lstd).double()
        rm, rs = mean64(hidden.double()), lstd64(hidden.double()).exp()
    assert tuple(qm.shape) == (M, Z) and tuple(qs.shape) == (M, Z)
    em, es = float((qm.double() - rm).abs().max()), float(((qs.double() - rs).abs() / rs).max())
    em32, es32 = float((rm32.double() - rm).abs().max()), float(((rs32.double() - rs).abs() / rs).max())
    assert em <= 2e-6 and es <= 2e-6, (em, es)
    assert em <= 4 * em32 + 3e-7 and es <= 4 * es32 + 3e-7, (em, em32, es, es32)
    # leading shape kept; deterministic
    if M % 2 == 0:
        q2, _ = kernels.coor_heads(hidden.view(2, M // 2, H), w1t, b1, mean[2].weight, mean[2].bias, lstd[2].weight, lstd[2].bias)
        assert torch.equal(q2.view(M, Z), qm)


def test_enc_coor_fused_heads_switch_and_weight_cache():
    from apgs_b200.bmnist import models as M
    torch.manual_seed(3)
    enc = M.Enc_coor(4761, 400, 2).to(DEV)
    conved = torch.softmax(torch.randn(10, 64, 4761, device=DEV) * 2, -1)
    with torch.no_grad():
        m1, s1 = enc.params(conved)
        M.Enc_coor.fused_heads = False
        try:
            m0, s0 = enc.params(conved)
        finally:
            M.Enc_coor.fused_heads = True
        assert float((m1 - m0).abs().max()) <= 2e-6 and float(((s1 - s0).abs() / s0).max()) <= 2e-6
        enc.where_mean[0].weight.mul_(0.5)                              # the cached transposed weights follow the parameter
        enc.where_log_std[0].bias.add_(0.1)
        m2, s2 = enc.params(conved)
        M.Enc_coor.fused_heads = False
        try:
            m3, s3 = enc.params(conved)
        finally:
            M.Enc_coor.fused_heads = True
    assert float((m2 - m3).abs().max()) <= 2e-6 and float(((s2 - s3).abs() / s3).max()) <= 2e-6
    assert not torch.equal(m2, m1)


@pytest.mark.parametrize("S,B", [(10, 64), (2, 3)])
def test_template_match_on_the_residual_frame_is_bit_identical(S, B):
    """apgs_bmnist_template_match_residual_f32: forming frame - paste(digit0) - paste(digit1) while the image is staged gives the
    bits of paste_subtract (k_stn_paste's residual mode) followed by the match (apgs/bmnist/objectives.py:72-89), for one and
    two previous digits, S-shared frames (stride 0), strided template views, fp32 and packed outputs; digits partly outside the
    frame included."""
    from apgs_b200.bmnist import kernels, Affine_Transformer
    torch.manual_seed(S + 7 * B)
    FP, DP, T, K = 96, 28, 3, 3
    AT = Affine_Transformer(FP, DP, True, DEV)
    frames = torch.rand(B, T, FP, FP, device=DEV).unsqueeze(0).expand(S, -1, -1, -1, -1)     # stride-0 particle axis
    tmpl = torch.rand(S, B, K, DP, DP, device=DEV)
    z = (torch.rand(S, B, K, 2, device=DEV) * 2.4 - 1.2)                                    # some digits stick out of the frame
    f_t = frames[:, :, 1]
    left1 = AT.subtract_digit(f_t, tmpl[:, :, 0], z[:, :, 0].contiguous())
    left2 = AT.subtract_digit(left1, tmpl[:, :, 1], z[:, :, 1].contiguous())
    p1 = [(tmpl[:, :, 0], z[:, :, 0].contiguous())]
    p2 = p1 + [(tmpl[:, :, 1], z[:, :, 1].contiguous())]
    for sm in (True, False):
        assert torch.equal(kernels.template_match(f_t, tmpl[:, :, 1], softmax=sm, prev=p1, lin_fp=AT.lin_fp),
                           kernels.template_match(left1, tmpl[:, :, 1], softmax=sm))
        assert torch.equal(kernels.template_match(f_t, tmpl[:, :, 2], softmax=sm, prev=p2, lin_fp=AT.lin_fp),
                           kernels.template_match(left2, tmpl[:, :, 2], softmax=sm))
    a = kernels.template_match(f_t, tmpl[:, :, 2], packed=True, prev=p2, lin_fp=AT.lin_fp)
    b = kernels.template_match(left2, tmpl[:, :, 2], packed=True)
    assert torch.equal(a.unpack(), b.unpack())
    assert not torch.equal(kernels.template_match(f_t, tmpl[:, :, 2]), kernels.template_match(left2, tmpl[:, :, 2]))
    with pytest.raises(ValueError):
        kernels.template_match(f_t, tmpl[:, :, 2], prev=p2 + p1, lin_fp=AT.lin_fp)
    with pytest.raises(ValueError):
        kernels.template_match(f_t, tmpl[:, :, 2], prev=p1)
