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
    M.Enc_coor.linear_kernel = "mma"
    try:
        with torch.no_grad():
            mean_x, std_x = enc.params(conved)
    finally:
        M.Enc_coor.linear_kernel = "tc5"
    assert float((mean_x - mean_t).abs().max()) <= 3e-6 and float(((std_x - std_t).abs() / std_t).max()) <= 1e-5
    # training mode keeps the differentiable torch modules
    for p in enc.parameters():
        p.requires_grad_(True)
    m2, _ = enc.params(conved)
    assert m2.requires_grad
