"""CPU tests of the I/O around the path: simulators with the reference generators' distributions, the
reference's `.npy` dataset files, and `weights/cp-<version>` checkpoints that load in the reference (and back) when the
reference is available (/root/reference here, oracle/_ref on the GPU box)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import ref_arm


def test_gmm_simulator_and_dataset_files(tmp_path):
    from apgs_b200.gmm import sim
    torch.manual_seed(0)
    x, z, tau, mu = sim.simulate(400, return_latents=True)
    assert tuple(x.shape) == (400, 60, 2) and tuple(z.shape) == (400, 60, 3) and bool((z.sum(-1) == 1).all())
    # x_n ~ N(mu_z, tau_z^-1/2): standardised residuals are N(0,1)
    lab = z.argmax(-1).unsqueeze(-1).expand(-1, -1, 2)
    r = (x - torch.gather(mu, 1, lab)) * torch.gather(tau, 1, lab).sqrt()
    assert abs(float(r.mean())) < 0.02 and abs(float(r.std()) - 1) < 0.02
    assert abs(float(tau.mean()) - 1.0) < 0.1                      # Gamma(2, 2): mean 1
    ob, asg = sim.save_dataset(str(tmp_path), 9)
    assert np.load(ob).shape == (9, 60, 2) and np.load(ob).dtype == np.float64 and np.load(asg).shape == (9, 60, 3)
    d, a = sim.load_dataset(str(tmp_path))
    assert d.dtype == torch.float32 and tuple(a.shape) == (9, 60, 3)


def test_dmm_simulator_rings(tmp_path):
    from apgs_b200.dmm import sim
    torch.manual_seed(1)
    x, st, c, ang = sim.simulate(50, return_latents=True)
    assert tuple(x.shape) == (50, 200, 2) and bool((st.sum(1) == 50).all())
    cen = torch.einsum('bnk,bkd->bnd', st, c)
    rad = (x - cen).norm(dim=-1)
    assert abs(float(rad.mean()) - 2.0) < 0.02 and float((rad - 2).abs().max()) < 0.8       # noise 0.1
    assert float(ang.min()) >= 0 and float(ang.max()) <= 2 * math.pi * (1 + 1 / 24) + 1e-5
    assert abs(float(c.std()) - 3.0) < 0.5
    with pytest.raises(ValueError):
        sim.simulate(2, N=203)
    f = sim.save_dataset(str(tmp_path / "dmm"), 6)
    assert tuple(sim.load_dataset(f).shape) == (6, 200, 2)


def test_bmnist_simulator_trajectories_and_frames(tmp_path):
    from apgs_b200.bmnist import sim
    from oracle import bmnist as ob
    torch.manual_seed(2)
    fr, X, V = sim.simulate(6, return_latents=True)
    assert tuple(fr.shape) == (6, 10, 96, 96) and float(fr.min()) >= 0 and float(fr.max()) <= 1
    assert float(X.abs().max()) <= 1.0 + 1e-6
    assert torch.allclose(V.norm(dim=-1), torch.ones(6, 10, 3), atol=1e-5)           # unit speed, sign flips only
    step = (X[:, 1:] - X[:, :-1]).abs()
    free = (V[:, 1:] == V[:, :-1]).all(-1)
    assert torch.allclose(step[free], (V[:, :-1].abs() * 0.3)[free], atol=1e-5)      # free flight: dX = V dt
    # frames are the transformer's paste of the templates, summed over digits and clamped (oracle == reference bit-exact)
    t = sim.blobs()
    ref = torch.clamp(ob.digit_to_frame(t.view(1, 1, 3, 28, 28).expand(1, 6, -1, -1, -1), X.unsqueeze(0)).sum(-3), 0, 1)[0]
    assert torch.equal(fr, ref)
    files = sim.save_dataset(str(tmp_path / "bm"), 5, chunk_size=3)
    assert [tuple(sim.load_dataset(f).shape) for f in files] == [(3, 10, 96, 96), (2, 10, 96, 96)]


@pytest.mark.parametrize("task", ["gmm", "dmm", "bmnist"])
def test_checkpoint_layout_roundtrip(task, tmp_path, monkeypatch):
    from apgs_b200 import checkpoint
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    if task == "gmm":
        from apgs_b200.gmm import models as M
        mk = lambda **kw: M.init_apg_models(3, 2, 32, False, None, **kw)            # noqa: E731
        save = M.save_apg_models
    elif task == "dmm":
        from apgs_b200.dmm import models as M
        mk = lambda **kw: M.init_apg_models(4, 2, 32, 8, 32, 32, 0.5, False, None, **kw)   # noqa: E731
        save = M.save_apg_models
    else:
        from apgs_b200.bmnist import models as M
        mk = lambda **kw: M.init_models(96, 28, 400, 400, 2, 10, False, None, **kw)  # noqa: E731
        save = M.save_models
    models = mk()
    f = save(models, "t1")
    assert f == os.path.join("weights", "cp-t1") and os.path.exists(f)
    ck = torch.load(f)
    assert sorted(ck.keys()) == sorted(k for k in checkpoint.LAYOUT[task] if k is not None)
    again = mk(load_version="t1")
    for a, b in zip(models, again):
        if hasattr(a, "state_dict"):
            for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
                assert ka == kb and torch.equal(va, vb)
    if not ref_arm.available():
        return
    # the reference loads the file this package wrote, and this package loads the reference's
    ref_arm.load()
    if task == "gmm":
        from apgs.gmm.apg_training import init_apg_models as rinit, save_apg_models as rsave
        rmodels = rinit(3, 2, 32, False, None, load_version="t1")
    elif task == "dmm":
        from apgs.dmm.apg_training import init_apg_models as rinit, save_apg_models as rsave
        rmodels = rinit(4, 2, 32, 8, 32, 32, 0.5, False, None, load_version="t1")
    else:
        from apgs.bmnist.apg_training import init_models as rinit, save_models as rsave
        rmodels = rinit(96, 28, 400, 400, 2, 10, False, None, "t1", None)
    for a, b in zip(models, rmodels):
        if hasattr(a, "state_dict") and hasattr(b, "state_dict"):
            sa, sb = a.state_dict(), b.state_dict()
            assert list(sa.keys()) == list(sb.keys())
            assert all(torch.equal(sa[k], sb[k]) for k in sa)
    rsave(rmodels, "t2")
    back = mk(load_version="t2")
    for a, b in zip(models, back):
        if hasattr(a, "state_dict"):
            assert all(torch.equal(v, b.state_dict()[k]) for k, v in a.state_dict().items())
