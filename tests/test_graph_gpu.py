"""GPU: a whole GMM objective captured into a CUDA graph replays with fresh random numbers and the same statistics."""
import pytest
import torch

from oracle import gmm as og

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_graphed_objective_replays():
    from apgs_b200.gmm import apg_objective, init_apg_models
    from apgs_b200.graph import GraphedObjective
    from apgs_b200.resampler import Resampler
    S, B, sweeps = 10, 20, 5                                    # BASELINE.json configs[0]
    torch.manual_seed(0)
    models = init_apg_models(3, 2, 32, True, DEV)
    x = og.simulate(B).to(DEV)
    flags = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}
    rs = Resampler('systematic', S, True, DEV)
    g = GraphedObjective(lambda xx: apg_objective(models, xx, flags, sweeps, 'decomposed', rs), x)
    a = g(x, clone=True)
    b = g(x, clone=True)
    assert tuple(a['ess'].shape) == (sweeps, B) and tuple(a['density'].shape) == (sweeps, S, B)
    assert bool(torch.isfinite(a['density']).all()) and bool(torch.isfinite(b['density']).all())
    assert not torch.equal(a['density'], b['density']), "replays must draw fresh random numbers"
    # a different batch through the same graph
    x2 = og.simulate(B).to(DEV)
    c = g(x2, clone=True)
    with torch.no_grad():
        ref = apg_objective(models, x2, flags, sweeps, 'decomposed', rs)
    assert abs(float(c['density'][0].mean()) - float(ref['density'][0].mean())) < 0.25 * abs(float(ref['density'][0].mean()))
