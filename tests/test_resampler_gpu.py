"""GPU parity: resampling kernels (through the C ABI) vs the oracle and the golden fixtures.

Contract (DESIGN.md "index exactness"):
  Level 1  given the normalised weights (systematic) / categorical probabilities (multinomial)
           and the uniforms, ancestors equal the reference run on CPU BIT FOR BIT.
  Level 2  from log-weights the kernel's own expf differs from ATen's vectorised softmax by
           ulps, so an ancestor may move by one only where a position sits on a CDF boundary.
"""
import numpy as np
import pytest
import torch

from oracle import resampler as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from apgs_b200 import ops
    return ops


def _cases(g):
    return [(i, int(S), int(B)) for i, (S, B, _) in enumerate(g["cases"])]


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ----------------------------------------------------------------------------- Level 1
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_L1_systematic_bit_exact_vs_reference_golden(golden, idx_dtype):
    g, ops = golden("resampler"), _ops()
    for ci, S, B in _cases(g):
        a = ops.sample_ancestors(w_norm=_cuda(g["c%d_w" % ci]), strategy="systematic",
                                 uniforms=_cuda(g["c%d_u_sys" % ci]), index_dtype=idx_dtype)
        assert a.dtype == idx_dtype and tuple(a.shape) == (S, B)
        assert np.array_equal(a.cpu().numpy(), g["c%d_anc_sys" % ci]), "case %d" % ci


def test_L1_multinomial_bit_exact_vs_reference_golden(golden):
    g, ops = golden("resampler"), _ops()
    for ci, S, B in _cases(g):
        u = _cuda(g["c%d_u_mul" % ci])                       # (B,S): the reference's generator order
        a = ops.sample_ancestors(w_norm=_cuda(g["c%d_p_mul" % ci]), strategy="multinomial", uniforms=u)
        assert np.array_equal(a.cpu().numpy(), g["c%d_anc_mul" % ci]), "case %d" % ci
        a2 = ops.sample_ancestors(w_norm=_cuda(g["c%d_p_mul" % ci]), strategy="multinomial",
                                  uniforms=u.t().contiguous())   # (S,B) layout, same numbers
        assert torch.equal(a, a2)


@pytest.mark.parametrize("S,B,sig", [(1000, 64, 2.0), (5000, 8, 10.0), (8192, 3, 0.1), (37, 1, 1.0), (2, 300, 3.0)])
def test_L1_seeded_vs_oracle(S, B, sig):
    ops = _ops()
    torch.manual_seed(S + B)
    w = torch.softmax(torch.randn(S, B) * sig, 0)
    u = torch.rand(B)
    a = ops.sample_ancestors(w_norm=w.to(DEV), strategy="systematic", uniforms=u.to(DEV))
    assert np.array_equal(a.cpu().numpy(), orc.systematic_ancestors(w.numpy(), u.numpy()))
    p = orc.categorical_probs(w)
    um = torch.rand(B, S, dtype=torch.float64)
    a = ops.sample_ancestors(w_norm=p.to(DEV), strategy="multinomial", uniforms=um.to(DEV))
    assert np.array_equal(a.cpu().numpy(), orc.multinomial_ancestors(p.numpy(), um.numpy()))


def _boundary_ok(ncdf_b, pos, a_ref, a_gpu, tol):
    lo, hi = min(a_ref, a_gpu), max(a_ref, a_gpu)
    return all(abs(float(ncdf_b[j]) - float(pos)) <= tol for j in range(lo, hi))


@pytest.mark.parametrize("S,B", [(20000, 4), (60000, 3)])
def test_blocked_scan_long_particle_axis(S, B):
    """S > 8192 -> strip-parallel fp64 scan (S=60000 also exercises the global-workspace tile)."""
    ops = _ops()
    torch.manual_seed(S)
    w = torch.softmax(torch.randn(S, B) * 2.0, 0)
    u = torch.rand(B)
    a = ops.sample_ancestors(w_norm=w.to(DEV), strategy="systematic", uniforms=u.to(DEV)).cpu().numpy()
    ref = orc.systematic_ancestors(w.numpy(), u.numpy())
    bad = np.argwhere(a != ref)
    assert len(bad) <= 2, "fp64 re-association should flip (almost) nothing: %d" % len(bad)
    um = torch.rand(S, B, dtype=torch.float64)
    a = ops.sample_ancestors(w_norm=w.to(DEV), strategy="multinomial", uniforms=um.to(DEV)).cpu().numpy()
    ref = orc.multinomial_ancestors(w.numpy(), um.t().numpy(), cdf_mode="f64")
    assert (a != ref).sum() <= 2


# ----------------------------------------------------------------------------- Level 2
@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_L2_from_log_weights(golden, strategy):
    g, ops = golden("resampler"), _ops()
    tot = bad_tot = 0
    for ci, S, B in _cases(g):
        lw = _cuda(g["c%d_log_w" % ci])
        if strategy == "systematic":
            u = g["c%d_u_sys" % ci]
            a, w, ess, lse = ops.sample_ancestors(log_w=lw, strategy=strategy, uniforms=_cuda(u), want_stats=True)
            ref = g["c%d_anc_sys" % ci]
            ncdf = orc.cdf_f64acc(g["c%d_w" % ci])
            pos = orc.systematic_positions(u, S)
        else:
            u = g["c%d_u_mul" % ci]
            a, w, ess, lse = ops.sample_ancestors(log_w=lw, strategy=strategy, uniforms=_cuda(u), want_stats=True)
            ref = g["c%d_anc_mul" % ci]
            ncdf = orc.cdf_seq_f32(g["c%d_p_mul" % ci])
            pos = u
        a = a.cpu().numpy()
        # W1 outputs: 1e-5 relative (north_star tolerance for fp32 log-weights / log-marginals)
        np.testing.assert_allclose(w.cpu().numpy(), g["c%d_w" % ci], rtol=1e-5, atol=1e-37)
        np.testing.assert_allclose(ess.cpu().numpy(), g["c%d_ess" % ci], rtol=1e-5)
        np.testing.assert_allclose(lse.cpu().numpy(), g["c%d_lse" % ci], rtol=1e-5, atol=1e-5)
        bad = np.argwhere(a != ref)
        for s, b in bad:
            assert abs(int(a[s, b]) - int(ref[s, b])) <= 2
            assert _boundary_ok(ncdf[b], pos[b, s], int(ref[s, b]), int(a[s, b]), 4e-6), (ci, s, b)
        tot += a.size
        bad_tot += len(bad)
    assert bad_tot <= 2e-3 * tot, "mismatch fraction %g" % (bad_tot / tot)
    print("L2 %s: %d / %d ancestors differ (all on CDF boundaries)" % (strategy, bad_tot, tot))


def test_weight_stats_and_local(golden):
    g, ops = golden("resampler"), _ops()
    for ci, S, B in _cases(g):
        w, ess, lse = ops.weight_stats(_cuda(g["c%d_log_w" % ci]))
        np.testing.assert_allclose(w.cpu().numpy(), g["c%d_w" % ci], rtol=1e-5, atol=1e-37)
        np.testing.assert_allclose(ess.cpu().numpy(), g["c%d_ess" % ci], rtol=1e-5)
        np.testing.assert_allclose(lse.cpu().numpy(), g["c%d_lse" % ci], rtol=1e-5, atol=1e-5)
    torch.manual_seed(5)
    x = torch.randn(37, 11, 60) * 3
    w_ref, lw_ref = orc.local_weight_stats(x)
    w, lw = ops.local_weight_stats(x.to(DEV))
    torch.testing.assert_close(w.cpu(), w_ref, rtol=1e-5, atol=1e-30)
    torch.testing.assert_close(lw.cpu(), lw_ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------- gather
def test_gather_vs_reference_golden(golden):
    g, ops = golden("resampler"), _ops()
    for ci, S, B in _cases(g):
        if "c%d_var4" % ci not in g:
            continue
        for dt in (torch.int64, torch.int32):
            o4, = ops.gather_ancestors([_cuda(g["c%d_var4" % ci])], _cuda(g["c%d_anc_sys" % ci]).to(dt))
            o5, = ops.gather_ancestors([_cuda(g["c%d_var5" % ci])], _cuda(g["c%d_anc_mul" % ci]).to(dt))
            assert np.array_equal(o4.cpu().numpy(), g["c%d_res4" % ci])
            assert np.array_equal(o5.cpu().numpy(), g["c%d_res5" % ci])


@pytest.mark.parametrize("S,B", [(1, 1), (7, 3), (100, 33), (1000, 64)])
def test_gather_every_vector_width_and_multi_latent(S, B):
    """inner chunk sizes hit the 16/8/4/2/1-byte vector paths, whole-row and split-row blocks,
    and the 8-latent single launch."""
    ops = _ops()
    torch.manual_seed(S * 131 + B)
    anc = torch.randint(0, S, (S, B))
    lat = [torch.randn(S, B, 60, 3), torch.randn(S, B, 3, 2), torch.randn(S, B, 1), torch.randn(S, B, 200, 4),
           torch.randn(S, B, 5, 1, 2), torch.randint(0, 255, (S, B, 3), dtype=torch.uint8),
           torch.randn(S, B, 7).to(torch.bfloat16), torch.randn(S, B, 2, dtype=torch.float64)]
    outs = ops.gather_ancestors([v.to(DEV) for v in lat], anc.to(DEV))
    for v, o in zip(lat, outs):
        ref = orc.gather_ancestors(v.view(torch.uint8).numpy() if v.dtype == torch.bfloat16 else v.numpy(), anc.numpy())
        got = o.cpu()
        got = got.view(torch.uint8).numpy() if v.dtype == torch.bfloat16 else got.numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), tuple(v.shape)


def test_gather_misaligned_views():
    ops = _ops()
    S, B = 50, 9
    anc = torch.randint(0, S, (S, B))
    base = torch.randn(S * B * 6 + 1, device=DEV)
    v = base[1:].view(S, B, 3, 2)                    # 4-byte aligned only
    o, = ops.gather_ancestors([v], anc.to(DEV))
    assert np.array_equal(o.cpu().numpy(), orc.gather_ancestors(v.cpu().numpy(), anc.numpy()))


# ---------------------------------------------------------------- reference-surface object
@pytest.mark.parametrize("strategy", ["systematic", "multinomial"])
def test_resampler_surface(strategy):
    from apgs_b200.resampler import Resampler
    S, B = 10, 20                                    # BASELINE.json configs[0]
    torch.manual_seed(0)
    r = Resampler(strategy, S, True, DEV)
    lw = (torch.randn(S, B) * 2).to(DEV)
    a = r.sample_ancestral_index(lw)
    assert a.dtype == torch.int64 and tuple(a.shape) == (S, B)
    assert int(a.min()) >= 0 and int(a.max()) < S
    if strategy == "systematic":
        assert bool((a[1:] >= a[:-1]).all()), "systematic ancestors are monotone in s"
    tau, z, zw = torch.randn(S, B, 3, 2, device=DEV), torch.randn(S, B, 60, 3, device=DEV), torch.randn(S, B, 4, 3, 2, device=DEV)
    t2, z2, zw2 = r.resample_4dims(tau, a), r.resample_4dims(z, a), r.resample_5dims(zw, a)
    idx = a.view(S, B, 1, 1)
    assert torch.equal(t2, torch.gather(tau, 0, idx.expand_as(tau)))
    assert torch.equal(z2, torch.gather(z, 0, idx.expand_as(z)))
    assert torch.equal(zw2, torch.gather(zw, 0, a.view(S, B, 1, 1, 1).expand_as(zw)))
    # fused entry == separate calls given the same uniforms
    u = r.draw_uniforms(S, B, DEV)
    a = r.sample_ancestral_index(lw, uniforms=u)
    (t3, z3), a3, (w, ess, lse) = r.resample([tau, z], lw, uniforms=u, want_stats=True, want_index=True)
    assert torch.equal(a, a3)
    assert torch.equal(t3, r.resample_4dims(tau, a)) and torch.equal(z3, r.resample_4dims(z, a))
    torch.testing.assert_close(w, torch.softmax(lw, 0), rtol=1e-5, atol=1e-30)
    torch.testing.assert_close(lse, torch.logsumexp(lw, 0), rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        Resampler("stratified", S, True, DEV)
    if strategy == "systematic":
        with pytest.raises(ValueError):
            r.sample_ancestral_index(lw[:5])


# --------------------------------------------------- size-independent properties at full size
def test_properties_at_gmm_full_size():
    """BASELINE.json configs[1]: S=1000, B=1024, GMM latents (tau, mu: 24 B; z: 720 B per particle)."""
    from apgs_b200.resampler import Resampler
    S, B = 1000, 1024
    torch.manual_seed(1)
    r = Resampler("systematic", S, True, DEV)
    lw = torch.randn(S, B, device=DEV) * 3
    lw[::7, 5] = -float("inf")
    tau = torch.randn(S, B, 3, 2, device=DEV)
    mu = torch.randn(S, B, 3, 2, device=DEV)
    z = torch.nn.functional.one_hot(torch.randint(0, 3, (S, B, 60), device=DEV), 3).float()
    u = r.draw_uniforms(S, B, DEV)
    (t2, m2, z2), a, (w, ess, lse) = r.resample([tau, mu, z], lw, uniforms=u, want_stats=True, want_index=True)
    assert int(a.min()) >= 0 and int(a.max()) < S
    assert bool((a[1:] >= a[:-1]).all())
    dead = torch.isinf(lw[:, 5]).nonzero().flatten()
    assert not torch.isin(a[:, 5], dead).any(), "-inf weights never selected"
    # gather == torch.gather with the reference's expanded index (apgs/resampler.py:48-49)
    assert torch.equal(z2, torch.gather(z, 0, a.view(S, B, 1, 1).expand_as(z)))
    assert torch.equal(t2, torch.gather(tau, 0, a.view(S, B, 1, 1).expand_as(tau)))
    assert torch.equal(m2, torch.gather(mu, 0, a.view(S, B, 1, 1).expand_as(mu)))
    # systematic counts: |count_i - S w_i| < 1 (+ fp slack)
    counts = torch.zeros(S, B, device=DEV).scatter_add_(0, a, torch.ones(S, B, device=DEV))
    assert float((counts - S * w).abs().max()) < 1.0 + 1e-2
    # permutation equivariance over instances
    perm = torch.randperm(B, device=DEV)
    a_p = r.sample_ancestral_index(lw[:, perm].contiguous(), uniforms=u[perm].contiguous())
    assert torch.equal(a_p, a[:, perm])
    # ess in [1, S], lse matches torch
    assert float(ess.min()) >= 1 - 1e-4 and float(ess.max()) <= S + 1e-2
    torch.testing.assert_close(lse, torch.logsumexp(lw, 0), rtol=1e-5, atol=1e-5)


def test_multinomial_distribution_chi2():
    """chi-square of multinomial offspring counts against S*w."""
    from apgs_b200.resampler import Resampler
    S, B = 4000, 8
    torch.manual_seed(3)
    r = Resampler("multinomial", S, True, DEV)
    lw = torch.randn(S, B, device=DEV) * 0.5
    a = r.sample_ancestral_index(lw)
    w = torch.softmax(lw, 0)
    counts = torch.zeros(S, B, device=DEV).scatter_add_(0, a, torch.ones(S, B, device=DEV))
    chi2 = (((counts - S * w) ** 2) / (S * w)).sum(0)           # dof = S-1, sd = sqrt(2(S-1))
    assert bool(((chi2 - (S - 1)).abs() < 6 * (2 * (S - 1)) ** 0.5).all()), chi2


@pytest.mark.parametrize("S,B,shape", [(203, 67, (60, 3)), (50, 33, (200, 4)), (1000, 13, (50, 4)), (77, 129, (64,))])
def test_gather_large_rows(S, B, shape):
    """> 8 MB latents with 720 B / 3200 B / 800 B / 256 B rows next to a 24 B latent in the same launch:
    ragged last chunk, int32 and int64 ancestors."""
    ops = _ops()
    torch.manual_seed(S + B)
    reps = max(1, int(9e6 // (S * B * int(np.prod(shape)) * 4)) + 1)
    Sx = S * reps                                           # grow S until the latent is > 8 MB
    anc = torch.randint(0, Sx, (Sx, B), device=DEV)
    v = torch.randn(Sx, B, *shape, device=DEV)
    small = torch.randn(Sx, B, 3, 2, device=DEV)
    for dt in (torch.int64, torch.int32):
        o, o2 = ops.gather_ancestors([v, small], anc.to(dt))
        idx = anc.view(Sx, B, *([1] * len(shape))).expand_as(v)
        assert torch.equal(o, torch.gather(v, 0, idx))
        assert torch.equal(o2, torch.gather(small, 0, anc.view(Sx, B, 1, 1).expand_as(small)))


# ------------------------------------------------------------------ split-S pipeline (S beyond a shared-memory tile)
@pytest.mark.parametrize("S,B", [(60000, 3), (100000, 1), (131072, 33), (70001, 40)])
def test_split_S_pipeline_from_log_weights(S, B):
    """S > ~51k runs resample_large.cu: weights / ESS / lse against torch, systematic ancestors against the oracle run
    on the kernel's own weights (level 1), and the fused resample against a plain gather."""
    ops = _ops()
    torch.manual_seed(S + B)
    lw = torch.randn(S, B) * 2.0
    lw[:, 0] = torch.randn(S) * 12.0                                      # one nearly degenerate instance
    u = torch.rand(B)
    w, ess, lse = ops.weight_stats(lw.to(DEV))
    w_ref = torch.softmax(lw.double(), 0)
    np.testing.assert_allclose(w.cpu().numpy(), w_ref.float().numpy(), rtol=2e-5, atol=1e-37)
    np.testing.assert_allclose(ess.cpu().numpy(), (1.0 / (w_ref ** 2).sum(0)).numpy(), rtol=2e-5)
    np.testing.assert_allclose(lse.cpu().numpy(), torch.logsumexp(lw.double(), 0).numpy(), rtol=1e-6, atol=1e-5)
    a = ops.sample_ancestors(log_w=lw.to(DEV), strategy="systematic", uniforms=u.to(DEV))
    assert a.dtype == torch.int64 and int(a.min()) >= 0 and int(a.max()) < S
    assert bool((a[1:] >= a[:-1]).all())
    ref = orc.systematic_ancestors(w.cpu().numpy(), u.numpy())
    bad = (a.cpu().numpy() != ref).sum()
    assert bad <= 2, "fp64 re-association should flip (almost) nothing: %d" % bad
    a32 = ops.sample_ancestors(log_w=lw.to(DEV), strategy="systematic", uniforms=u.to(DEV), index_dtype=torch.int32)
    assert torch.equal(a32.long(), a)
    # fused entry
    lat = torch.randn(S, B, 3, device=DEV)
    (out,), _, _ = ops.resample([lat], log_w=lw.to(DEV), strategy="systematic", uniforms=u.to(DEV))
    assert torch.equal(out, torch.gather(lat, 0, a.unsqueeze(-1).expand_as(lat)))


def test_split_S_pipeline_multinomial_and_given_weights():
    ops = _ops()
    S, B = 80000, 5
    torch.manual_seed(1)
    w = torch.softmax(torch.randn(S, B) * 2.0, 0)
    um = torch.rand(S, B, dtype=torch.float64)
    a = ops.sample_ancestors(w_norm=w.to(DEV), strategy="multinomial", uniforms=um.to(DEV)).cpu().numpy()
    ref = orc.multinomial_ancestors(w.numpy(), um.t().numpy(), cdf_mode="f64")
    assert (a != ref).sum() <= 2
    u = torch.rand(B)
    a = ops.sample_ancestors(w_norm=w.to(DEV), strategy="systematic", uniforms=u.to(DEV)).cpu().numpy()
    assert (a != orc.systematic_ancestors(w.numpy(), u.numpy())).sum() <= 2
    # the exact sequential chain is still available on request (single-CTA global-tile variant)
    a = ops.sample_ancestors(w_norm=w.to(DEV), strategy="systematic", uniforms=u.to(DEV), cdf_mode=1).cpu().numpy()     # APGS_CDF_EXACT
    assert np.array_equal(a, orc.systematic_ancestors(w.numpy(), u.numpy()))


# ------------------------------------------------------------------ guard bands (the pool has no compute-sanitizer)
@pytest.mark.parametrize("S,B,strategy", [(70001, 5, "systematic"), (60000, 33, "multinomial"), (1000, 37, "systematic"),
                                          (9000, 3, "multinomial")])
def test_ancestor_kernels_stay_inside_their_buffers(S, B, strategy):
    """raw C-ABI call with every output and the workspace embedded in sentinel-filled guard bands: nothing outside the
    documented extents may be written (covers the shared-memory tile, blocked and split-S variants)."""
    from apgs_b200 import _lib
    from apgs_b200._lib import ptr
    lib = _lib.load()
    torch.manual_seed(S)
    G = 4096                                                     # guard elements on both sides
    lw = (torch.randn(S, B) * 2).to(DEV)
    if strategy == "systematic":
        u, us, ub = torch.rand(B, device=DEV), 0, 0
    else:
        u, us, ub = torch.rand(B, S, dtype=torch.float64, device=DEV), 1, S

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=DEV)
        return buf, buf[G:G + n]
    b64, a64 = guarded(S * B, torch.int64, -7)
    b32, a32 = guarded(S * B, torch.int32, -7)
    bw, w = guarded(S * B, torch.float32, -7.0)
    be, ess = guarded(B, torch.float32, -7.0)
    bl, lse = guarded(B, torch.float32, -7.0)
    nws = lib.apgs_ancestors_workspace_bytes(S, B)
    bws = torch.full((nws + 2 * 256 * 16,), 0x5A, dtype=torch.uint8, device=DEV)
    off = (-bws.data_ptr()) % 256 + 256 * 8                      # keep the interior 256-byte aligned
    ws = bws[off:off + nws]
    rc = lib.apgs_sample_ancestors_f32(ptr(lw), None, S, B, _lib.SYSTEMATIC if strategy == "systematic" else _lib.MULTINOMIAL,
                                       ptr(u), us, ub, 0, ptr(a64), ptr(a32), ptr(w), ptr(ess), ptr(lse),
                                       ptr(ws) if nws else None, nws, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.apgs_last_error()
    torch.cuda.synchronize()
    for buf, fill in ((b64, -7), (b32, -7), (bw, -7.0), (be, -7.0), (bl, -7.0)):
        assert bool((buf[:G] == fill).all()) and bool((buf[-G:] == fill).all())
    assert bool((bws[:off] == 0x5A).all()) and bool((bws[off + nws:] == 0x5A).all())
    assert int(a64.min()) >= 0 and int(a64.max()) < S and torch.equal(a64, a32.long())
    assert bool(torch.isfinite(ess).all()) and bool(torch.isfinite(lse).all())
    np.testing.assert_allclose(w.view(S, B).sum(0).cpu().numpy(), 1.0, rtol=1e-4)


def test_multinomial_uniform_layout_is_explicit_when_ambiguous():
    """with S == B a (B,S) tensor in the reference's generator order must not be read as (S,B) silently."""
    from apgs_b200.resampler import Resampler
    ops = _ops()
    S = B = 16
    torch.manual_seed(0)
    w = torch.softmax(torch.randn(S, B) * 2, 0)
    p = orc.categorical_probs(w)
    u_bs = torch.rand(B, S, dtype=torch.float64)
    ref = orc.multinomial_ancestors(p.numpy(), u_bs.numpy())
    with pytest.raises(ValueError):
        ops.sample_ancestors(w_norm=p.to(DEV), strategy="multinomial", uniforms=u_bs.to(DEV))
    a = ops.sample_ancestors(w_norm=p.to(DEV), strategy="multinomial", uniforms=u_bs.to(DEV), uniform_layout="bs")
    assert np.array_equal(a.cpu().numpy(), ref)
    a2 = ops.sample_ancestors(w_norm=p.to(DEV), strategy="multinomial", uniforms=u_bs.t().contiguous().to(DEV), uniform_layout="sb")
    assert np.array_equal(a2.cpu().numpy(), ref)
    # the Resampler draws torch.rand(B, S, float64) -- the reference's generator stream -- and reads it as such
    r = Resampler("multinomial", S, True, DEV)
    u = r.draw_uniforms(S, B, DEV)
    assert tuple(u.shape) == (B, S) and u.dtype == torch.float64
    lw = torch.log(w).to(DEV)
    a3 = r.sample_ancestral_index(lw, uniforms=u)
    assert int(a3.min()) >= 0 and int(a3.max()) < S


def test_resample_dims_propagate_gradients_like_torch_gather():
    """resample_4dims / 5dims are differentiable w.r.t. the latent, as the reference's torch.gather is."""
    from apgs_b200.resampler import Resampler
    S, B = 12, 5
    torch.manual_seed(1)
    r = Resampler("systematic", S, True, DEV)
    a = r.sample_ancestral_index((torch.randn(S, B) * 2).to(DEV))
    for shape in ((S, B, 3, 2), (S, B, 4, 3, 2)):
        v = torch.randn(*shape, device=DEV, requires_grad=True)
        g = torch.randn(*shape, device=DEV)
        out = r.resample_4dims(v, a) if len(shape) == 4 else r.resample_5dims(v, a)
        assert out.requires_grad
        (out * g).sum().backward()
        v2 = v.detach().clone().requires_grad_(True)
        idx = a.view((S, B) + (1,) * (len(shape) - 2)).expand(*shape)
        (torch.gather(v2, 0, idx) * g).sum().backward()
        assert torch.equal(out.detach(), torch.gather(v2, 0, idx).detach())
        torch.testing.assert_close(v.grad, v2.grad, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("S,B,N", [(1000, 64, 60), (100, 37, 200), (5, 3, 7), (3300, 4, 9), (1, 2, 5)])
def test_W2_local_softmax_tile_kernel(S, B, N):
    """per-point softmax over S (apgs/gmm/objectives.py:216-218): the single-pass shared-memory tile kernel (and, beyond
    a tile, the three-pass one: S = 3300) against torch, incl. -inf entries and ragged column tiles; columns sum to 1."""
    ops = _ops()
    torch.manual_seed(S + N)
    x = (torch.randn(S, B, N) * 4).to(DEV)
    if S > 4:
        x[::3, 0, 0] = -float("inf")
    w, lw = ops.local_weight_stats(x)
    ref = torch.softmax(x.double(), 0)
    torch.testing.assert_close(w.double(), ref, rtol=1e-5, atol=1e-30)
    torch.testing.assert_close(w.sum(0), torch.ones(B, N, device=DEV), rtol=1e-5, atol=1e-5)
    fin = torch.isfinite(x).all(0)
    torch.testing.assert_close(lw[:, fin.all(-1)].double(), x.double().sum(-1)[:, fin.all(-1)], rtol=1e-5, atol=1e-4)
