"""Host wrappers of the GMM reweighting kernels (apgs_b200/csrc/gmm.cu) -- SURVEY.md G1-G4."""
import ctypes as C

import torch

from .. import _lib
from .._lib import GmmOut, GmmPost, GmmPrior, check, ptr
from ..ops import _call, _req, _stream
from . import torch_path

BASE, ZFUSED, ETA_FUSED, LABELS = 1, 2, 4, 8
_SUPPORTED = {}


def supported(K, D):
    """bit mask of the kernel families compiled for (K, D) (apgs_gmm_supported): BASE = G1-G4, ZFUSED = the fused z step,
    ETA_FUSED = the fused eta heads, LABELS = the label-mode entries.  Shapes without BASE run through stock PyTorch
    CUDA ops (torch_path.py) instead of failing -- the reference accepts any --num_clusters / data dimension."""
    key = (int(K), int(D))
    if key not in _SUPPORTED:
        _SUPPORTED[key] = int(_lib.load().apgs_gmm_supported(key[0], key[1]))
    return _SUPPORTED[key]


def _cuda_only(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.ApgsError("apgs_b200 has no CPU fallback: tensors must be CUDA tensors")


def _prior_struct(prior):
    """prior: object with prior_alpha/beta/mu/nu (K,D) and prior_pi (K) CUDA tensors (Generative)."""
    t = [_req(getattr(prior, "prior_" + n), "prior_" + n) for n in ("alpha", "beta", "mu", "nu", "pi")]
    s = GmmPrior(*[v.data_ptr() for v in t])
    return s, t


def _post_struct(post):
    t = [_req(v, "posterior parameter") for v in post]
    return GmmPost(*[v.data_ptr() for v in t]), t


def _x3(x):
    """the kernels read ONE copy of the data per instance: accept (B,N,D) or the reference's
    S-replicated (S,B,N,D) (apgs/gmm/apg_training.py:19) and use its first slice."""
    if x.dim() == 4:
        x = x[0]
    return _req(x, "x", ndim=3)


def nss_posterior(xp, gamma, prior):
    """G1: data_to_stats + posterior_eta (apgs/gmm/kls_gmm.py:92-119) -> alpha, beta, mu, nu (S,B,K,D)."""
    lib = _lib.load()
    xp, gamma = _req(xp, "xp", ndim=4), _req(gamma, "gamma", ndim=4)
    S, B, N, D = xp.shape
    K = gamma.shape[-1]
    if not supported(K, D) & BASE:
        return torch_path.nss_posterior(xp, gamma, prior)
    if tuple(gamma.shape[:3]) != (S, B, N):
        raise ValueError("xp %s and gamma %s disagree" % (tuple(xp.shape), tuple(gamma.shape)))
    pr = [_req(getattr(prior, "prior_" + n), "prior_" + n) for n in ("alpha", "beta", "mu", "nu")]
    outs = [torch.empty(S, B, K, D, dtype=torch.float32, device=xp.device) for _ in range(4)]
    with torch.cuda.device(xp.device), _call("gmm_nss_posterior", 1, 4 * S * B * (N * (K + D) + 4 * K * D)):
        check(lib.apgs_gmm_nss_posterior_f32(ptr(xp), ptr(gamma), S, B, N, K, D, *[ptr(v) for v in pr],
                                             *[ptr(v) for v in outs], _stream()), "apgs_gmm_nss_posterior_f32")
    return tuple(outs)


class Weights(object):
    """outputs of one reweighting step (all (S,B) unless noted; None when not requested)."""
    __slots__ = ("log_w", "log_q_f", "density", "log_w_f", "log_w_b", "log_w_local", "log_q_local")


def _outs(S, B, N, dev, parts, local):
    w = Weights()
    o = GmmOut()
    keep = []
    for name in Weights.__slots__:
        t = None
        if name in ("log_w", "log_q_f", "density") or (parts and name in ("log_w_f", "log_w_b")):
            t = torch.empty(S, B, dtype=torch.float32, device=dev)
        elif local and name in ("log_w_local", "log_q_local"):
            t = torch.empty(S, B, N, dtype=torch.float32, device=dev)
        setattr(w, name, t)
        setattr(o, name, None if t is None else t.data_ptr())
        keep.append(t)
    return w, o


def logw_oneshot(x, post, prior, tau, mu, q_probs, z, parts=False):
    """G4a: oneshot weights (apgs/gmm/objectives.py:74-110)."""
    if not supported(tau.shape[-2], tau.shape[-1]) & BASE:
        _cuda_only(x, tau, mu, z)
        with torch.no_grad():
            return torch_path.logw_oneshot(x, post, prior, tau, mu, q_probs, z, parts)
    lib = _lib.load()
    x = _x3(x)
    z, q_probs, tau, mu = _req(z, "z", ndim=4), _req(q_probs, "q_probs", ndim=4), _req(tau, "tau", ndim=4), _req(mu, "mu", ndim=4)
    S, B, N, K = z.shape
    D = x.shape[-1]
    ps, pk = _post_struct(post)
    rs, rk = _prior_struct(prior)
    w, o = _outs(S, B, N, z.device, parts, False)
    with torch.cuda.device(z.device), _call("gmm_logw_oneshot", 1, 4 * S * B * (2 * N * K + 6 * K * D + 3)):
        check(lib.apgs_gmm_logw_oneshot_f32(ptr(x), C.byref(ps), C.byref(rs), ptr(tau), ptr(mu), ptr(q_probs), ptr(z),
                                            S, B, N, K, D, C.byref(o), _stream()), "apgs_gmm_logw_oneshot_f32")
    return w


def logw_eta(x, z, post, prior, tau, mu, tau_old, mu_old, parts=False):
    """G2: eta-step weights (apgs/gmm/objectives.py:161-195)."""
    if not supported(tau.shape[-2], tau.shape[-1]) & BASE:
        _cuda_only(x, tau, mu, z)
        with torch.no_grad():
            return torch_path.logw_eta(x, z, post, prior, tau, mu, tau_old, mu_old, parts)
    lib = _lib.load()
    x = _x3(x)
    z = _req(z, "z", ndim=4)
    tau, mu, tau_old, mu_old = [_req(v, "tau/mu", ndim=4) for v in (tau, mu, tau_old, mu_old)]
    S, B, N, K = z.shape
    D = x.shape[-1]
    ps, pk = _post_struct(post)
    rs, rk = _prior_struct(prior)
    w, o = _outs(S, B, N, z.device, parts, False)
    with torch.cuda.device(z.device), _call("gmm_logw_eta", 1, 4 * S * B * (N * K + 8 * K * D + 3)):
        check(lib.apgs_gmm_logw_eta_f32(ptr(x), ptr(z), C.byref(ps), C.byref(rs), ptr(tau), ptr(mu), ptr(tau_old),
                                        ptr(mu_old), S, B, N, K, D, C.byref(o), _stream()), "apgs_gmm_logw_eta_f32")
    return w


def logw_eta_stats(stats, center, post, prior, tau, mu, tau_old, mu_old, parts=False):
    """G2 evaluated from the centred one-hot statistics of z (eta_posterior_fused(..., center=)): same outputs as
    logw_eta without a second pass over z (apgs/gmm/objectives.py:161-195)."""
    lib = _lib.load()
    stats, center = _req(stats, "stats", ndim=4), _req(center, "center", ndim=4)
    tau, mu, tau_old, mu_old = [_req(v, "tau/mu", ndim=4) for v in (tau, mu, tau_old, mu_old)]
    S, B, K, D = tau.shape
    if tuple(stats.shape) != (S, B, K, 1 + 2 * D) or tuple(center.shape) != (S, B, K, D):
        raise ValueError("stats/center have shapes %s %s" % (tuple(stats.shape), tuple(center.shape)))
    for v in (mu, tau_old, mu_old):
        if tuple(v.shape) != (S, B, K, D):
            raise ValueError("tau/mu shapes differ")
    ps, pk = _post_struct(post)
    rs, rk = _prior_struct(prior)
    w, o = _outs(S, B, 1, tau.device, parts, False)
    with torch.cuda.device(tau.device), _call("gmm_logw_eta_stats", 1, 4 * S * B * (K * (1 + 2 * D) + 9 * K * D + 3)):
        check(lib.apgs_gmm_logw_eta_stats_f32(ptr(stats), ptr(center), C.byref(ps), C.byref(rs), ptr(tau), ptr(mu),
                                              ptr(tau_old), ptr(mu_old), S, B, K, D, C.byref(o), _stream()),
              "apgs_gmm_logw_eta_stats_f32")
    return w


def draw_logw_eta_stats(stats, center, post, prior, gamma_std, normal_std, tau_old, mu_old, parts=False):
    """logw_eta_stats with the draw fused in: gamma_std = torch._standard_gamma(alpha), normal_std = torch.randn(...)
    (S,B,K,D).  Returns (Weights, tau, mu) with tau, mu bit-identical to Gamma(alpha, beta).sample() and
    Normal(q_mu, 1/sqrt(q_nu tau)).sample() formed from the same variates."""
    lib = _lib.load()
    stats, center = _req(stats, "stats", ndim=4), _req(center, "center", ndim=4)
    g, e, tau_old, mu_old = [_req(v, "draws / old eta", ndim=4) for v in (gamma_std, normal_std, tau_old, mu_old)]
    S, B, K, D = g.shape
    if tuple(stats.shape) != (S, B, K, 1 + 2 * D) or tuple(center.shape) != (S, B, K, D):
        raise ValueError("stats/center have shapes %s %s" % (tuple(stats.shape), tuple(center.shape)))
    for v in (e, tau_old, mu_old):
        if tuple(v.shape) != (S, B, K, D):
            raise ValueError("draw / old eta shapes differ")
    ps, pk = _post_struct(post)
    rs, rk = _prior_struct(prior)
    tau, mu = torch.empty_like(g), torch.empty_like(g)
    w, o = _outs(S, B, 1, g.device, parts, False)
    with torch.cuda.device(g.device), _call("gmm_logw_eta_stats", 1, 4 * S * B * (K * (1 + 2 * D) + 11 * K * D + 3)):
        check(lib.apgs_gmm_draw_logw_eta_stats_f32(ptr(stats), ptr(center), C.byref(ps), C.byref(rs), ptr(g), ptr(e),
                                                   ptr(tau), ptr(mu), ptr(tau_old), ptr(mu_old), S, B, K, D,
                                                   C.byref(o), _stream()), "apgs_gmm_draw_logw_eta_stats_f32")
    return w, tau, mu


def draw_logw_eta_stats_rng(stats, center, post, prior, seed, tau_old, mu_old, parts=False, old_ancestors=None):
    """draw_logw_eta_stats with the Gamma / Normal variates made inside the kernel (Philox + Marsaglia-Tsang + Box-Muller):
    `seed` is an int64 CUDA tensor of two words drawn from torch's generator (philox_seed()).  Returns (Weights, tau, mu).
    old_ancestors (S,B) int32: center / tau_old / mu_old of particle (s,b) are read from row (old_ancestors[s,b], b)."""
    lib = _lib.load()
    stats, center = _req(stats, "stats", ndim=4), _req(center, "center", ndim=4)
    tau_old, mu_old = _req(tau_old, "tau_old", ndim=4), _req(mu_old, "mu_old", ndim=4)
    seed = _req(seed, "seed", dtype=torch.int64, ndim=1)
    S, B, K, D = tau_old.shape
    if tuple(stats.shape) != (S, B, K, 1 + 2 * D) or tuple(center.shape) != (S, B, K, D) or seed.numel() < 2:
        raise ValueError("stats / center / seed have unexpected shapes")
    ps, pk = _post_struct(post)
    rs, rk = _prior_struct(prior)
    tau, mu = torch.empty_like(tau_old), torch.empty_like(tau_old)
    w, o = _outs(S, B, 1, tau_old.device, parts, False)
    with torch.cuda.device(tau_old.device), _call("gmm_logw_eta_stats", 1, 4 * S * B * (K * (1 + 2 * D) + 9 * K * D + 3)):
        if old_ancestors is not None:
            old_ancestors = _req(old_ancestors, "old_ancestors", dtype=torch.int32, ndim=2)
            if tuple(old_ancestors.shape) != (S, B):
                raise ValueError("old_ancestors must be (S, B) int32")
            check(lib.apgs_gmm_draw_logw_eta_stats_rng_via_f32(ptr(stats), ptr(center), C.byref(ps), C.byref(rs), ptr(seed),
                                                               ptr(tau), ptr(mu), ptr(tau_old), ptr(mu_old), ptr(old_ancestors),
                                                               S, B, K, D, C.byref(o), _stream()),
                  "apgs_gmm_draw_logw_eta_stats_rng_via_f32")
            return w, tau, mu
        check(lib.apgs_gmm_draw_logw_eta_stats_rng_f32(ptr(stats), ptr(center), C.byref(ps), C.byref(rs), ptr(seed), ptr(tau),
                                                       ptr(mu), ptr(tau_old), ptr(mu_old), S, B, K, D, C.byref(o), _stream()),
              "apgs_gmm_draw_logw_eta_stats_rng_f32")
    return w, tau, mu


def philox_seed(device, generator=None):
    """two 64-bit words from torch's generator, on the device (graph-safe: fresh values on every replay)"""
    return torch.randint(-2 ** 62, 2 ** 62, (2,), dtype=torch.int64, device=device, generator=generator)


def logw_z(x, tau, mu, q_probs, prior, z, z_old, parts=False, local=False):
    """G3 (+W2): z-step weights (apgs/gmm/objectives.py:197-227)."""
    if not supported(tau.shape[-2], tau.shape[-1]) & BASE:
        _cuda_only(x, tau, mu, z)
        with torch.no_grad():
            return torch_path.logw_z(x, tau, mu, q_probs, prior, z, z_old, parts, local)
    lib = _lib.load()
    x = _x3(x)
    z, z_old, q_probs = _req(z, "z", ndim=4), _req(z_old, "z_old", ndim=4), _req(q_probs, "q_probs", ndim=4)
    tau, mu = _req(tau, "tau", ndim=4), _req(mu, "mu", ndim=4)
    S, B, N, K = z.shape
    D = x.shape[-1]
    rs, rk = _prior_struct(prior)
    w, o = _outs(S, B, N, z.device, parts, local)
    with torch.cuda.device(z.device), _call("gmm_logw_z", 1, 4 * S * B * (3 * N * K + 2 * K * D + 3 + (2 * N if local else 0))):
        check(lib.apgs_gmm_logw_z_f32(ptr(x), ptr(tau), ptr(mu), ptr(q_probs), C.byref(rs), ptr(z), ptr(z_old),
                                      S, B, N, K, D, C.byref(o), _stream()), "apgs_gmm_logw_z_f32")
    return w


def logw_joint(x, post, prior, tau, mu, tau_old, mu_old, q_probs, z, z_old, parts=False):
    """G4b: joint-step weights (apgs/gmm/objectives.py:112-158)."""
    if not supported(tau.shape[-2], tau.shape[-1]) & BASE:
        _cuda_only(x, tau, mu, z)
        with torch.no_grad():
            return torch_path.logw_joint(x, post, prior, tau, mu, tau_old, mu_old, q_probs, z, z_old, parts)
    lib = _lib.load()
    x = _x3(x)
    z, z_old, q_probs = _req(z, "z", ndim=4), _req(z_old, "z_old", ndim=4), _req(q_probs, "q_probs", ndim=4)
    tau, mu, tau_old, mu_old = [_req(v, "tau/mu", ndim=4) for v in (tau, mu, tau_old, mu_old)]
    S, B, N, K = z.shape
    D = x.shape[-1]
    ps, pk = _post_struct(post)
    rs, rk = _prior_struct(prior)
    w, o = _outs(S, B, N, z.device, parts, False)
    with torch.cuda.device(z.device), _call("gmm_logw_joint", 1, 4 * S * B * (3 * N * K + 8 * K * D + 3)):
        check(lib.apgs_gmm_logw_joint_f32(ptr(x), C.byref(ps), C.byref(rs), ptr(tau), ptr(mu), ptr(tau_old),
                                          ptr(mu_old), ptr(q_probs), ptr(z), ptr(z_old), S, B, N, K, D,
                                          C.byref(o), _stream()), "apgs_gmm_logw_joint_f32")
    return w


def z_loop_probe(ctas, passes, out=None, device=None):
    """Measurement only (apgs_gmm_z_loop_probe_f32): launches the fused z step's main loop alone on synthetic tables --
    `ctas` CTAs of 32 particles, `passes` passes of eight quads of hidden units -- and returns the checksum buffer."""
    lib = _lib.load()
    if out is None:
        out = torch.empty(int(ctas) * 256, dtype=torch.float32, device=device if device is not None else "cuda")
    out = _req(out, "out", ndim=1)
    if out.numel() < int(ctas) * 256:
        raise ValueError("out needs ctas * 256 floats")
    with torch.cuda.device(out.device):
        check(lib.apgs_gmm_z_loop_probe_f32(int(ctas), int(passes), ptr(out), _stream()), "apgs_gmm_z_loop_probe_f32")
    return out


def z_tables(x, mlp, prior):
    """Per-instance tables of the fused z step (apgs_gmm_z_tables_f32): the part of Enc_apg_z that depends only on the data
    and the encoder weights, built once per objective and handed to z_step_fused(seed=..., tables=...).  x is (B,N,D) or
    an (S,B,N,D) view of it; returns a float32 CUDA tensor (B, floats per instance)."""
    lib = _lib.load()
    x = _x3(x)
    B, N, D = x.shape
    W1, b1, W2, b2 = [_req(v, "mlp weight") for v in mlp]
    H = W1.shape[0]
    K = prior.prior_pi.numel()
    if tuple(W1.shape) != (H, 3 * D) or W2.numel() != H or b1.numel() != H or b2.numel() != 1:
        raise ValueError("unexpected Enc_apg_z weight shapes")
    nbytes = lib.apgs_gmm_z_tables_bytes(B, N, H, K, D)
    if nbytes == 0 or nbytes % (4 * B):
        raise ValueError("apgs_gmm_z_tables_bytes rejected the shape")
    tables = torch.empty(B, nbytes // (4 * B), dtype=torch.float32, device=x.device)
    rs, rk = _prior_struct(prior)
    with torch.cuda.device(x.device), _call("gmm_z_tables", 1, nbytes):
        check(lib.apgs_gmm_z_tables_f32(ptr(x), ptr(W1), ptr(b1), ptr(W2), ptr(b2), H, C.byref(rs), B, N, K, D,
                                        ptr(tables), _stream()), "apgs_gmm_z_tables_f32")
    return tables


def z_step_fused(x, tau, mu, mlp, prior, z_old=None, uniforms=None, z_forced=None, want_q=False,
                 parts=False, local=False, also_labels=0, want_onehot=True, eta_heads=None, seed=None, tables=None,
                 src_ancestors=None, src_ancestors_z=None):
    """N1+G3: Enc_apg_z MLP + softmax + categorical draw + z-step weights in one kernel.
    mlp = (W1 (H,3D), b1 (H), W2 (1,H) or (H), b2 (1)).  Returns (z_new, q_probs or None, Weights).
    z_old as uint8 labels (S,B,>=N) selects label mode (z_new comes back as labels of the same row stride).
    also_labels = row stride (>= N): sample-only call that returns ((z_new one-hot, labels), q_probs, Weights);
    want_onehot=False leaves the one-hot form out (None): the draw then costs ~N bytes per particle instead of 4 K N,
    and Weights.density / .log_q_f carry log p(x, z | eta) and log q(z) of the draw.
    eta_heads = (Wg, bg, Wo, bo) of Enc_apg_eta (label mode only): the same launch also emits the NEXT eta step's inputs
    for the labels just drawn -- returns (z_new, None, Weights, (alpha, beta, mu_q, nu), stats) with stats the one-hot
    statistics of z_new centred at `mu` -- bit-identical to eta_posterior_fused(x, z_new, heads, prior, center=mu).
    seed = philox_seed(device) (a one-element int64 device tensor): the uniforms of the draw are made inside the kernel
    (same distribution, Philox stream keyed by the seed) instead of being read from `uniforms`.
    tables = z_tables(x, mlp, prior) (with seed): the per-instance part of the encoder is fetched instead of rebuilt by
    every CTA; same results bit for bit.
    src_ancestors (S,B) int32 (with seed): tau, mu and z_old of particle (s,b) are read from row (src_ancestors[s,b], b) --
    the ancestors of a resampling step whose gather is deferred (ops.gather_ancestors(..., first=...)); same results as
    gathering first.  src_ancestors_z: the map for z_old alone when it differs (default: src_ancestors)."""
    lib = _lib.load()
    if seed is not None:
        if uniforms is not None or z_forced is not None:
            raise ValueError("seed replaces uniforms; it has no meaning with z_forced")
        return _z_step_fused_rng(lib, x, tau, mu, mlp, prior, z_old, want_q, parts, local, also_labels, want_onehot,
                                 eta_heads, seed, tables, src_ancestors, src_ancestors_z)
    if tables is not None or src_ancestors is not None or src_ancestors_z is not None:
        raise ValueError("tables / src_ancestors ride on the seed entry (apgs_gmm_z_step_fused_tab_f32)")
    x = _x3(x)
    tau, mu = _req(tau, "tau", ndim=4), _req(mu, "mu", ndim=4)
    S, B, K, D = tau.shape
    N = x.shape[-2]
    W1, b1, W2, b2 = [_req(v, "mlp weight") for v in mlp]
    H = W1.shape[0]
    if tuple(W1.shape) != (H, 3 * D) or W2.numel() != H or b1.numel() != H or b2.numel() != 1:
        raise ValueError("unexpected Enc_apg_z weight shapes")
    dev = tau.device
    if also_labels:
        if z_old is not None or z_forced is not None:
            raise ValueError("also_labels is a sample-only mode")
        LN = int(also_labels)
        uniforms = torch.rand(S, B, N, device=dev) if uniforms is None else _req(uniforms, "uniforms", ndim=3)
        z_hot = torch.empty(S, B, N, K, dtype=torch.float32, device=dev) if want_onehot else None
        z_lab = torch.empty(S, B, LN, dtype=torch.uint8, device=dev)
        q = torch.empty(S, B, N, K, dtype=torch.float32, device=dev) if want_q else None
        rs, rk = _prior_struct(prior)
        w, o = _outs(S, B, N, dev, False, local)         # local: log_q_local of the draw (log_w_local has no meaning here)
        with torch.cuda.device(dev), _call("gmm_z_step_fused", 1, S * B * (LN + 4 * ((N * K if want_onehot else 0) + N + 2 * K * D + (N * K if want_q else 0)))):
            check(lib.apgs_gmm_z_step_fused_labels_f32(ptr(x), ptr(tau), ptr(mu), ptr(W1), ptr(b1), ptr(W2), ptr(b2), H,
                                                       C.byref(rs), None, ptr(uniforms), S, B, N, K, D,
                                                       ptr(z_lab), LN, ptr(z_hot), ptr(q), C.byref(o), _stream()),
                  "apgs_gmm_z_step_fused_labels_f32")
        return (z_hot, z_lab), q, w
    labels = z_old is not None and z_old.dtype == torch.uint8
    if labels:
        # label mode: assignments as one byte per point, in and out (compact latents inside a sweep)
        if z_forced is not None:
            raise ValueError("label mode samples from uniforms; z_forced needs the one-hot entry")
        z_old = _req(z_old, "z_old labels", dtype=torch.uint8, ndim=3)
        LN = z_old.shape[-1]                              # label rows may be padded (16-byte gather vectors)
        if tuple(z_old.shape[:2]) != (S, B) or LN < N:
            raise ValueError("z_old labels have shape %s, expected (%d,%d,>=%d)" % (tuple(z_old.shape), S, B, N))
        uniforms = torch.rand(S, B, N, device=dev) if uniforms is None else _req(uniforms, "uniforms", ndim=3)
        z_new = torch.empty(S, B, LN, dtype=torch.uint8, device=dev)
        q = torch.empty(S, B, N, K, dtype=torch.float32, device=dev) if want_q else None
        rs, rk = _prior_struct(prior)
        w, o = _outs(S, B, N, dev, parts, local)
        nbytes = S * B * (2 * N + 4 * (N + 2 * K * D + 3 + (N * K if want_q else 0)))
        if eta_heads is not None:
            if want_q or parts or local:
                raise ValueError("the eta statistics ride on the lean launch: no q_probs / parts / per-point outputs")
            Wg, bg, Wo, bo = [_req(v, "head weight") for v in eta_heads]
            if tuple(Wg.shape) != (K, D + K) or tuple(Wo.shape) != (D, D + K):
                raise ValueError("unexpected head weight shapes %s %s" % (tuple(Wg.shape), tuple(Wo.shape)))
            post = [torch.empty(S, B, K, D, dtype=torch.float32, device=dev) for _ in range(4)]
            stats = torch.empty(S, B, K, 1 + 2 * D, dtype=torch.float32, device=dev)
            po = GmmPost(*[v.data_ptr() for v in post])
            with torch.cuda.device(dev), _call("gmm_z_step_fused", 1, nbytes + 4 * S * B * (4 * K * D + K * (1 + 2 * D))):
                check(lib.apgs_gmm_z_step_fused_labels_eta_f32(ptr(x), ptr(tau), ptr(mu), ptr(W1), ptr(b1), ptr(W2), ptr(b2), H,
                                                               C.byref(rs), ptr(z_old), ptr(uniforms), S, B, N, K, D,
                                                               ptr(z_new), LN, ptr(Wg), ptr(bg), ptr(Wo), ptr(bo),
                                                               C.byref(po), ptr(stats), C.byref(o), _stream()),
                      "apgs_gmm_z_step_fused_labels_eta_f32")
            return z_new, None, w, tuple(post), stats
        with torch.cuda.device(dev), _call("gmm_z_step_fused", 1, nbytes):
            check(lib.apgs_gmm_z_step_fused_labels_f32(ptr(x), ptr(tau), ptr(mu), ptr(W1), ptr(b1), ptr(W2), ptr(b2), H,
                                                       C.byref(rs), ptr(z_old), ptr(uniforms), S, B, N, K, D,
                                                       ptr(z_new), LN, None, ptr(q), C.byref(o), _stream()),
                  "apgs_gmm_z_step_fused_labels_f32")
        return z_new, q, w
    if z_old is not None:
        z_old = _req(z_old, "z_old", ndim=4)
    if z_forced is not None:
        z_forced = _req(z_forced, "z_forced", ndim=4)
    elif uniforms is None:
        uniforms = torch.rand(S, B, N, device=dev)
    if uniforms is not None:
        uniforms = _req(uniforms, "uniforms", ndim=3)
    z_new = torch.empty(S, B, N, K, dtype=torch.float32, device=dev)
    q = torch.empty(S, B, N, K, dtype=torch.float32, device=dev) if want_q else None
    rs, rk = _prior_struct(prior)
    w, o = _outs(S, B, N, dev, parts, local)
    nbytes = 4 * S * B * (N * K * (2 if z_old is not None else 1) + N + 2 * K * D + 3 + (N * K if want_q else 0))
    with torch.cuda.device(dev), _call("gmm_z_step_fused", 1, nbytes):
        check(lib.apgs_gmm_z_step_fused_f32(ptr(x), ptr(tau), ptr(mu), ptr(W1), ptr(b1), ptr(W2), ptr(b2), H,
                                            C.byref(rs), ptr(z_old), ptr(z_forced), ptr(uniforms), S, B, N, K, D,
                                            ptr(z_new), ptr(q), C.byref(o), _stream()), "apgs_gmm_z_step_fused_f32")
    return z_new, q, w


def _z_step_fused_rng(lib, x, tau, mu, mlp, prior, z_old, want_q, parts, local, also_labels, want_onehot, eta_heads, seed,
                      tables=None, src_ancestors=None, src_ancestors_z=None):
    """z_step_fused with in-kernel uniforms: the same modes and return values, one entry point."""
    x = _x3(x)
    tau, mu = _req(tau, "tau", ndim=4), _req(mu, "mu", ndim=4)
    S, B, K, D = tau.shape
    N = x.shape[-2]
    W1, b1, W2, b2 = [_req(v, "mlp weight") for v in mlp]
    H = W1.shape[0]
    if tuple(W1.shape) != (H, 3 * D) or W2.numel() != H or b1.numel() != H or b2.numel() != 1:
        raise ValueError("unexpected Enc_apg_z weight shapes")
    dev = tau.device
    seed = _req(seed, "seed", dtype=torch.int64)
    if tables is not None:
        tables = _req(tables, "tables", ndim=2)
        if tables.shape[0] != B or tables.numel() * 4 != lib.apgs_gmm_z_tables_bytes(B, N, H, K, D):
            raise ValueError("tables were built for another shape")
    if src_ancestors is not None:
        src_ancestors = _req(src_ancestors, "src_ancestors", dtype=torch.int32, ndim=2)
        if tuple(src_ancestors.shape) != (S, B):
            raise ValueError("src_ancestors must be (S, B) int32")
    if src_ancestors_z is not None:
        src_ancestors_z = _req(src_ancestors_z, "src_ancestors_z", dtype=torch.int32, ndim=2)
        if tuple(src_ancestors_z.shape) != (S, B):
            raise ValueError("src_ancestors_z must be (S, B) int32")
    rs, rk = _prior_struct(prior)
    z_hot_old = z_lab_old = z_hot = z_lab = post = stats = po = None
    heads = [None] * 4
    LN = N
    if also_labels:
        if z_old is not None:
            raise ValueError("also_labels is a sample-only mode")
        LN = int(also_labels)
        z_hot = torch.empty(S, B, N, K, dtype=torch.float32, device=dev) if want_onehot else None
        z_lab = torch.empty(S, B, LN, dtype=torch.uint8, device=dev)
        w, o = _outs(S, B, N, dev, False, local)
        nbytes = S * B * (LN + 4 * ((N * K if want_onehot else 0) + 2 * K * D + (N * K if want_q else 0)))
    elif z_old is not None and z_old.dtype == torch.uint8:
        z_lab_old = _req(z_old, "z_old labels", dtype=torch.uint8, ndim=3)
        LN = z_lab_old.shape[-1]
        if tuple(z_lab_old.shape[:2]) != (S, B) or LN < N:
            raise ValueError("z_old labels have shape %s, expected (%d,%d,>=%d)" % (tuple(z_lab_old.shape), S, B, N))
        z_lab = torch.empty(S, B, LN, dtype=torch.uint8, device=dev)
        w, o = _outs(S, B, N, dev, parts, local)
        nbytes = S * B * (2 * N + 4 * (2 * K * D + 3 + (N * K if want_q else 0)))
        if eta_heads is not None:
            if want_q or parts or local:
                raise ValueError("the eta statistics ride on the lean launch: no q_probs / parts / per-point outputs")
            heads = [_req(v, "head weight") for v in eta_heads]
            if tuple(heads[0].shape) != (K, D + K) or tuple(heads[2].shape) != (D, D + K):
                raise ValueError("unexpected head weight shapes %s %s" % (tuple(heads[0].shape), tuple(heads[2].shape)))
            post = [torch.empty(S, B, K, D, dtype=torch.float32, device=dev) for _ in range(4)]
            stats = torch.empty(S, B, K, 1 + 2 * D, dtype=torch.float32, device=dev)
            po = GmmPost(*[v.data_ptr() for v in post])
            nbytes += 4 * S * B * (4 * K * D + K * (1 + 2 * D))
    else:
        if eta_heads is not None:
            raise ValueError("the eta statistics need label mode")
        z_hot_old = _req(z_old, "z_old", ndim=4) if z_old is not None else None
        z_hot = torch.empty(S, B, N, K, dtype=torch.float32, device=dev)
        w, o = _outs(S, B, N, dev, parts, local)
        nbytes = 4 * S * B * (N * K * (2 if z_old is not None else 1) + 2 * K * D + 3 + (N * K if want_q else 0))
    q = torch.empty(S, B, N, K, dtype=torch.float32, device=dev) if want_q else None
    with torch.cuda.device(dev), _call("gmm_z_step_fused", 1, nbytes):
        check(lib.apgs_gmm_z_step_fused_tab_f32(ptr(x), ptr(tau), ptr(mu), ptr(W1), ptr(b1), ptr(W2), ptr(b2), H,
                                                C.byref(rs), ptr(z_hot_old), ptr(z_lab_old), ptr(seed), S, B, N, K, D,
                                                ptr(z_hot), ptr(z_lab), LN, ptr(q), ptr(heads[0]), ptr(heads[1]),
                                                ptr(heads[2]), ptr(heads[3]), C.byref(po) if po is not None else None,
                                                ptr(stats), C.byref(o), ptr(tables), ptr(src_ancestors), ptr(src_ancestors_z),
                                                _stream()),
              "apgs_gmm_z_step_fused_tab_f32")
    if also_labels:
        return (z_hot, z_lab), q, w
    z_new = z_lab if z_lab is not None else z_hot
    if post is not None:
        return z_new, None, w, tuple(post), stats
    return z_new, q, w


def eta_tables(x, heads):
    """Per-instance tables of the eta step (apgs_gmm_eta_tables_f32): Enc_apg_eta's head factors over (point, label) and the
    points, built once per objective for eta_posterior_fused(..., tables=...) in label mode.  None when (K, D) is not
    supported by the instance-major statistics kernel."""
    lib = _lib.load()
    x = _x3(x)
    B, N, D = x.shape
    Wg, bg, Wo, bo = [_req(v, "head weight") for v in heads]
    K = Wg.shape[0]
    if tuple(Wg.shape) != (K, D + K) or tuple(Wo.shape) != (D, D + K):
        raise ValueError("unexpected head weight shapes %s %s" % (tuple(Wg.shape), tuple(Wo.shape)))
    nbytes = lib.apgs_gmm_eta_tables_bytes(B, N, K, D)
    if nbytes == 0:
        return None
    tables = torch.empty(B, nbytes // (4 * B), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _call("gmm_eta_tables", 1, nbytes):
        check(lib.apgs_gmm_eta_tables_f32(ptr(x), ptr(Wg), ptr(bg), ptr(Wo), ptr(bo), B, N, K, D, ptr(tables), _stream()),
              "apgs_gmm_eta_tables_f32")
    return tables


def eta_posterior_fused(x, z, heads, prior, center=None, z_ancestors=None, center_ancestors=None, tables=None):
    """N1b+G1: Linear heads of Enc_apg_eta (z given) / Enc_rws_eta (z None) + softmax + NSS posterior.
    heads = (Wg, bg, Wo, bo).  Returns alpha, beta, mu, nu of shape (S,B,K,D) ((1,B,K,D) when z is None).
    With `center` (S,B,K,D; the old mu) the same pass also returns the centred one-hot statistics of z,
    (S,B,K,1+2D), as a fifth tensor -- the input of logw_eta_stats.
    z_ancestors / center_ancestors (S,B) int32 (label mode): the label row / the centre of particle (s,b) are read from
    row (map[s,b], b) -- ancestors of resampling steps whose gathers were deferred."""
    lib = _lib.load()
    x = _x3(x)
    B, N, D = x.shape
    Wg, bg, Wo, bo = [_req(v, "head weight") for v in heads]
    if (z_ancestors is not None or center_ancestors is not None or tables is not None) and not (z is not None and z.dtype == torch.uint8):
        raise ValueError("ancestor maps / tables are taken in label mode only")
    K = Wg.shape[0]
    if z is not None and z.dtype == torch.uint8:
        # label mode: z (S,B,N) uint8
        z = _req(z, "z labels", dtype=torch.uint8, ndim=3)
        S, LN = z.shape[0], z.shape[-1]
        if z.shape[1] != B or LN < N:
            raise ValueError("z labels have shape %s, expected (S,%d,>=%d)" % (tuple(z.shape), B, N))
        if tuple(Wg.shape) != (K, D + K) or tuple(Wo.shape) != (D, D + K):
            raise ValueError("unexpected head weight shapes %s %s" % (tuple(Wg.shape), tuple(Wo.shape)))
        pr = [_req(getattr(prior, "prior_" + n), "prior_" + n) for n in ("alpha", "beta", "mu", "nu")]
        outs = [torch.empty(S, B, K, D, dtype=torch.float32, device=x.device) for _ in range(4)]
        stats = None
        if center is not None:
            center = _req(center, "center", ndim=4)
            if tuple(center.shape) != (S, B, K, D):
                raise ValueError("center has shape %s, expected %s" % (tuple(center.shape), (S, B, K, D)))
            stats = torch.empty(S, B, K, 1 + 2 * D, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device), _call("gmm_eta_posterior_stats", 1, S * B * (N + 4 * (6 * K * D + K * (1 + 2 * D)))):
            maps = [None if m is None else _req(m, "ancestor map", dtype=torch.int32, ndim=2) for m in (z_ancestors, center_ancestors)]
            if any(m is not None and tuple(m.shape) != (S, B) for m in maps):
                raise ValueError("ancestor maps must be (S, B) int32")
            if tables is not None:
                tables = _req(tables, "tables", ndim=2)
                if tables.shape[0] != B or tables.numel() * 4 != lib.apgs_gmm_eta_tables_bytes(B, N, K, D):
                    raise ValueError("tables were built for another shape")
            if maps[0] is not None or maps[1] is not None or tables is not None:
                check(lib.apgs_gmm_eta_posterior_stats_labels_via_f32(ptr(x), ptr(z), ptr(Wg), ptr(bg), ptr(Wo), ptr(bo), S, B, N,
                                                                      K, D, *[ptr(v) for v in pr], *[ptr(v) for v in outs],
                                                                      ptr(center), ptr(stats), LN, ptr(maps[0]), ptr(maps[1]),
                                                                      ptr(tables), _stream()),
                      "apgs_gmm_eta_posterior_stats_labels_via_f32")
            else:
                check(lib.apgs_gmm_eta_posterior_stats_labels_f32(ptr(x), ptr(z), ptr(Wg), ptr(bg), ptr(Wo), ptr(bo), S, B, N, K, D,
                                                                  *[ptr(v) for v in pr], *[ptr(v) for v in outs], ptr(center),
                                                                  ptr(stats), LN, _stream()), "apgs_gmm_eta_posterior_stats_labels_f32")
        return tuple(outs) + ((stats,) if stats is not None else ())
    if z is not None:
        z = _req(z, "z", ndim=4)
        S = z.shape[0]
        if tuple(z.shape) != (S, B, N, K):
            raise ValueError("z has shape %s, expected (S,%d,%d,%d)" % (tuple(z.shape), B, N, K))
    else:
        S = 1
    IN = D + (K if z is not None else 0)
    if tuple(Wg.shape) != (K, IN) or tuple(Wo.shape) != (D, IN):
        raise ValueError("unexpected head weight shapes %s %s" % (tuple(Wg.shape), tuple(Wo.shape)))
    pr = [_req(getattr(prior, "prior_" + n), "prior_" + n) for n in ("alpha", "beta", "mu", "nu")]
    outs = [torch.empty(S, B, K, D, dtype=torch.float32, device=x.device) for _ in range(4)]
    if center is not None:
        if z is None:
            raise ValueError("centred statistics need z")
        center = _req(center, "center", ndim=4)
        if tuple(center.shape) != (S, B, K, D):
            raise ValueError("center has shape %s, expected %s" % (tuple(center.shape), (S, B, K, D)))
        stats = torch.empty(S, B, K, 1 + 2 * D, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device), _call("gmm_eta_posterior_stats", 1, 4 * S * B * (N * K + 6 * K * D + K * (1 + 2 * D))):
            check(lib.apgs_gmm_eta_posterior_stats_f32(ptr(x), ptr(z), ptr(Wg), ptr(bg), ptr(Wo), ptr(bo), S, B, N, K, D,
                                                       *[ptr(v) for v in pr], *[ptr(v) for v in outs], ptr(center),
                                                       ptr(stats), _stream()), "apgs_gmm_eta_posterior_stats_f32")
        return tuple(outs) + (stats,)
    with torch.cuda.device(x.device), _call("gmm_eta_posterior_fused", 1, 4 * S * B * (N * K + 4 * K * D)):
        check(lib.apgs_gmm_eta_posterior_fused_f32(ptr(x), ptr(z), ptr(Wg), ptr(bg), ptr(Wo), ptr(bo), S, B, N, K, D,
                                                   *[ptr(v) for v in pr], *[ptr(v) for v in outs], _stream()),
              "apgs_gmm_eta_posterior_fused_f32")
    return tuple(outs)
