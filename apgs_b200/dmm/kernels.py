"""Host wrappers of the DMM kernels (apgs_b200/csrc/dmm.cu) -- SURVEY.md D1-D3."""
import ctypes as C

import torch

from .. import _lib
from .._lib import DmmMlp, DmmOut, DmmPrior, check, ptr
from ..ops import _call, _req, _stream


_SUPPORTED = {}


def supported(K, D):
    """apgs_dmm_supported: bit 0 the fused local encoder is compiled for (K, D), bit 1 the decoder kernel for D."""
    key = (int(K), int(D))
    if key not in _SUPPORTED:
        _SUPPORTED[key] = int(_lib.load().apgs_dmm_supported(key[0], key[1]))
    return _SUPPORTED[key]


def _x3(x):
    if x.dim() == 4:
        x = x[0]
    return _req(x, "x", ndim=3)


def _prior(dec):
    """dec: object with prior_mu_mu, prior_mu_sigma (D), prior_pi (K), prior_con1/con0 (1), radi (1), recon_sigma (1)."""
    t = [_req(getattr(dec, n), n) for n in ("prior_mu_mu", "prior_mu_sigma", "prior_pi", "prior_con1", "prior_con0",
                                             "radi", "recon_sigma")]
    return DmmPrior(*[v.data_ptr() for v in t]), t


class Weights(object):
    __slots__ = ("log_w", "log_q_f", "density", "ll_f", "log_w_f", "log_w_b", "log_w_local", "log_q_local", "ll_local")


def _outs(S, B, N, dev, parts, local):
    w, o = Weights(), DmmOut()
    for name in Weights.__slots__:
        t = None
        if name in ("log_w", "log_q_f", "density", "ll_f") or (parts and name in ("log_w_f", "log_w_b")):
            t = torch.empty(S, B, dtype=torch.float32, device=dev)
        elif local and name in ("log_w_local", "log_q_local", "ll_local"):
            t = torch.empty(S, B, N, dtype=torch.float32, device=dev)
        setattr(w, name, t)
        setattr(o, name, None if t is None else t.data_ptr())
    return w, o


def nss_pool(nss1, nss2, prior_mu_mu, prior_mu_sigma):
    """D1 -> (S,B,K,J+2D)."""
    lib = _lib.load()
    nss1, nss2 = _req(nss1, "nss1", ndim=4), _req(nss2, "nss2", ndim=4)
    S, B, N, J = nss1.shape
    K = nss2.shape[-1]
    pm, ps = _req(prior_mu_mu, "prior_mu_mu"), _req(prior_mu_sigma, "prior_mu_sigma")
    D = pm.numel()
    out = torch.empty(S, B, K, J + 2 * D, dtype=torch.float32, device=nss1.device)
    with torch.cuda.device(nss1.device), _call("dmm_nss_pool", 1, 4 * S * B * (N * (J + K) + K * (J + 2 * D))):
        check(lib.apgs_dmm_nss_pool_f32(ptr(nss1), ptr(nss2), ptr(pm), ptr(ps), S, B, N, K, J, D, ptr(out), _stream()),
              "apgs_dmm_nss_pool_f32")
    return out


def mu_heads_supported(mean_seq, sigma_seq):
    """True when (mean_mu, mean_log_sigma) are Linear + Tanh + Linear of equal shapes that apgs_dmm_mu_heads_f32 takes."""
    import torch.nn as nn
    ok = all(len(q) == 3 and isinstance(q[0], nn.Linear) and isinstance(q[1], nn.Tanh) and isinstance(q[2], nn.Linear)
             and q[0].bias is not None and q[2].bias is not None for q in (mean_seq, sigma_seq))
    return (ok and mean_seq[0].weight.shape == sigma_seq[0].weight.shape and mean_seq[2].weight.shape == sigma_seq[2].weight.shape
            and mean_seq[0].in_features <= 32 and mean_seq[2].out_features <= 4 and mean_seq[0].out_features <= 1024)


def mu_heads(v, mean_seq, sigma_seq):
    """(q_mu, q_sigma) = (mean_mu(v), mean_log_sigma(v).exp()) in one launch (apgs_dmm_mu_heads_f32; apgs/dmm/models.py:43-44).
    Forward values only."""
    lib = _lib.load()
    v = _req(v, "nss_prior")
    H = v.shape[-1]
    M = v.numel() // H
    ws = [_req(t, "weight") for t in (mean_seq[0].weight, mean_seq[0].bias, mean_seq[2].weight, mean_seq[2].bias,
                                      sigma_seq[0].weight, sigma_seq[0].bias, sigma_seq[2].weight, sigma_seq[2].bias)]
    H2, Z = ws[0].shape[0], ws[2].shape[0]
    if ws[0].shape[1] != H:
        raise ValueError("mu_heads: input width %d does not match the first layer %s" % (H, tuple(ws[0].shape)))
    mean = torch.empty(tuple(v.shape[:-1]) + (Z,), dtype=torch.float32, device=v.device)
    std = torch.empty_like(mean)
    with torch.cuda.device(v.device), _call("dmm_mu_heads", 1, 4 * M * (H + 2 * Z)):
        check(lib.apgs_dmm_mu_heads_f32(ptr(v), *[ptr(w) for w in ws], M, H, H2, Z, ptr(mean), ptr(std), _stream()),
              "apgs_dmm_mu_heads_f32")
    return mean, std


def _flat(t, name):
    """(S,B,N,1) or (S,B,N) -> contiguous tensor read as (S,B,N)."""
    t = _req(t, name)
    return t


def logw_mu(x, z, hidden2, q_mu, q_sigma, mu, mu_old, dec, parts=False):
    """D2: mu-step weights (apgs/dmm/objectives.py:116-151)."""
    lib = _lib.load()
    x = _x3(x)
    z, hidden2 = _req(z, "z", ndim=4), _req(hidden2, "hidden2", ndim=4)
    q_mu, q_sigma, mu, mu_old = [_req(v, "mu params", ndim=4) for v in (q_mu, q_sigma, mu, mu_old)]
    S, B, N, K = z.shape
    D = x.shape[-1]
    ps, keep = _prior(dec)
    w, o = _outs(S, B, N, z.device, parts, False)
    with torch.cuda.device(z.device), _call("dmm_logw_mu", 1, 4 * S * B * (N * (K + D) + 4 * K * D + 4)):
        check(lib.apgs_dmm_logw_mu_f32(ptr(x), ptr(z), ptr(hidden2), ptr(q_mu), ptr(q_sigma), ptr(mu), ptr(mu_old),
                                       C.byref(ps), S, B, N, K, D, C.byref(o), _stream()), "apgs_dmm_logw_mu_f32")
    return w


def logw_local(x, mu, q_probs, z, beta, con1, con0, hidden2, z_old, beta_old, con1_old, con0_old, hidden2_old, dec,
               parts=False, local=False):
    """D3: local-step weights (apgs/dmm/objectives.py:153-188)."""
    lib = _lib.load()
    x = _x3(x)
    mu = _req(mu, "mu", ndim=4)
    q_probs, z, z_old = [_req(v, "z/q", ndim=4) for v in (q_probs, z, z_old)]
    hidden2, hidden2_old = _req(hidden2, "hidden2", ndim=4), _req(hidden2_old, "hidden2_old", ndim=4)
    b = [_flat(v, "beta/con") for v in (beta, con1, con0, beta_old, con1_old, con0_old)]
    S, B, N, K = z.shape
    D = x.shape[-1]
    ps, keep = _prior(dec)
    w, o = _outs(S, B, N, z.device, parts, local)
    with torch.cuda.device(z.device), _call("dmm_logw_local", 1, 4 * S * B * (N * (3 * K + 2 * D + 6) + K * D + 3)):
        check(lib.apgs_dmm_logw_local_f32(ptr(x), ptr(mu), ptr(q_probs), ptr(z), ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(hidden2),
                                          ptr(z_old), ptr(b[3]), ptr(b[4]), ptr(b[5]), ptr(hidden2_old), C.byref(ps),
                                          S, B, N, K, D, C.byref(o), _stream()), "apgs_dmm_logw_local_f32")
    return w


def logw_oneshot(x, q_mu, q_sigma, mu, q_probs, z, beta, con1, con0, hidden2, dec, local=False):
    """one-shot weights (apgs/dmm/objectives.py:80-114)."""
    lib = _lib.load()
    x = _x3(x)
    q_mu, q_sigma, mu = [_req(v, "mu params", ndim=4) for v in (q_mu, q_sigma, mu)]
    q_probs, z, hidden2 = [_req(v, "z/q/h", ndim=4) for v in (q_probs, z, hidden2)]
    b = [_flat(v, "beta/con") for v in (beta, con1, con0)]
    S, B, N, K = z.shape
    D = x.shape[-1]
    ps, keep = _prior(dec)
    w, o = _outs(S, B, N, z.device, False, local)
    with torch.cuda.device(z.device), _call("dmm_logw_oneshot", 1, 4 * S * B * (N * (2 * K + D + 3) + 3 * K * D + 4)):
        check(lib.apgs_dmm_logw_oneshot_f32(ptr(x), ptr(q_mu), ptr(q_sigma), ptr(mu), ptr(q_probs), ptr(z), ptr(b[0]), ptr(b[1]),
                                            ptr(b[2]), ptr(hidden2), C.byref(ps), S, B, N, K, D, C.byref(o), _stream()),
              "apgs_dmm_logw_oneshot_f32")
    return w


# ------------------------------------------------------------------------------ fused encoders / decoder (N1)
def _mlp_struct(seq, name):
    """seq: nn.Sequential(Linear, act, Linear) -> (DmmMlp, tensors kept alive, H)."""
    t = [_req(v, name) for v in (seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias)]
    return DmmMlp(*[v.data_ptr() for v in t]), t, seq[0].weight.shape[0]


def nss_pool_onehot(tab1, tab2, z, prior_mu_mu, prior_mu_sigma):
    """D1 for one-hot z from per-instance label tables: tab1 (B,N,K,J) = nss1(cat(x_n, e_l)), tab2 (B,N,K,K) =
    softmax(nss2(cat(x_n, e_l))) -> (S,B,K,J+2D)  (apgs/dmm/models.py:129-139)."""
    lib = _lib.load()
    tab1, tab2, z = _req(tab1, "tab1", ndim=4), _req(tab2, "tab2", ndim=4), _req(z, "z", ndim=4)
    S, B, N, K = z.shape
    J = tab1.shape[-1]
    if tuple(tab1.shape) != (B, N, K, J) or tuple(tab2.shape) != (B, N, K, K):
        raise ValueError("label tables have shapes %s %s, expected (%d,%d,%d,J) and (%d,%d,%d,%d)"
                         % (tuple(tab1.shape), tuple(tab2.shape), B, N, K, B, N, K, K))
    pm, ps = _req(prior_mu_mu, "prior_mu_mu"), _req(prior_mu_sigma, "prior_mu_sigma")
    D = pm.numel()
    out = torch.empty(S, B, K, J + 2 * D, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device), _call("dmm_nss_pool_onehot", 1, 4 * (S * B * (N * K + K * (J + 2 * D)) + B * N * K * (J + K))):
        check(lib.apgs_dmm_nss_pool_onehot_f32(ptr(tab1), ptr(tab2), ptr(z), ptr(pm), ptr(ps), S, B, N, K, J, D, ptr(out),
                                               _stream()), "apgs_dmm_nss_pool_onehot_f32")
    return out


def local_encoder_fused(x, mu, enc, uniforms=None, z_forced=None, z_old=None, seed=None, decoder=None):
    """Enc_apg_local in one kernel (apgs/dmm/models.py:76-101).  enc: module with pi_log_prob, angle_log_con1,
    angle_log_con0.  Returns q_probs (S,B,N,K), z (S,B,N,K), con1, con0 (S,B,N,1) and, when z_old is given,
    con1_old, con0_old (S,B,N,1) (else None, None).
    seed (two int64 words on the device, gmm.kernels.philox_seed): the draws are made in the kernel -- the categorical draw's
    uniform (unless `uniforms` is given) and beta ~ Beta(con1, con0) -- and two more values come back: beta (S,B,N,1) and,
    with decoder = Decoder.recon_mu, its unit-normalised output hidden2 (S,B,N,D) (else None)."""
    lib = _lib.load()
    x = _x3(x)
    mu = _req(mu, "mu", ndim=4)
    S, B, K, D = mu.shape
    N = x.shape[1]
    if tuple(x.shape) != (B, N, D):
        raise ValueError("x has shape %s, expected (%d,N,%d)" % (tuple(x.shape), B, D))
    pi, k1, H = _mlp_struct(enc.pi_log_prob, "pi_log_prob")
    c1, k2, H1 = _mlp_struct(enc.angle_log_con1, "angle_log_con1")
    c0, k3, H0 = _mlp_struct(enc.angle_log_con0, "angle_log_con0")
    if not (H == H1 == H0):
        raise ValueError("the three local MLPs must share num_hidden")
    dev = mu.device
    if uniforms is None and z_forced is None and seed is None:
        uniforms = torch.rand(S, B, N, dtype=torch.float32, device=dev)
    for name, t, shp in (("uniforms", uniforms, (S, B, N)), ("z_forced", z_forced, (S, B, N, K)), ("z_old", z_old, (S, B, N, K))):
        if t is not None:
            _req(t, name)
            if tuple(t.shape) != shp:
                raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), shp))
    q = torch.empty(S, B, N, K, dtype=torch.float32, device=dev)
    z = torch.empty(S, B, N, K, dtype=torch.float32, device=dev)
    # con1 / con0 are the two halves of one buffer so that beta_sample() can draw both Gammas in one launch
    cbuf = torch.empty(2, S, B, N, 1, dtype=torch.float32, device=dev)
    con = [cbuf[0], cbuf[1]] + ([torch.empty(S, B, N, 1, dtype=torch.float32, device=dev) for _ in range(2)]
                                if z_old is not None else [])
    co = con + [None, None] if z_old is None else con
    nbytes = 4 * S * B * (N * (2 * K + 1 + len(con) + (K if z_old is not None else 0)) + K * D)
    if seed is not None:
        seed = _req(seed, "seed", dtype=torch.int64)
        if seed.numel() < 2:
            raise ValueError("seed must hold two int64 words")
        beta = torch.empty(S, B, N, 1, dtype=torch.float32, device=dev)
        h2 = dnet = None
        HD = 0
        if decoder is not None:
            dnet, keepd, HD = _mlp_struct(decoder, "recon_mu")
            if decoder[0].weight.shape[1] != 1 or decoder[2].weight.shape[0] != D:
                raise ValueError("decoder must map 1 -> H -> D")
            h2 = torch.empty(S, B, N, D, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _call("dmm_local_encoder_fused", 1, nbytes + 4 * S * B * N * (1 + (D if h2 is not None else 0))):
            check(lib.apgs_dmm_local_encoder_draw_f32(ptr(x), ptr(mu), C.byref(pi), C.byref(c1), C.byref(c0), H, ptr(uniforms),
                                                      ptr(z_forced), ptr(z_old), ptr(seed), S, B, N, K, D, ptr(q), ptr(z),
                                                      ptr(co[0]), ptr(co[1]), ptr(co[2]), ptr(co[3]), ptr(beta),
                                                      C.byref(dnet) if dnet is not None else None, HD, ptr(h2), _stream()),
                  "apgs_dmm_local_encoder_draw_f32")
        return q, z, co[0], co[1], co[2], co[3], beta, h2
    with torch.cuda.device(dev), _call("dmm_local_encoder_fused", 1, nbytes):
        check(lib.apgs_dmm_local_encoder_fused_f32(ptr(x), ptr(mu), C.byref(pi), C.byref(c1), C.byref(c0), H, ptr(uniforms),
                                                   ptr(z_forced), ptr(z_old), S, B, N, K, D, ptr(q), ptr(z), ptr(co[0]),
                                                   ptr(co[1]), ptr(co[2]), ptr(co[3]), _stream()),
              "apgs_dmm_local_encoder_fused_f32")
    return q, z, co[0], co[1], co[2], co[3]


def decoder_hidden2(beta, recon_mu):
    """unit-normalised decoder output (apgs/dmm/models.py:202-203): beta (...,1) -> (...,D)."""
    lib = _lib.load()
    beta = _req(beta, "beta")
    net, keep, H = _mlp_struct(recon_mu, "recon_mu")
    D = recon_mu[2].weight.shape[0]
    if recon_mu[0].weight.shape[1] != 1:
        raise ValueError("recon_mu must map 1 -> H -> D")
    M = beta.numel()
    shape = tuple(beta.shape[:-1]) if beta.shape[-1] == 1 else tuple(beta.shape)
    out = torch.empty(shape + (D,), dtype=torch.float32, device=beta.device)
    with torch.cuda.device(beta.device), _call("dmm_decoder_hidden2", 1, 4 * M * (1 + D)):
        check(lib.apgs_dmm_decoder_hidden2_f32(ptr(beta), C.byref(net), H, D, M, ptr(out), _stream()),
              "apgs_dmm_decoder_hidden2_f32")
    return out


def beta_sample(con1, con0):
    """Beta(con1, con0).sample() as torch.distributions draws it (Dirichlet on [con1, con0]: two standard Gammas,
    normalise, clamp to [tiny, 1 - eps]) without the stack / cat / select copies: when con1 and con0 are adjacent
    halves of one buffer (local_encoder_fused) both Gammas come from ONE launch."""
    n = con1.numel()
    if (con1.is_contiguous() and con0.is_contiguous() and con1.dtype == con0.dtype and con1.shape == con0.shape
            and con1.untyped_storage().data_ptr() == con0.untyped_storage().data_ptr()
            and con0.storage_offset() == con1.storage_offset() + n):
        both = con1.as_strided((2,) + tuple(con1.shape), (n,) + tuple(con1.stride()))
    else:
        both = torch.stack((con1, con0), 0)
    g = torch._standard_gamma(both)
    fi = torch.finfo(g.dtype)
    return (g[0] / (g[0] + g[1])).clamp_(fi.tiny, 1.0 - fi.eps)
