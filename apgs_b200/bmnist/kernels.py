"""Host wrappers of the bouncing-MNIST kernels (apgs_b200/csrc/bmnist.cu) -- SURVEY.md B1-B4."""
import torch

from .. import _lib
from .._lib import check, ptr
from ..ops import _call, _req, _stream


def _frames_view(frames, S):
    """frames as (S,B,T,FP,FP) with arbitrary (possibly zero) strides over s, b, t and dense pixels."""
    if frames.dim() == 4:
        frames = frames.unsqueeze(0).expand(S, -1, -1, -1, -1)
    if not frames.is_cuda or frames.dtype != torch.float32:
        raise _lib.ApgsError("frames must be a float32 CUDA tensor (no CPU fallback)")
    if frames.stride(-1) != 1 or frames.stride(-2) != frames.shape[-1]:
        frames = frames.contiguous()
    return frames.detach()


def crop(frames, z_where, lin_dp, DP, mode="nearest"):
    """B1: (S,B,T,FP,FP) x (S,B,T,K,2) -> (S,B,T,K,DP,DP).  mode: 'nearest' (the reference's) or 'bilinear'."""
    lib = _lib.load()
    fn, fname = ((lib.apgs_bmnist_crop_f32, "apgs_bmnist_crop_f32") if mode == "nearest" else
                 (lib.apgs_bmnist_crop_bilinear_f32, "apgs_bmnist_crop_bilinear_f32"))
    z = _req(z_where, "z_where", ndim=5)
    S, B, T, K, _ = z.shape
    fr = _frames_view(frames, S)
    FP = fr.shape[-1]
    out = torch.empty(S, B, T, K, DP, DP, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device), _call("bmnist_crop", 1, 8 * S * B * T * K * DP * DP):
        check(fn(ptr(fr), fr.stride(0), fr.stride(1), fr.stride(2), ptr(z), ptr(lin_dp),
                 S, B, T, K, FP, DP, ptr(out), _stream()), fname)
    return out


def paste(digits, z_where, lin_fp, FP, mode="nearest"):
    """B2: (S,B,K,DP,DP) x (S,B,T,K,2) -> (S,B,T,K,FP,FP).  mode: 'nearest' (the reference's) or 'bilinear'."""
    lib = _lib.load()
    fn, fname = ((lib.apgs_bmnist_paste_f32, "apgs_bmnist_paste_f32") if mode == "nearest" else
                 (lib.apgs_bmnist_paste_bilinear_f32, "apgs_bmnist_paste_bilinear_f32"))
    z = _req(z_where, "z_where", ndim=5)
    d = _req(digits, "digit", ndim=5)
    S, B, T, K, _ = z.shape
    DP = d.shape[-1]
    out = torch.empty(S, B, T, K, FP, FP, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device), _call("bmnist_paste", 1, 4 * S * B * (K * DP * DP + T * K * FP * FP)):
        check(fn(ptr(d), ptr(z), ptr(lin_fp), S, B, T, K, FP, DP, ptr(out), _stream()), fname)
    return out


def paste_ll(digits, z_where, frames, lin_fp, want_recon=False):
    """B2+B3: -> ll (S,B,T), recon (S,B,T,FP,FP) or None."""
    lib = _lib.load()
    z = _req(z_where, "z_where", ndim=5)
    d = _req(digits, "digit", ndim=5)
    S, B, T, K, _ = z.shape
    DP = d.shape[-1]
    fr = _frames_view(frames, S)
    FP = fr.shape[-1]
    ll = torch.empty(S, B, T, dtype=torch.float32, device=z.device)
    recon = torch.empty(S, B, T, FP, FP, dtype=torch.float32, device=z.device) if want_recon else None
    uniq = (1 if fr.stride(0) == 0 else S)
    nbytes = 4 * (uniq * B * T * FP * FP + S * B * K * DP * DP + S * B * T * (1 + (FP * FP if want_recon else 0)))
    with torch.cuda.device(z.device), _call("bmnist_paste_ll", 1, nbytes):
        check(lib.apgs_bmnist_paste_ll_f32(ptr(d), ptr(z), ptr(lin_fp), ptr(fr), fr.stride(0), fr.stride(1), fr.stride(2),
                                           S, B, T, K, FP, DP, ptr(ll), ptr(recon), _stream()), "apgs_bmnist_paste_ll_f32")
    return ll, recon


def logw_where(q_mean, q_std, z_new, z_prev, z_old, z_old_prev, dec_coor, ll_f=None, ll_b=None, want_w=True,
               want_b=False):
    """B4a: per-frame z_where log-densities and weight -> (log_w or None, log_q_f, log_p_f), each (S,B);
    with want_b also (log_q_b, log_p_b)."""
    lib = _lib.load()
    q_mean, q_std, z_new = _req(q_mean, "q_mean", ndim=4), _req(q_std, "q_std", ndim=4), _req(z_new, "z_new", ndim=4)
    S, B, K, D = z_new.shape
    opt = [None if v is None else _req(v, "z", ndim=4) for v in (z_prev, z_old, z_old_prev)]
    lls = [None if v is None else _req(v, "ll", ndim=2) for v in (ll_f, ll_b)]
    pr = [_req(getattr(dec_coor, n), n) for n in ("prior_mu0", "prior_Sigma0", "prior_Sigmat")]
    dev = z_new.device
    lw = torch.empty(S, B, dtype=torch.float32, device=dev) if want_w else None
    lq = torch.empty(S, B, dtype=torch.float32, device=dev)
    lp = torch.empty(S, B, dtype=torch.float32, device=dev)
    lqb = torch.empty(S, B, dtype=torch.float32, device=dev) if want_b else None
    lpb = torch.empty(S, B, dtype=torch.float32, device=dev) if want_b else None
    with torch.cuda.device(dev), _call("bmnist_logw_where", 1, 4 * S * B * (6 * K * D + 5)):
        check(lib.apgs_bmnist_logw_where_f32(ptr(q_mean), ptr(q_std), ptr(z_new), *[ptr(v) for v in opt],
                                             *[ptr(v) for v in pr], *[ptr(v) for v in lls], S, B, K, D,
                                             ptr(lw), ptr(lq), ptr(lp), ptr(lqb), ptr(lpb), _stream()),
              "apgs_bmnist_logw_where_f32")
    if want_b:
        return lw, lq, lp, lqb, lpb
    return lw, lq, lp


def logw_what(q_mu, q_std, z_new, z_old, ll_f, ll_b=None, extra_log_p=None, extra_log_q=None):
    """B4b: z_what step / one-shot -> (log_w, log_q_f, density, ll_sum), each (S,B)."""
    lib = _lib.load()
    q_mu, q_std, z_new = _req(q_mu, "q_mu", ndim=4), _req(q_std, "q_std", ndim=4), _req(z_new, "z_what", ndim=4)
    S, B, K, Z = z_new.shape
    z_old = None if z_old is None else _req(z_old, "z_what_old", ndim=4)
    ll_f = _req(ll_f, "ll_f", ndim=3)
    T = ll_f.shape[-1]
    ll_b = None if ll_b is None else _req(ll_b, "ll_b", ndim=3)
    ep = None if extra_log_p is None else _req(extra_log_p, "extra_log_p", ndim=2)
    eq = None if extra_log_q is None else _req(extra_log_q, "extra_log_q", ndim=2)
    dev = z_new.device
    outs = [torch.empty(S, B, dtype=torch.float32, device=dev) for _ in range(4)]
    with torch.cuda.device(dev), _call("bmnist_logw_what", 1, 4 * S * B * (4 * K * Z + 2 * T + 6)):
        check(lib.apgs_bmnist_logw_what_f32(ptr(q_mu), ptr(q_std), ptr(z_new), ptr(z_old), ptr(ll_f), ptr(ll_b), ptr(ep),
                                            ptr(eq), S, B, K, Z, T, *[ptr(v) for v in outs], _stream()),
              "apgs_bmnist_logw_what_f32")
    return tuple(outs)


def _dense_images(t, name):
    """(S,B,H,W) float32 CUDA tensor with dense H x W images and arbitrary non-negative (possibly zero) s / b strides."""
    if not t.is_cuda or t.dtype != torch.float32:
        raise _lib.ApgsError("%s must be a float32 CUDA tensor (no CPU fallback)" % name)
    if t.dim() != 4:
        raise ValueError("%s must be (S,B,H,W), got %s" % (name, tuple(t.shape)))
    if t.stride(-1) != 1 or t.stride(-2) != t.shape[-1] or t.stride(0) < 0 or t.stride(1) < 0:
        t = t.contiguous()
    return t.detach()


def _linear(x, weight, bias, relu, which):
    lib = _lib.load()
    x = _req(x, "x")
    w = _req(weight, "weight", ndim=2)
    b = _req(bias, "bias", ndim=1) if bias is not None else None
    N, K = w.shape
    if x.shape[-1] != K or (b is not None and b.numel() != N):
        raise ValueError("linear: x %s, weight %s, bias %s do not match" % (tuple(x.shape), tuple(w.shape), None if b is None else tuple(b.shape)))
    M = x.numel() // K
    out = torch.empty(tuple(x.shape[:-1]) + (N,), dtype=torch.float32, device=x.device)
    wsb, fn, name = ((lib.apgs_linear_tc5_workspace_bytes, lib.apgs_linear_tc5_f32, "apgs_linear_tc5_f32") if which == "tc5" else
                     (lib.apgs_linear_tf32x3_workspace_bytes, lib.apgs_linear_tf32x3_f32, "apgs_linear_tf32x3_f32"))
    nws = wsb(M, N, K)
    ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device), _call("linear_" + which, 2, 4 * (M * K + N * K + M * N)):
        check(fn(ptr(x), ptr(w), ptr(b), M, K, N, 1 if relu else 0, ptr(out), ptr(ws), nws, _stream()), name)
    return out


def linear_tf32x3(x, weight, bias=None, relu=False):
    """act(x W^T + b) on the tensor cores with fp32-level accuracy (apgs_linear_tf32x3_f32: 3xTF32 error compensation, split-K):
    x (..., K), weight (N, K), bias (N) -> (..., N).  Forward values only (no autograd graph)."""
    return _linear(x, weight, bias, relu, "tf32x3")


def linear_tc5(x, weight, bias=None, relu=False):
    """act(x W^T + b) as 3xTF32 through tcgen05 (apgs_linear_tc5_f32: tcgen05.mma kind::tf32 on shared-memory operand halves,
    accumulator in tensor memory, deterministic split-K) -- the faster form on B200.  Forward values only."""
    return _linear(x, weight, bias, relu, "tc5")


class PackedRows:
    """A (rows, K) fp32 matrix held as the packed TF32 hi / lo operand of apgs_linear_tc5_packed_f32 (include/apgs_b200.h,
    csrc/tf32_pack.cuh): what template_match(..., packed=True) returns and linear_tc5_packed consumes.  `lead` is the leading
    shape of the rows ((S, B) for the template-match map)."""

    def __init__(self, buf, lead, K, which):
        self.buf, self.lead, self.K, self.which = buf, tuple(lead), int(K), int(which)
        self.rows = 1
        for n in self.lead:
            self.rows *= int(n)

    def unpack(self):
        """the matrix as an ordinary (..., K) fp32 tensor (hi + lo: exact to ~2^-22 relative) -- tests and fallbacks"""
        TR = 128 if self.which == 0 else 80
        nch = (self.K + 31) // 32
        t = self.buf.view(-1, nch, 2, TR // 8, 8, 8, 4)                  # (row tile, chunk, half, 8-row group, k group, row, k)
        full = (t[:, :, 0] + t[:, :, 1]).permute(0, 2, 4, 1, 3, 5).reshape(-1, nch * 32)
        return full[:self.rows, :self.K].reshape(self.lead + (self.K,)).contiguous()


def pack_rows(x, which):
    """x (..., K) fp32 -> PackedRows (apgs_linear_tc5_pack_f32); which = 0: activations (128-row tiles), 1: weights (80-row tiles)."""
    lib = _lib.load()
    x = _req(x, "x")
    K = x.shape[-1]
    rows = x.numel() // K
    nbytes = lib.apgs_linear_tc5_packed_bytes(rows, K, which)
    buf = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _call("linear_tc5_pack", 1, 4 * rows * K + nbytes):
        check(lib.apgs_linear_tc5_pack_f32(ptr(x), rows, K, K, which, ptr(buf), _stream()), "apgs_linear_tc5_pack_f32")
    return PackedRows(buf, x.shape[:-1], K, which)


def linear_tc5_packed(xp, wp, bias=None, relu=False):
    """act(x W^T + b) from packed operands (PackedRows which=0 / which=1): apgs_linear_tc5_packed_f32 -- a stage of the GEMM is
    two bulk copies, warp-specialised tcgen05 pipeline; bit-identical to linear_tc5 on the same matrices."""
    lib = _lib.load()
    if not (isinstance(xp, PackedRows) and isinstance(wp, PackedRows)) or xp.which != 0 or wp.which != 1 or xp.K != wp.K:
        raise ValueError("linear_tc5_packed: PackedRows activations (which=0) and weights (which=1) of equal K required")
    b = _req(bias, "bias", ndim=1) if bias is not None else None
    M, N, K = xp.rows, wp.rows, xp.K
    if b is not None and b.numel() != N:
        raise ValueError("linear_tc5_packed: bias %s does not match N = %d" % (tuple(b.shape), N))
    out = torch.empty(xp.lead + (N,), dtype=torch.float32, device=xp.buf.device)
    nws = lib.apgs_linear_tc5_packed_workspace_bytes(M, N, K)
    ws = torch.empty(nws, dtype=torch.uint8, device=xp.buf.device)
    with torch.cuda.device(xp.buf.device), _call("linear_tc5_packed", 2, xp.buf.numel() * 4 + wp.buf.numel() * 4 + 4 * M * N):
        check(lib.apgs_linear_tc5_packed_f32(ptr(xp.buf), ptr(wp.buf), ptr(b), M, K, N, 1 if relu else 0, ptr(out), ptr(ws), nws,
                                             _stream()), "apgs_linear_tc5_packed_f32")
    return out


def coor_heads_weights(w1_mean, b1_mean, w1_std, b1_std):
    """the first layers of Enc_coor's two heads in apgs_bmnist_coor_heads_f32's layout: (w1t [ceil(H/4)][2 H2][4], b1 [2 H2])"""
    w = torch.cat([w1_mean.detach(), w1_std.detach()], 0).float()               # (2 H2, H)
    U, H = w.shape
    K4 = (H + 3) // 4
    if K4 * 4 != H:
        w = torch.nn.functional.pad(w, (0, K4 * 4 - H))
    return w.view(U, K4, 4).permute(1, 0, 2).contiguous(), torch.cat([b1_mean.detach(), b1_std.detach()], 0).float().contiguous()


def coor_heads(hidden, w1t, b1, w2_mean, b2_mean, w2_std, b2_std):
    """(q_mean, q_std) = (tanh(W2m relu(W1m h + b1m) + b2m), exp(W2s relu(W1s h + b1s) + b2s)) in one launch
    (apgs_bmnist_coor_heads_f32; apgs/bmnist/models.py:31-32); w1t, b1 from coor_heads_weights.  Forward values only."""
    lib = _lib.load()
    h = _req(hidden, "hidden")
    H = h.shape[-1]
    M = h.numel() // H
    w2m, b2m, w2s, b2s = (_req(t, n) for t, n in ((w2_mean, "w2_mean"), (b2_mean, "b2_mean"), (w2_std, "w2_std"), (b2_std, "b2_std")))
    Z, H2 = w2m.shape
    if tuple(w1t.shape) != ((H + 3) // 4, 2 * H2, 4) or tuple(b1.shape) != (2 * H2,) or tuple(w2s.shape) != (Z, H2) \
            or b2m.numel() != Z or b2s.numel() != Z or not w1t.is_contiguous():
        raise ValueError("coor_heads: weight shapes do not match hidden %s" % (tuple(h.shape),))
    mean = torch.empty(tuple(h.shape[:-1]) + (Z,), dtype=torch.float32, device=h.device)
    std = torch.empty_like(mean)
    with torch.cuda.device(h.device), _call("bmnist_coor_heads", 1, 4 * (M * H + w1t.numel() + 2 * M * Z)):
        check(lib.apgs_bmnist_coor_heads_f32(ptr(h), ptr(w1t), ptr(b1), ptr(w2m), ptr(b2m), ptr(w2s), ptr(b2s), M, H, H2, Z,
                                             ptr(mean), ptr(std), _stream()), "apgs_bmnist_coor_heads_f32")
    return mean, std


def template_match_supports_packed(FP, DP):
    return FP == 96 and DP == 28


def template_match_supports_residual(FP, DP):
    return FP == 96 and DP == 28


def template_match(frame, template, softmax=True, packed=False, prev=None, lin_fp=None):
    """N3: per-(s,b) cross-correlation of frame (S,B,FP,FP) with template (S,B,DP,DP) -> (S,B,CP*CP), CP = FP-DP+1,
    followed by the softmax over the scores (apgs/bmnist/objectives.py:72-74).  Strided views (frames[:, :, t],
    template[:, :, k], S-expanded frames) are read in place.  packed=True: the softmax map as PackedRows (the operand layout of
    Enc_coor's tensor-core layer; FP = 96, DP = 28).  prev = [(digit (S,B,DP,DP), z_where (S,B,2)), ...] (at most two) with
    lin_fp: the match runs on frame - paste(digit0) - paste(digit1), the residual frame of the sequential proposals (:87-89),
    formed inside the kernel (bit-identical to paste_subtract followed by the match)."""
    lib = _lib.load()
    f, t = _dense_images(frame, "frame"), _dense_images(template, "template")
    S, B, FP, _ = f.shape
    DP = t.shape[-1]
    if tuple(t.shape) != (S, B, DP, DP) or f.shape[-2] != FP or DP > FP:
        raise ValueError("frame %s / template %s shapes do not match" % (tuple(f.shape), tuple(t.shape)))
    CP = FP - DP + 1
    if packed and not softmax:
        raise ValueError("template_match: packed output is the softmax map")
    prev = list(prev) if prev else []
    if len(prev) > 2:
        raise ValueError("template_match: at most two previous digits")
    if packed:                                      # softmax written as the packed activation operand of linear_tc5_packed
        nbytes = lib.apgs_linear_tc5_packed_bytes(S * B, CP * CP, 0)
        out = torch.empty(nbytes // 4, dtype=torch.float32, device=f.device)
    else:
        nbytes = 4 * S * B * CP * CP
        out = torch.empty(S, B, CP * CP, dtype=torch.float32, device=f.device)
    pd, pz, lin = [], [], None
    if prev:                                        # argument checks outside the timed call
        if lin_fp is None:
            raise ValueError("template_match: lin_fp is required with prev")
        lin = _req(lin_fp, "lin_fp", ndim=1)
        for d, z in prev:
            d, z = _dense_images(d, "digit"), _req(z, "z_where", ndim=3)
            if tuple(d.shape) != (S, B, DP, DP) or tuple(z.shape) != (S, B, 2) or lin.numel() != FP:
                raise ValueError("template_match: prev digit %s / z_where %s do not match" % (tuple(d.shape), tuple(z.shape)))
            pd.append(d), pz.append(z)
    with torch.cuda.device(f.device), _call("bmnist_template_match", 1, 4 * S * B * (FP * FP + (1 + len(prev)) * DP * DP) + nbytes):
        if prev:
            d1, z1 = (pd[1], pz[1]) if len(pd) > 1 else (None, None)
            check(lib.apgs_bmnist_template_match_residual_f32(
                ptr(f), f.stride(0), f.stride(1), ptr(t), t.stride(0), t.stride(1),
                ptr(pd[0]), pd[0].stride(0), pd[0].stride(1), ptr(pz[0]),
                ptr(d1), d1.stride(0) if d1 is not None else 0, d1.stride(1) if d1 is not None else 0, ptr(z1),
                ptr(lin), S, B, FP, DP, 1 if softmax else 0, 1 if packed else 0, ptr(out), _stream()),
                "apgs_bmnist_template_match_residual_f32")
        elif packed:
            check(lib.apgs_bmnist_template_match_packed_f32(ptr(f), f.stride(0), f.stride(1), ptr(t), t.stride(0), t.stride(1),
                                                            S, B, FP, DP, ptr(out), _stream()),
                  "apgs_bmnist_template_match_packed_f32")
        else:
            check(lib.apgs_bmnist_template_match_f32(ptr(f), f.stride(0), f.stride(1), ptr(t), t.stride(0), t.stride(1),
                                                     S, B, FP, DP, 1 if softmax else 0, ptr(out), _stream()),
                  "apgs_bmnist_template_match_f32")
    return PackedRows(out, (S, B), CP * CP, 0) if packed else out


def paste_subtract(frame, digit, z_where, lin_fp):
    """frame (S,B,FP,FP) - digit_to_frame(digit (S,B,DP,DP), z_where (S,B,2)) -> (S,B,FP,FP): the frame_left update of
    the sequential template matches (apgs/bmnist/objectives.py:87-89) in one pass.  Strided views are read in place."""
    lib = _lib.load()
    f, d = _dense_images(frame, "frame"), _dense_images(digit, "digit")
    S, B, FP, _ = f.shape
    DP = d.shape[-1]
    z = _req(z_where, "z_where", ndim=3)
    if tuple(d.shape) != (S, B, DP, DP) or tuple(z.shape) != (S, B, 2):
        raise ValueError("frame %s / digit %s / z_where %s shapes do not match" % (tuple(f.shape), tuple(d.shape), tuple(z.shape)))
    if d.stride(1) == 0:
        d = d.contiguous()
    out = torch.empty(S, B, FP, FP, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device), _call("bmnist_paste_subtract", 1, 4 * S * B * (2 * FP * FP + DP * DP + 2)):
        check(lib.apgs_bmnist_paste_subtract_f32(ptr(f), f.stride(0), f.stride(1), ptr(d), d.stride(0), d.stride(1), ptr(z),
                                                 ptr(lin_fp), S, B, FP, DP, ptr(out), _stream()),
              "apgs_bmnist_paste_subtract_f32")
    return out
