"""Host-side wrappers: validate tensors, allocate outputs with torch, call the C ABI on the
current CUDA stream.  PyTorch is plumbing here (device memory + streams); all arithmetic is
in apgs_b200/csrc/*.cu.  No CPU path exists: a non-CUDA tensor raises.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Latent, check, ptr

STRATEGIES = {"systematic": _lib.SYSTEMATIC, "multinomial": _lib.MULTINOMIAL}


class _Profile(object):
    """Optional per-call accounting used by bench.py: counts the kernels this package launches and,
    when `events` is on, brackets every C-ABI call with CUDA events on the launching stream."""

    def __init__(self):
        self.launches = 0
        self.events = False
        self.records = []          # (name, start_event, end_event, algorithmic_bytes)
        self.notes = {}            # name -> list of device scalars recorded by callers (read after a synchronize)

    def reset(self, events=False):
        self.launches, self.events, self.records, self.notes = 0, events, [], {}

    def summary(self):
        """name -> dict(calls, ms_total, ms_avg, bytes_per_call, gbs) (call after a synchronize)."""
        out = {}
        for name, e0, e1, nbytes in self.records:
            d = out.setdefault(name, {"calls": 0, "ms_total": 0.0, "bytes": 0})
            d["calls"] += 1
            d["ms_total"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
        for d in out.values():
            d["ms_avg"] = d["ms_total"] / d["calls"]
            d["bytes_per_call"] = d["bytes"] / d["calls"]
            d["gbs"] = d["bytes"] / max(d["ms_total"], 1e-9) / 1e6
        return out


PROFILE = _Profile()


class _call(object):
    __slots__ = ("name", "n", "nbytes", "e0")

    def __init__(self, name, n_launches, nbytes=0):
        self.name, self.n, self.nbytes = name, n_launches, nbytes

    def __enter__(self):
        PROFILE.launches += self.n
        if PROFILE.events:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE.events and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.records.append((self.name, self.e0, e1, self.nbytes))
        return False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, name, dtype=torch.float32, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.ApgsError("%s must be a CUDA tensor: apgs_b200 has no CPU fallback" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    return t.detach().contiguous()


def _workspace(nbytes, device):
    if nbytes == 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------- W1 / W2
def weight_stats(log_w, want_w=True):
    """(S,B) log-weights -> (w or None, ess (B), lse (B)).  One kernel; replaces the softmax +
    1/sum(w^2) pairs listed in include/apgs_b200.h (W1)."""
    lib = _lib.load()
    log_w = _req(log_w, "log_w", ndim=2)
    S, B = log_w.shape
    w = torch.empty_like(log_w) if want_w else None
    ess = torch.empty(B, dtype=torch.float32, device=log_w.device)
    lse = torch.empty(B, dtype=torch.float32, device=log_w.device)
    nws = lib.apgs_ancestors_workspace_bytes(S, B)
    ws = _workspace(nws, log_w.device)
    with torch.cuda.device(log_w.device), _call("weight_stats", 1, 4 * S * B * (2 if want_w else 1)):
        check(lib.apgs_weight_stats_f32(ptr(log_w), S, B, ptr(w), ptr(ess), ptr(lse), ptr(ws), nws, _stream()),
              "apgs_weight_stats_f32")
    return w, ess, lse


def local_weight_stats(log_w_local, want_w=True, want_sum=True):
    """(S,B,N) per-point log-weights -> (w_local or None, log_w (S,B) or None)  (W2)."""
    lib = _lib.load()
    x = _req(log_w_local, "log_w_local", ndim=3)
    S, B, N = x.shape
    w = torch.empty_like(x) if want_w else None
    lw = torch.empty(S, B, dtype=torch.float32, device=x.device) if want_sum else None
    with torch.cuda.device(x.device), _call("local_weight_stats", int(want_w) + int(want_sum),
                                            4 * S * B * N * (int(want_sum) + 2 * int(want_w))):
        check(lib.apgs_local_weight_stats_f32(ptr(x), S, B, N, ptr(w), ptr(lw), _stream()),
              "apgs_local_weight_stats_f32")
    return w, lw


# ------------------------------------------------------------------------------------- R2 / R3
def _uniform_args(strategy, uniforms, S, B, layout=None):
    if strategy == "systematic":
        u = _req(uniforms, "uniforms", torch.float32)
        if u.numel() != B:
            raise ValueError("systematic resampling needs one uniform per instance: %d != B=%d" % (u.numel(), B))
        return u, 0, 0
    if strategy == "multinomial":
        if not uniforms.is_cuda:
            raise _lib.ApgsError("uniforms must be a CUDA tensor")
        if uniforms.dtype != torch.float64:
            raise TypeError("multinomial resampling takes float64 uniforms")
        u = uniforms.detach()
        if layout not in (None, "sb", "bs"):
            raise ValueError("uniform_layout must be 'sb', 'bs' or None")
        if layout is None:
            if S == B and tuple(u.shape) == (S, B):
                raise ValueError("multinomial uniforms of shape (%d,%d) are ambiguous when S == B: pass uniform_layout='sb' "
                                 "((S,B)) or 'bs' ((B,S), the reference's generator order)" % (S, B))
            layout = "sb" if tuple(u.shape) == (S, B) else "bs"
        if layout == "sb" and tuple(u.shape) == (S, B):
            return u, u.stride(0), u.stride(1)
        if layout == "bs" and tuple(u.shape) == (B, S):       # torch.rand(B, S): the reference's generator order
            return u, u.stride(1), u.stride(0)
        raise ValueError("multinomial uniforms must be (S,B) ['sb'] or (B,S) ['bs'], got %s with layout %r" % (tuple(u.shape), layout))
    raise ValueError("unknown resampling strategy %r" % (strategy,))


def sample_ancestors(log_w=None, w_norm=None, strategy="systematic", uniforms=None,
                     cdf_mode=_lib.CDF_AUTO, index_dtype=torch.int64, want_stats=False, want_w=True, uniform_layout=None):
    """Ancestor indices (S,B) from log-weights (or already-normalised weights) and explicit
    uniforms.  Returns anc or (anc, w, ess, lse) when want_stats (log_w input only; w is None unless want_w):
    the softmax the ancestor kernel computes anyway IS the W1 normalisation, so a caller that needs ESS / lse of the
    weights it resamples with gets them from the same launch ("duplicated")."""
    lib = _lib.load()
    src = log_w if log_w is not None else w_norm
    if (log_w is None) == (w_norm is None):
        raise ValueError("give exactly one of log_w / w_norm")
    src = _req(src, "weights", ndim=2)
    S, B = src.shape
    u, us, ub = _uniform_args(strategy, uniforms, S, B, uniform_layout)
    dev = src.device
    anc = torch.empty(S, B, dtype=index_dtype, device=dev)
    a64 = anc if index_dtype == torch.int64 else None
    a32 = anc if index_dtype == torch.int32 else None
    if a64 is None and a32 is None:
        raise TypeError("index_dtype must be int64 or int32")
    w = ess = lse = None
    if want_stats:
        if log_w is None:
            raise ValueError("want_stats needs log_w")
        w = torch.empty_like(src) if want_w else None
        ess = torch.empty(B, dtype=torch.float32, device=dev)
        lse = torch.empty(B, dtype=torch.float32, device=dev)
    nws = lib.apgs_ancestors_workspace_bytes(S, B)
    ws = _workspace(nws, dev)
    with torch.cuda.device(dev), _call("sample_ancestors", 1, S * B * (4 + anc.element_size())):
        check(lib.apgs_sample_ancestors_f32(
            ptr(src) if log_w is not None else None, ptr(src) if w_norm is not None else None,
            S, B, STRATEGIES[strategy], ptr(u), us, ub, cdf_mode, ptr(a64), ptr(a32),
            ptr(w), ptr(ess), ptr(lse), ptr(ws), nws, _stream()), "apgs_sample_ancestors_f32")
    return (anc, w, ess, lse) if want_stats else anc


# ------------------------------------------------------------------------------------------ R4
def _latent_array(latents, S, B):
    if len(latents) > _lib.MAX_LATENTS:
        raise ValueError("at most %d latents per call" % _lib.MAX_LATENTS)
    arr = (Latent * max(1, len(latents)))()
    srcs, outs = [], []
    for i, v in enumerate(latents):
        v = _req(v, "latent %d" % i, dtype=None)
        if v.dim() < 2 or v.shape[0] != S or v.shape[1] != B:
            raise ValueError("latent %d has shape %s, expected (S=%d, B=%d, ...)" % (i, tuple(v.shape), S, B))
        o = torch.empty_like(v)
        inner = v[0, 0].numel() * v.element_size() if v.dim() > 2 else v.element_size()
        arr[i].src, arr[i].dst, arr[i].inner_bytes = v.data_ptr(), o.data_ptr(), inner
        srcs.append(v)
        outs.append(o)
    return arr, srcs, outs


def gather_ancestors(latents, anc, first=None):
    """out_l[s,b,...] = latent_l[anc[s,b], b, ...] for every latent, in ONE launch (R4).
    first: optional list (one entry per latent) of (S,B) int32 ancestor maps or None -- a latent with a map still sits in the
    order before an earlier resampling step whose gather was deferred, and is read through both maps:
    out_l[s,b] = latent_l[first_l[anc[s,b], b], b] (apgs_gather_ancestors_via)."""
    lib = _lib.load()
    if anc.dtype not in (torch.int64, torch.int32):
        raise TypeError("ancestor index must be int64 or int32")
    anc = _req(anc, "ancestral_index", dtype=None, ndim=2)
    S, B = anc.shape
    latents = list(latents)
    arr, srcs, outs = _latent_array(latents, S, B)
    if first is not None and any(f is not None for f in first):
        if anc.dtype != torch.int32 or len(first) != len(latents):
            raise ValueError("composed gathers take int32 ancestors and one `first` entry per latent")
        fl = [None if f is None else _req(f, "first-level ancestors", dtype=torch.int32, ndim=2) for f in first]
        if any(f is not None and tuple(f.shape) != (S, B) for f in fl):
            raise ValueError("first-level ancestor maps must be (S, B)")
        fp = (C.c_void_p * len(latents))(*[None if f is None else f.data_ptr() for f in fl])
        nbytes = S * B * (4 + 2 * sum(int(a.inner_bytes) for a in arr[:len(latents)]) + 4 * sum(f is not None for f in fl))
        with torch.cuda.device(anc.device), _call("gather_ancestors", 1, nbytes):
            check(lib.apgs_gather_ancestors_via(ptr(anc), S, B, arr, fp, len(latents), _stream()), "apgs_gather_ancestors_via")
        return outs
    a32 = anc if anc.dtype == torch.int32 else None
    a64 = anc if anc.dtype == torch.int64 else None
    nbytes = S * B * (anc.element_size() + 2 * sum(int(a.inner_bytes) for a in arr[:len(latents)]))
    with torch.cuda.device(anc.device), _call("gather_ancestors", 1, nbytes):
        check(lib.apgs_gather_ancestors(ptr(a32), ptr(a64), S, B, arr, len(latents), _stream()),
              "apgs_gather_ancestors")
    return outs


def compose_ancestors(prev, nxt):
    """out[s,b] = prev[nxt[s,b], b] (int32): the ancestor map of two consecutive resampling steps for a latent that was not
    gathered in between (apgs_compose_ancestors)."""
    lib = _lib.load()
    prev = _req(prev, "prev", dtype=torch.int32, ndim=2)
    nxt = _req(nxt, "next", dtype=torch.int32, ndim=2)
    if prev.shape != nxt.shape:
        raise ValueError("ancestor maps must have the same (S, B) shape")
    S, B = nxt.shape
    out = torch.empty_like(nxt)
    with torch.cuda.device(nxt.device), _call("compose_ancestors", 1, S * B * 12):
        check(lib.apgs_compose_ancestors(ptr(prev), ptr(nxt), S, B, ptr(out), _stream()), "apgs_compose_ancestors")
    return out


# ------------------------------------------------------------------------------------------ R5
def resample(latents, log_w=None, w_norm=None, strategy="systematic", uniforms=None,
             cdf_mode=_lib.CDF_AUTO, want_index=False, want_stats=False, want_w=True, uniform_layout=None):
    """Fused resample: ancestors (int32, in workspace) + gather of every latent.
    Returns (outs, anc64 or None, (w, ess, lse) or None); w is None unless want_w."""
    lib = _lib.load()
    src = log_w if log_w is not None else w_norm
    if (log_w is None) == (w_norm is None):
        raise ValueError("give exactly one of log_w / w_norm")
    src = _req(src, "weights", ndim=2)
    S, B = src.shape
    dev = src.device
    u, us, ub = _uniform_args(strategy, uniforms, S, B, uniform_layout)
    latents = list(latents)
    arr, srcs, outs = _latent_array(latents, S, B)
    anc64 = torch.empty(S, B, dtype=torch.int64, device=dev) if want_index else None
    w = ess = lse = None
    if want_stats:
        if log_w is None:
            raise ValueError("want_stats needs log_w")
        w = torch.empty_like(src) if want_w else None
        ess = torch.empty(B, dtype=torch.float32, device=dev)
        lse = torch.empty(B, dtype=torch.float32, device=dev)
    nws = lib.apgs_resample_workspace_bytes(S, B)
    ws = _workspace(nws, dev)
    nbytes = S * B * (4 + 2 * sum(int(a.inner_bytes) for a in arr[:len(latents)]))
    with torch.cuda.device(dev), _call("resample", 2, nbytes):
        check(lib.apgs_resample_f32(
            ptr(src) if log_w is not None else None, ptr(src) if w_norm is not None else None,
            S, B, STRATEGIES[strategy], ptr(u), us, ub, cdf_mode, arr, len(latents),
            ptr(anc64), ptr(w), ptr(ess), ptr(lse), ptr(ws), nws, _stream()), "apgs_resample_f32")
    return outs, anc64, ((w, ess, lse) if want_stats else None)
