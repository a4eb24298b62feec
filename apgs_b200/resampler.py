"""Drop-in for apgs/resampler.py:7-54 on top of the sm_100a kernels.

Same constructor, same method names and argument meaning:
    Resampler(strategy, sample_size, CUDA, device)
    .sample_ancestral_index(log_weights: f32[S,B]) -> i64[S,B]      (apgs/resampler.py:22-44)
    .resample_4dims(var, ancestral_index) / .resample_5dims(...)     (apgs/resampler.py:46-54)
plus one fused entry the reference lacks:
    .resample(latents, log_weights) -> list of resampled latents     (= resample_variables)

Random numbers stay in PyTorch (north_star): systematic draws torch.rand(B, 1) -- exactly what
Uniform(0,1).sample((B,)) with shape-[1] parameters consumes (quirk Q5); multinomial draws
float64 uniforms and inverts the CDF, which is what ATen's CPU multinomial does internally.
"""
import torch

from . import _lib, ops


class _GatherFn(torch.autograd.Function):
    """out[s,b,...] = var[anc[s,b],b,...] through k_gather, differentiable w.r.t. `var` like the reference's
    torch.gather (apgs/resampler.py:46-54): the backward scatters the incoming gradient onto the ancestors' rows."""

    @staticmethod
    def forward(ctx, var, anc):
        ctx.save_for_backward(anc)
        return ops.gather_ancestors([var], anc)[0]

    @staticmethod
    def backward(ctx, grad):
        (anc,) = ctx.saved_tensors
        idx = anc.long().reshape(anc.shape + (1,) * (grad.dim() - 2)).expand_as(grad)
        return torch.zeros_like(grad).scatter_add_(0, idx, grad.contiguous()), None


def _gather(var, ancestral_index):
    if var.requires_grad and torch.is_grad_enabled():
        return _GatherFn.apply(var, ancestral_index)
    return ops.gather_ancestors([var], ancestral_index)[0]


class Resampler():
    def __init__(self, strategy, sample_size, CUDA=True, device=None, generator=None):
        if strategy not in ("systematic", "multinomial"):
            # the reference's assert at apgs/resampler.py:11 is always true; its else-branch prints
            raise ValueError("ERROR! specify resampling strategy as either systematic or multinomial.")
        if not CUDA:
            raise _lib.ApgsError("apgs_b200.Resampler runs on CUDA only (no CPU fallback)")
        _lib.load()
        self.strategy = strategy
        self.S = int(sample_size)
        self.CUDA = CUDA
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.generator = generator
        self.cdf_mode = _lib.CDF_AUTO

    # ------------------------------------------------------------------ uniforms (host RNG)
    def draw_uniforms(self, S, B, device):
        """systematic: (B,1) float32, one uniform per instance (quirk Q5); multinomial: (B,S) float64 -- torch.rand(B, S) is
        the order in which ATen's multinomial consumes its generator, i.e. the reference's stream (layout 'bs')."""
        if self.strategy == "systematic":
            return torch.rand(B, 1, device=device, generator=self.generator)
        return torch.rand(B, S, dtype=torch.float64, device=device, generator=self.generator)

    def _layout(self, uniforms, S, B, uniform_layout):
        if uniform_layout is not None or self.strategy == "systematic":
            return uniform_layout
        return "bs" if tuple(uniforms.shape) == (B, S) else "sb"        # (B,S) is what draw_uniforms makes

    def _check(self, log_weights):
        if log_weights.dim() != 2:
            raise ValueError("log_weights must be (S, B)")
        if self.strategy == "systematic" and log_weights.shape[0] != self.S:
            raise ValueError("log_weights has %d particles but the resampler was built for sample_size=%d"
                             % (log_weights.shape[0], self.S))

    # ------------------------------------------------------------------ reference surface
    def sample_ancestral_index(self, log_weights, uniforms=None, uniform_layout=None):
        """sample ancestral indices"""
        self._check(log_weights)
        S, B = log_weights.shape
        if uniforms is None:
            uniforms = self.draw_uniforms(S, B, log_weights.device)
        return ops.sample_ancestors(log_w=log_weights, strategy=self.strategy, uniforms=uniforms,
                                    cdf_mode=self.cdf_mode, uniform_layout=self._layout(uniforms, S, B, uniform_layout))

    def ancestors32(self, log_weights, want_stats=False, want_w=True, uniforms=None, uniform_layout=None):
        """sample_ancestral_index as the kernels keep it (int32), optionally with (w, ess, lse) of the same softmax: the
        building block of a deferred gather (ops.gather_ancestors(..., first=...)).  Returns anc or (anc, (w, ess, lse))."""
        self._check(log_weights)
        S, B = log_weights.shape
        if uniforms is None:
            uniforms = self.draw_uniforms(S, B, log_weights.device)
        r = ops.sample_ancestors(log_w=log_weights, strategy=self.strategy, uniforms=uniforms, cdf_mode=self.cdf_mode,
                                 index_dtype=torch.int32, want_stats=want_stats, want_w=want_w,
                                 uniform_layout=self._layout(uniforms, S, B, uniform_layout))
        if ops.PROFILE.events:                                 # bench.py: fraction of source rows actually read
            a = r[0] if want_stats else r
            srt = a if self.strategy == "systematic" else a.sort(0).values
            ops.PROFILE.notes.setdefault("distinct_ancestor_frac", []).append(
                ((srt[1:] != srt[:-1]).sum() + B).float() / float(S * B))
        return (r[0], tuple(r[1:])) if want_stats else r

    def resample_4dims(self, var, ancestral_index):
        if var.dim() != 4:
            raise ValueError("resample_4dims expects a 4-D tensor")
        return _gather(var, ancestral_index)

    def resample_5dims(self, var, ancestral_index):
        if var.dim() != 5:
            raise ValueError("resample_5dims expects a 5-D tensor")
        return _gather(var, ancestral_index)

    # ------------------------------------------------------------------ fused entry
    def resample(self, latents, log_weights, uniforms=None, want_stats=False, want_index=False, want_w=True,
                 uniform_layout=None):
        """One ancestor draw + gather of every latent (2 launches, no int64 index tensor).
        Returns the list of resampled latents; with want_stats / want_index: (latents, ancestors or None,
        (w, ess, lse) or None) -- the normalised weights, ESS and logsumexp of `log_weights` come out of the ancestor
        kernel's own softmax, so the W1 bookkeeping of a sweep costs no extra launch (w is None unless want_w)."""
        self._check(log_weights)
        S, B = log_weights.shape
        if uniforms is None:
            uniforms = self.draw_uniforms(S, B, log_weights.device)
        lay = self._layout(uniforms, S, B, uniform_layout)
        if ops.PROFILE.events and not want_index:
            # bench.py roofline accounting: same two kernels, timed separately
            r = ops.sample_ancestors(log_w=log_weights, strategy=self.strategy, uniforms=uniforms,
                                     cdf_mode=self.cdf_mode, index_dtype=torch.int32, want_stats=want_stats, want_w=want_w,
                                     uniform_layout=lay)
            anc32, stats = (r[0], tuple(r[1:])) if want_stats else (r, None)
            # fraction of source rows that are actually read (distinct ancestors): the rest is served by L2
            srt = anc32 if self.strategy == "systematic" else anc32.sort(0).values
            distinct = ((srt[1:] != srt[:-1]).sum() + B).float() / float(S * B)
            ops.PROFILE.notes.setdefault("distinct_ancestor_frac", []).append(distinct)
            outs = ops.gather_ancestors(latents, anc32)
            return (outs, None, stats) if want_stats else outs
        outs, anc, stats = ops.resample(latents, log_w=log_weights, strategy=self.strategy,
                                        uniforms=uniforms, cdf_mode=self.cdf_mode,
                                        want_index=want_index, want_stats=want_stats, want_w=want_w, uniform_layout=lay)
        if want_stats or want_index:
            return outs, anc, stats
        return outs
