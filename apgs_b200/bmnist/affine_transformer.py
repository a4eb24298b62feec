"""Drop-in for apgs/bmnist/affine_transformer.py:24-65 on the sm_100a crop / paste kernels.

Same constructor and methods:
    Affine_Transformer(frame_pixels, digit_pixels, CUDA, DEVICE)
    .digit_to_frame(digit[S,B,K,DP,DP], z_where[S,B,T,K,2]) -> [S,B,T,K,FP,FP]
    .frame_to_digit(frames[S,B,T,FP,FP], z_where[S,B,T,K,2]) -> [S,B,T,K,DP,DP]
Nearest sampling with align_corners=True, exactly as the reference; z_where is not modified.
`frames` may also be (B,T,FP,FP) or carry a stride-0 particle axis.
"""
import torch

from .. import _lib
from . import kernels


class Affine_Transformer():
    def __init__(self, frame_pixels, digit_pixels, CUDA=True, DEVICE=None, mode="nearest"):
        """mode: 'nearest' -- what the reference samples with (affine_transformer.py:50-51,63-64; the parity mode) -- or
        'bilinear' (north_star's mode switch) for digit_to_frame / frame_to_digit; the fused likelihood and residual
        entries follow the reference (nearest)."""
        super().__init__()
        if mode not in ("nearest", "bilinear"):
            raise ValueError("mode must be 'nearest' or 'bilinear'")
        self.mode = mode
        if not CUDA:
            raise _lib.ApgsError("apgs_b200 Affine_Transformer runs on CUDA only (no CPU fallback)")
        _lib.load()
        self.digit_pixels = digit_pixels
        self.frame_pixels = frame_pixels
        self.translation_dtof = (self.frame_pixels - self.digit_pixels) / self.digit_pixels
        self.translation_ftod = (self.frame_pixels - self.digit_pixels) / self.frame_pixels
        self.scale_dtof = torch.FloatTensor([[self.frame_pixels / self.digit_pixels, 0], [0, self.frame_pixels / self.digit_pixels]])
        self.scale_ftod = torch.FloatTensor([[self.digit_pixels / self.frame_pixels, 0], [0, self.digit_pixels / self.frame_pixels]])
        dev = torch.device("cuda", torch.cuda.current_device()) if DEVICE is None else torch.device(DEVICE)
        self.device = dev
        # affine_grid's base axes, computed by the same CPU linspace the reference's affine_grid uses
        self.lin_fp = torch.linspace(-1, 1, frame_pixels).to(dev)
        self.lin_dp = torch.linspace(-1, 1, digit_pixels).to(dev)
        self.scale_dtof = self.scale_dtof.to(dev)
        self.scale_ftod = self.scale_ftod.to(dev)

    def digit_to_frame(self, digit, z_where):
        """transfer the digits to the frame
        [digit: S * B * K * DP * DP, z_where: S * B * T * K * 2 ===> frame: S * B * T * K * FP * FP]"""
        return kernels.paste(digit, z_where, self.lin_fp, self.frame_pixels, mode=self.mode)

    def frame_to_digit(self, frames, z_where):
        """transfer the frames to the digits
        [frame: S * B * T * FP * FP, z_where: S * B * T * K * 2 ===> digit: S * B * T * K * DP * DP]"""
        return kernels.crop(frames, z_where, self.lin_dp, self.digit_pixels, mode=self.mode)

    def subtract_digit(self, frame, digit, z_where):
        """frame (S,B,FP,FP) - digit_to_frame(digit (S,B,DP,DP), z_where (S,B,2)): the residual-frame update of the
        sequential template matches (apgs/bmnist/objectives.py:87-89) -- the reference has no such entry."""
        return kernels.paste_subtract(frame, digit, z_where, self.lin_fp)

    # fused tail of Dec_digit.forward (apgs/bmnist/models.py:148-154) -- the reference has no such entry
    def reconstruct_ll(self, digit, z_where, frames, want_recon=False):
        return kernels.paste_ll(digit, z_where, frames, self.lin_fp, want_recon=want_recon)

    # ---- differentiable twins for training mode: same pixels as the kernels, PyTorch ops only.
    # The transformer's theta is diagonal, so the nearest-neighbour sampling grid is separable: one source row per
    # output row and one source column per output column (the index formula of csrc/bmnist.cu stn_index); the paste
    # is two gathers, whose autograd is the scatter-add grid_sample(nearest) gives w.r.t. its input.
    @staticmethod
    def _axis_index(lin, scale, t, n_in):
        u = ((lin * scale + t.unsqueeze(-1)) + 1) * ((n_in - 1) / 2)
        idx = torch.round(u)
        ok = (idx >= 0) & (idx <= n_in - 1)
        return idx.clamp(0, n_in - 1).long(), ok

    def digit_to_frame_torch(self, digit, z_where):
        """digit_to_frame with an autograd graph w.r.t. `digit` (z_where carries none: nearest sampling)."""
        FP, DP = self.frame_pixels, self.digit_pixels
        T = z_where.shape[2]
        scale = float(torch.tensor(FP / DP, dtype=torch.float32))
        z = z_where.detach()
        col, cok = self._axis_index(self.lin_fp, scale, -1 * (z[..., 0] * self.translation_dtof), DP)
        row, rok = self._axis_index(self.lin_fp, scale, z[..., 1] * self.translation_dtof, DP)
        img = digit.unsqueeze(2).expand(-1, -1, T, -1, -1, -1)
        rows = torch.gather(img, -2, row.unsqueeze(-1).expand(row.shape + (DP,)))
        out = torch.gather(rows, -1, col.unsqueeze(-2).expand(col.shape[:-1] + (FP, FP)))
        return out * (rok.unsqueeze(-1) & cok.unsqueeze(-2)).to(out.dtype)

    def reconstruct_ll_torch(self, digit, z_where, frames, EPS=1e-9):
        """ll (S,B,T) of apgs/bmnist/models.py:148-162 with an autograd graph w.r.t. `digit`."""
        S = z_where.shape[0]
        if frames.dim() == 4:
            frames = frames.unsqueeze(0).expand(S, -1, -1, -1, -1)
        recon = torch.clamp(self.digit_to_frame_torch(digit, z_where).sum(-3), min=0.0, max=1.0)
        return (torch.log(recon + EPS) * frames + torch.log(1 - recon + EPS) * (1 - frames)).sum(-1).sum(-1)
