"""Training mode of the GMM sweep on the kernels.

The APG loss is  sum_s w_s.detach() * (-log q_s)  (apgs/gmm/objectives.py:183,220): autograd needs only the gradients of
the FORWARD proposal log-densities w.r.t. the encoder parameters.  Two `torch.autograd.Function`s put the kernels on that
path, so `loss_required` no longer switches the sweep back to PyTorch encoders:

  FusedZStep      the fused z step (csrc/gmm_fused.cu: Enc_apg_z MLP + softmax + draw + weights) with a hand-written
                  backward kernel (csrc/gmm_train.cu: k_gmm_z_logq_bwd) for d sum(G * log q(z)) / d{W1, b1, W2, b2};
  LabelTableSum   the label-indexed table sums behind Enc_apg_eta's neural sufficient statistics and their adjoint
                  (k_label_table_sum / _bwd); the per-instance head table and the O(K D) Normal-Gamma formulas stay
                  ordinary differentiable PyTorch ops on small tensors.
"""
import ctypes as C

import torch

from .. import _lib
from .._lib import check, ptr
from ..ops import _call, _req, _stream
from . import kernels


def z_logq_backward(x, tau, mu, mlp, q_probs, z, grad_logq):
    """(gW1, gb1, gW2, gb2) of sum(grad_logq * log q(z | x, tau, mu)); grad_logq (S,B,N) (any strides incl. 0 over N)."""
    lib = _lib.load()
    x = kernels._x3(x)
    tau, mu, q_probs = _req(tau, "tau", ndim=4), _req(mu, "mu", ndim=4), _req(q_probs, "q_probs", ndim=4)
    W1, b1, W2 = [_req(v, "mlp weight") for v in mlp[:3]]
    S, B, K, D = tau.shape
    N, H = x.shape[-2], W1.shape[0]
    labels = z.dtype == torch.uint8
    z = _req(z, "z", dtype=torch.uint8 if labels else torch.float32, ndim=3 if labels else 4)
    g = grad_logq.detach()
    if not g.is_cuda or g.dtype != torch.float32 or tuple(g.shape) != (S, B, N):
        raise ValueError("grad_logq must be a float32 CUDA tensor of shape (S,B,N)")
    # element (r = s*B + b, n) must sit at r * g_sr + n * g_sn (expanded views -- stride 0 over the points, the one-shot's
    # sum over n -- are read in place)
    if B > 1 and S > 1 and g.stride(0) != B * g.stride(1):
        g = g.contiguous()
    g_sr = g.stride(1) if B > 1 else g.stride(0)
    g_sn = g.stride(2)
    dev = tau.device
    gW1, gb1, gW2, gb2 = torch.empty_like(W1), torch.empty_like(b1), torch.empty(H, dtype=torch.float32, device=dev), \
        torch.empty(1, dtype=torch.float32, device=dev)
    nws = lib.apgs_gmm_z_logq_backward_workspace_bytes(S, B, H, D)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev), _call("gmm_z_logq_backward", 3, 4 * S * B * (N * K + N + 2 * K * D) + S * B * N * (1 if labels else 4 * K)):
        check(lib.apgs_gmm_z_logq_backward_f32(ptr(x), ptr(tau), ptr(mu), ptr(W1), ptr(b1), ptr(W2), H, ptr(q_probs),
                                               None if labels else ptr(z), ptr(z) if labels else None,
                                               z.shape[-1] if labels else 0, ptr(g), g_sr, g_sn, S, B, N, K, D,
                                               ptr(gW1), ptr(gb1), ptr(gW2), ptr(gb2), ptr(ws), nws, _stream()),
              "apgs_gmm_z_logq_backward_f32")
    return gW1, gb1, gW2.view_as(mlp[2]), gb2.view_as(mlp[3])


class FusedZStep(torch.autograd.Function):
    """z ~ q(z | x, tau, mu) and everything the z step needs from ONE forward launch; differentiable w.r.t. the MLP through
    log_q_local only (the weights are detached in the reference too, objectives.py:216-220).

    apply(W1, b1, W2, b2, x, tau, mu, z_old, prior, label_stride, seed=None, tables=None)
        -> z_new, q_probs, log_w, density, log_w_local, log_q_local, z_onehot
    z_old one-hot (S,B,N,K) or uint8 labels (S,B,LN) -> z_new in the same form.  z_old None: sample-only mode of the
    one-shot / joint step (log_w / log_w_local are then meaningless); with label_stride > 0 the draw comes back as labels
    (z_new) AND one-hot (z_onehot, what the one-shot's weight kernel reads), else z_onehot is z_new."""

    @staticmethod
    def forward(ctx, W1, b1, W2, b2, x, tau, mu, z_old, prior, label_stride=0, seed=None, tables=None):
        tables = tables if seed is not None else None       # kernels.z_tables of (x, weights): the forward launch fetches them
        if z_old is None and label_stride:
            (z_hot, z_new), q, w = kernels.z_step_fused(x, tau, mu, (W1, b1, W2, b2), prior, want_q=True, local=True,
                                                        also_labels=label_stride, seed=seed, tables=tables)
        else:
            z_new, q, w = kernels.z_step_fused(x, tau, mu, (W1, b1, W2, b2), prior, z_old=z_old, want_q=True, local=True,
                                               seed=seed, tables=tables)
            z_hot = z_new
        ctx.save_for_backward(W1, b1, W2, b2, x, tau, mu, q, z_new)
        log_w_local = w.log_w_local if w.log_w_local is not None else w.log_q_local.new_zeros(())
        outs = (z_new, q, w.log_w, w.density, log_w_local, w.log_q_local, z_hot if z_hot is not z_new else z_new.new_zeros(()))
        ctx.mark_non_differentiable(outs[0], outs[1], outs[2], outs[3], outs[4], outs[6])
        return outs

    @staticmethod
    def backward(ctx, gz, gq, glw, gd, glwl, glq, gzh):
        W1, b1, W2, b2, x, tau, mu, q, z_new = ctx.saved_tensors
        if glq is None:
            return (None,) * 12
        gW1, gb1, gW2, gb2 = z_logq_backward(x, tau, mu, (W1, b1, W2, b2), q, z_new, glq)
        return gW1, gb1, gW2, gb2, None, None, None, None, None, None, None, None


def label_table_sum(table, z):
    lib = _lib.load()
    table = _req(table, "table", ndim=4)
    B, N, K, Cc = table.shape
    labels = z.dtype == torch.uint8
    z = _req(z, "z", dtype=torch.uint8 if labels else torch.float32, ndim=3 if labels else 4)
    S = z.shape[0]
    out = torch.empty(S, B, Cc, dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device), _call("gmm_label_table_sum", 1, S * B * ((N if labels else 4 * N * K) + 4 * Cc)):
        check(lib.apgs_gmm_label_table_sum_f32(ptr(table), None if labels else ptr(z), ptr(z) if labels else None,
                                               z.shape[-1] if labels else 0, S, B, N, K, Cc, ptr(out), _stream()),
              "apgs_gmm_label_table_sum_f32")
    return out


def label_table_sum_backward(grad_out, z, N, K):
    lib = _lib.load()
    g = _req(grad_out, "grad_out", ndim=3)
    S, B, Cc = g.shape
    labels = z.dtype == torch.uint8
    P = int(lib.apgs_gmm_label_table_sum_backward_splits(S, B))     # partial tables, one per slice of the particles
    gt = torch.empty(P, B, N, K, Cc, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device), _call("gmm_label_table_sum_backward", 1, S * B * ((N if labels else 4 * N * K) + 4 * Cc)):
        check(lib.apgs_gmm_label_table_sum_backward_f32(ptr(g), None if labels else ptr(z), ptr(z) if labels else None,
                                                        z.shape[-1] if labels else 0, S, B, N, K, Cc, ptr(gt), _stream()),
              "apgs_gmm_label_table_sum_backward_f32")
    return gt[0] if P == 1 else gt.sum(0)


class LabelTableSum(torch.autograd.Function):
    """out[s,b,:] = sum_n table[b, n, label(s,b,n), :]  -- differentiable w.r.t. `table` (B,N,K,C)."""

    @staticmethod
    def forward(ctx, table, z):
        ctx.save_for_backward(z)
        ctx.dims = (table.shape[1], table.shape[2])
        return label_table_sum(table, z)

    @staticmethod
    def backward(ctx, grad_out):
        (z,) = ctx.saved_tensors
        return label_table_sum_backward(grad_out.contiguous(), z, *ctx.dims), None


class EtaLogQ(torch.autograd.Function):
    """log q(tau, mu) (S,B) of Enc_apg_eta as a differentiable function of its neural statistics nss (S,B,K+2KD): the
    Normal-Gamma update and the Gamma / Normal log-densities in one kernel that also returns the Jacobian."""

    @staticmethod
    def forward(ctx, nss, tau, mu, prior_alpha, prior_beta, prior_mu, prior_nu):
        lib = _lib.load()
        nss = _req(nss, "nss", ndim=3)
        tau, mu = _req(tau, "tau", ndim=4), _req(mu, "mu", ndim=4)
        S, B, K, D = tau.shape
        if nss.shape[-1] != K + 2 * K * D:
            raise ValueError("nss has %d columns, expected K + 2 K D = %d" % (nss.shape[-1], K + 2 * K * D))
        pr = [_req(v, "prior") for v in (prior_alpha, prior_beta, prior_mu, prior_nu)]
        lq = torch.empty(S, B, dtype=torch.float32, device=tau.device)
        jac = torch.empty_like(nss)
        with torch.cuda.device(tau.device), _call("gmm_eta_logq_jac", 1, 4 * S * B * (2 * nss.shape[-1] + 2 * K * D + 1)):
            check(lib.apgs_gmm_eta_logq_jac_f32(ptr(nss), ptr(tau), ptr(mu), *[ptr(v) for v in pr], S, B, K, D, ptr(lq), ptr(jac),
                                                _stream()), "apgs_gmm_eta_logq_jac_f32")
        ctx.save_for_backward(jac)
        return lq

    @staticmethod
    def backward(ctx, g):
        (jac,) = ctx.saved_tensors
        return g.unsqueeze(-1) * jac, None, None, None, None, None, None
