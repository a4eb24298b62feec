"""ctypes binding of the C ABI in include/apgs_b200.h (the drop-in boundary).

There is deliberately NO fallback: if the shared library is missing or a call fails, an
exception is raised.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`apgs_b200/csrc/build.sh`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libapgs_b200.so")

MAX_LATENTS = 8
SYSTEMATIC, MULTINOMIAL = 0, 1
CDF_AUTO, CDF_EXACT, CDF_BLOCKED = 0, 1, 2


class ApgsError(RuntimeError):
    pass


class Latent(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("inner_bytes", C.c_int64)]


class GmmPrior(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("alpha", "beta", "mu", "nu", "pi")]


class GmmPost(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("alpha", "beta", "mu", "nu")]


class GmmOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("log_w", "log_q_f", "density", "log_w_f", "log_w_b",
                                          "log_w_local", "log_q_local")]


class DmmPrior(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("mu_mu", "mu_sigma", "pi", "con1", "con0", "radi", "recon_sigma")]


class DmmMlp(C.Structure):
    """nn.Sequential(Linear, act, Linear): .0.weight, .0.bias, .2.weight, .2.bias (device pointers)."""
    _fields_ = [(n, C.c_void_p) for n in ("W0", "b0", "W2", "b2")]


class DmmOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("log_w", "log_q_f", "density", "ll_f", "log_w_f", "log_w_b",
                                          "log_w_local", "log_q_local", "ll_local")]


_p, _i64, _int, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_size_t
_PR, _PO, _OUT = C.POINTER(GmmPrior), C.POINTER(GmmPost), C.POINTER(GmmOut)
_DPR, _DOUT, _DMLP = C.POINTER(DmmPrior), C.POINTER(DmmOut), C.POINTER(DmmMlp)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "apgs_version": (_int, []),
    "apgs_last_error": (C.c_char_p, []),
    "apgs_gmm_supported": (_int, [_i64, _i64]),
    "apgs_dmm_supported": (_int, [_i64, _i64]),
    "apgs_weight_stats_f32": (_int, [_p, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "apgs_local_weight_stats_f32": (_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "apgs_ancestors_workspace_bytes": (_sz, [_i64, _i64]),
    "apgs_sample_ancestors_f32": (_int, [_p, _p, _i64, _i64, _int, _p, _i64, _i64, _int,
                                         _p, _p, _p, _p, _p, _p, _sz, _p]),
    "apgs_gather_ancestors": (_int, [_p, _p, _i64, _i64, C.POINTER(Latent), _int, _p]),
    "apgs_resample_workspace_bytes": (_sz, [_i64, _i64]),
    "apgs_resample_f32": (_int, [_p, _p, _i64, _i64, _int, _p, _i64, _i64, _int,
                                 C.POINTER(Latent), _int, _p, _p, _p, _p, _p, _sz, _p]),
    "apgs_gmm_nss_posterior_f32": (_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "apgs_gmm_logw_oneshot_f32": (_int, [_p, _PO, _PR, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_gmm_logw_eta_f32": (_int, [_p, _p, _PO, _PR, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_gmm_logw_z_f32": (_int, [_p, _p, _p, _p, _PR, _p, _p, _i64, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_gmm_logw_joint_f32": (_int, [_p, _PO, _PR, _p, _p, _p, _p, _p, _p, _p,
                                       _i64, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_dmm_mu_heads_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "apgs_dmm_nss_pool_f32": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_dmm_nss_pool_onehot_f32": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_dmm_local_encoder_fused_f32": (_int, [_p, _p, _DMLP, _DMLP, _DMLP, _i64, _p, _p, _p] + [_i64] * 5 + [_p] * 7),
    "apgs_dmm_local_encoder_draw_f32": (_int, [_p, _p, _DMLP, _DMLP, _DMLP, _i64, _p, _p, _p, _p] + [_i64] * 5 + [_p] * 7 + [_DMLP, _i64, _p, _p]),
    "apgs_dmm_decoder_hidden2_f32": (_int, [_p, _DMLP, _i64, _i64, _i64, _p, _p]),
    "apgs_dmm_logw_mu_f32": (_int, [_p] * 7 + [_DPR] + [_i64] * 5 + [_DOUT, _p]),
    "apgs_dmm_logw_local_f32": (_int, [_p] * 13 + [_DPR] + [_i64] * 5 + [_DOUT, _p]),
    "apgs_dmm_logw_oneshot_f32": (_int, [_p] * 10 + [_DPR] + [_i64] * 5 + [_DOUT, _p]),
    "apgs_bmnist_crop_f32": (_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_bmnist_crop_bilinear_f32": (_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_bmnist_paste_bilinear_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_bmnist_paste_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_bmnist_paste_ll_f32": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "apgs_bmnist_paste_subtract_f32": (_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _p, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_linear_tc5_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "apgs_linear_tc5_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _p, _p, _sz, _p]),
    "apgs_linear_tf32x3_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "apgs_linear_tf32x3_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _p, _p, _sz, _p]),
    "apgs_bmnist_coor_heads_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "apgs_linear_tc5_packed_bytes": (_sz, [_i64, _i64, _int]),
    "apgs_linear_tc5_pack_f32": (_int, [_p, _i64, _i64, _i64, _int, _p, _p]),
    "apgs_linear_tc5_packed_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "apgs_linear_tc5_packed_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _p, _p, _sz, _p]),
    "apgs_bmnist_template_match_residual_f32": (_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _p, _i64, _i64, _p, _p,
                                                      _i64, _i64, _i64, _i64, _int, _int, _p, _p]),
    "apgs_bmnist_template_match_packed_f32": (_int, [_p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_bmnist_template_match_f32": (_int, [_p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _p]),
    "apgs_bmnist_logw_where_f32": (_int, [_p] * 11 + [_i64] * 4 + [_p] * 6),
    "apgs_bmnist_logw_what_f32": (_int, [_p] * 8 + [_i64] * 5 + [_p] * 5),
    "apgs_gmm_eta_posterior_fused_f32": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64,
                                                _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "apgs_gmm_eta_posterior_stats_f32": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64,
                                                _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "apgs_gmm_eta_posterior_stats_labels_f32": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64,
                                                       _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "apgs_gmm_z_step_fused_labels_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _PR, _p, _p,
                                                _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _p, _OUT, _p]),
    "apgs_gmm_z_step_fused_labels_eta_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _PR, _p, _p,
                                                    _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _p, _p, _p, _PO, _p, _OUT, _p]),
    "apgs_gmm_z_step_fused_rng_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _PR, _p, _p, _p,
                                             _i64, _i64, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _PO, _p, _OUT, _p]),
    "apgs_gmm_z_step_fused_tab_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _PR, _p, _p, _p,
                                             _i64, _i64, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _PO, _p, _OUT, _p, _p, _p, _p]),
    "apgs_compose_ancestors": (_int, [_p, _p, _i64, _i64, _p, _p]),
    "apgs_gmm_eta_posterior_stats_labels_via_f32": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64,
                                                           _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "apgs_gmm_eta_tables_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "apgs_gmm_eta_tables_f32": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_gmm_draw_logw_eta_stats_rng_via_f32": (_int, [_p, _p, _PO, _PR, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_gather_ancestors_via": (_int, [_p, _i64, _i64, C.POINTER(Latent), C.POINTER(C.c_void_p), _int, _p]),
    "apgs_gmm_z_loop_probe_f32": (_int, [_i64, _i64, _p, _p]),
    "apgs_gmm_z_tables_bytes": (_sz, [_i64, _i64, _i64, _i64, _i64]),
    "apgs_gmm_z_tables_f32": (_int, [_p, _p, _p, _p, _p, _i64, _PR, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_gmm_z_logq_backward_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "apgs_gmm_z_logq_backward_f32": (_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _p, _i64, _i64,
                                            _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "apgs_gmm_eta_logq_jac_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "apgs_gmm_label_table_sum_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_gmm_label_table_sum_backward_splits": (_int, [_i64, _i64]),
    "apgs_gmm_label_table_sum_backward_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "apgs_gmm_draw_logw_eta_stats_f32": (_int, [_p, _p, _PO, _PR, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_gmm_draw_logw_eta_stats_rng_f32": (_int, [_p, _p, _PO, _PR, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_gmm_logw_eta_stats_f32": (_int, [_p, _p, _PO, _PR, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _OUT, _p]),
    "apgs_gmm_z_step_fused_f32": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _PR, _p, _p, _p,
                                         _i64, _i64, _i64, _i64, _i64, _p, _p, _OUT, _p]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library handle.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ApgsError(
            "apgs_b200: %s not found -- the CUDA extension has not been built "
            "(run apgs_b200/csrc/build.sh).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().apgs_last_error()
        raise ApgsError("%s failed (%d): %s" % (what, code, msg.decode() if msg else ""))


def ptr(t):
    """device pointer of a tensor, or NULL"""
    return None if t is None else C.c_void_p(t.data_ptr())
