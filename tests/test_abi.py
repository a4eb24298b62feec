"""CPU: the C-ABI library loads and exports every symbol include/apgs_b200.h declares
(no compute calls -- there is no GPU here), and argument validation happens before launch."""
import ctypes as C
import os
import re

import pytest

from apgs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    out = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            out += re.findall(r"APGS_API\s+[\w\s\*]+?\b(apgs_\w+)\s*\(", txt)
    return out


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 9
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), "symbol %s missing from %s" % (n, _lib.LIB_PATH)
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    for n in _lib.SIGNATURES:
        assert n in names, "%s bound but not declared in include/*.h" % n


def test_version_and_pure_host_queries():
    lib = _lib.load()
    assert lib.apgs_version() == 2
    assert lib.apgs_ancestors_workspace_bytes(1000, 1024) == 0            # CDF tile lives in smem
    need = lib.apgs_ancestors_workspace_bytes(100000, 10)                 # transposed CDF + per-chunk partials
    assert 100000 * 10 * 4 <= need <= 100000 * 10 * 4 + (1 << 16) and need % 256 == 0
    assert lib.apgs_ancestors_workspace_bytes(8192, 7) == 0
    assert lib.apgs_resample_workspace_bytes(1000, 1024) >= 1000 * 1024 * 4


def test_bad_arguments_are_rejected_before_launch():
    lib = _lib.load()
    # shape / null checks return a negative code and set the thread-local message; nothing is launched
    assert lib.apgs_weight_stats_f32(None, 0, 4, None, None, None, None, 0, None) < 0
    assert b"positive" in lib.apgs_last_error()
    assert lib.apgs_weight_stats_f32(None, 4, 4, None, None, None, None, 0, None) < 0
    assert lib.apgs_sample_ancestors_f32(None, None, 4, 4, 0, None, 0, 0, 0, None, None, None, None, None, None, 0, None) < 0
    assert lib.apgs_sample_ancestors_f32(C.c_void_p(16), None, 4, 4, 7, C.c_void_p(16), 0, 0, 0, C.c_void_p(16), None,
                                         None, None, None, None, 0, None) < 0
    assert b"strategy" in lib.apgs_last_error()
    assert lib.apgs_gather_ancestors(None, None, 4, 4, None, 0, None) < 0
    assert lib.apgs_resample_f32(C.c_void_p(16), None, 4, 4, 0, C.c_void_p(16), 0, 0, 0, None, 0, None, None, None, None,
                                 None, 0, None) == -4     # APGS_E_WORKSPACE


def test_no_cpu_fallback():
    import torch
    from apgs_b200 import ops
    from apgs_b200.resampler import Resampler
    with pytest.raises(_lib.ApgsError):
        ops.weight_stats(torch.zeros(4, 4))
    with pytest.raises(_lib.ApgsError):
        Resampler("systematic", 4, False, None)


def test_bad_arguments_of_the_fused_entry_points_are_rejected_before_launch():
    """every check below happens on the host, before any CUDA call: negative APGS_E_* code + a message."""
    lib = _lib.load()
    P = C.c_void_p(64)                       # a non-NULL dummy; never dereferenced on these paths
    pr = _lib.GmmPrior(64, 64, 64, 64, 64)
    out = _lib.GmmOut()
    mlp = _lib.DmmMlp(64, 64, 64, 64)
    # fused z step: uniforms or z_forced required; label variant needs its output and a stride >= N
    assert lib.apgs_gmm_z_step_fused_f32(P, P, P, P, P, P, P, 32, C.byref(pr), None, None, None, 4, 4, 60, 3, 2, P, None,
                                         C.byref(out), None) < 0
    assert b"uniforms" in lib.apgs_last_error()
    assert lib.apgs_gmm_z_step_fused_f32(P, P, P, P, P, P, P, 0, C.byref(pr), None, None, P, 4, 4, 60, 3, 2, P, None,
                                         C.byref(out), None) < 0
    assert lib.apgs_gmm_z_step_fused_f32(P, P, P, P, P, P, P, 32, C.byref(pr), None, None, P, 4, 4, 60, 7, 5, P, None,
                                         C.byref(out), None) == -5            # APGS_E_UNSUPPORTED (K, D) not instantiated
    assert lib.apgs_gmm_z_step_fused_labels_f32(P, P, P, P, P, P, P, 32, C.byref(pr), P, P, 4, 4, 60, 3, 2, None, 64, None, None,
                                                C.byref(out), None) < 0
    assert lib.apgs_gmm_z_step_fused_labels_f32(P, P, P, P, P, P, P, 32, C.byref(pr), P, P, 4, 4, 60, 3, 2, P, 48, None, None,
                                                C.byref(out), None) < 0
    assert b"label_stride" in lib.apgs_last_error()
    # statistics route of the eta step
    assert lib.apgs_gmm_eta_posterior_stats_f32(P, P, P, P, P, P, 4, 4, 60, 3, 2, P, P, P, P, P, P, P, P, None, P, None) < 0
    assert lib.apgs_gmm_eta_posterior_stats_labels_f32(P, P, P, P, P, P, 4, 4, 60, 3, 2, P, P, P, P, P, P, P, P, P, None, 64, None) < 0
    assert lib.apgs_gmm_eta_posterior_stats_labels_f32(P, P, P, P, P, P, 4, 4, 60, 3, 2, P, P, P, P, P, P, P, P, P, P, 10, None) < 0
    post = _lib.GmmPost(64, 64, 64, 64)
    assert lib.apgs_gmm_logw_eta_stats_f32(None, P, C.byref(post), C.byref(pr), P, P, P, P, 4, 4, 3, 2, C.byref(out), None) < 0
    assert lib.apgs_gmm_draw_logw_eta_stats_f32(P, P, C.byref(post), C.byref(pr), None, P, P, P, P, P, 4, 4, 3, 2, C.byref(out), None) < 0
    # DMM fused encoders
    assert lib.apgs_dmm_nss_pool_onehot_f32(P, P, None, P, P, 4, 4, 200, 4, 8, 2, P, None) < 0
    assert lib.apgs_dmm_local_encoder_fused_f32(P, P, C.byref(mlp), C.byref(mlp), C.byref(mlp), 32, None, None, None,
                                                4, 4, 200, 4, 2, P, P, P, P, None, None, None) < 0
    assert b"uniforms" in lib.apgs_last_error()
    assert lib.apgs_dmm_local_encoder_fused_f32(P, P, C.byref(mlp), C.byref(mlp), C.byref(mlp), 32, P, None, P,
                                                4, 4, 200, 4, 2, P, P, P, P, None, None, None) < 0     # z_old needs its outputs
    assert lib.apgs_dmm_decoder_hidden2_f32(P, C.byref(mlp), 32, 2, 0, P, None) < 0
    # bmnist
    assert lib.apgs_bmnist_template_match_f32(P, 0, 9216, P, 784, 784, 4, 4, 28, 96, 1, P, None) < 0      # DP > FP
    assert lib.apgs_bmnist_template_match_f32(P, -1, 9216, P, 784, 784, 4, 4, 96, 28, 1, P, None) < 0
    assert lib.apgs_bmnist_paste_subtract_f32(P, 0, 9216, P, 784, 0, P, P, 4, 4, 96, 28, P, None) < 0


def test_supported_shape_queries_need_no_gpu():
    """apgs_gmm_supported / apgs_dmm_supported are pure host functions: the compiled (K, D) list the host layer consults
    before it routes a shape to the kernels or to the generic PyTorch path."""
    from apgs_b200 import _lib
    lib = _lib.load()
    assert lib.apgs_gmm_supported(3, 2) == 1 | 2 | 4 | 8          # reweighting, fused z step, fused eta heads, label mode
    assert lib.apgs_gmm_supported(5, 2) == 1 | 2 | 4               # K D = 10 > 8: no label mode
    assert lib.apgs_gmm_supported(4, 3) == 1 | 4                   # no fused z step
    assert lib.apgs_gmm_supported(6, 2) == 0 and lib.apgs_gmm_supported(3, 4) == 0
    assert lib.apgs_dmm_supported(4, 2) == 3 and lib.apgs_dmm_supported(7, 2) == 2 and lib.apgs_dmm_supported(4, 5) == 0
    assert lib.apgs_gmm_label_table_sum_backward_splits(1000, 1024) >= 1
