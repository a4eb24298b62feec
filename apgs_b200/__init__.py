"""apgs_b200 -- B200-native (sm_100a) implementation of the APG reweight -> resample hot path
of hao-w/apgs, behind the reference's Python call surface.

Public surface (mirrors the reference):
    apgs_b200.resampler.Resampler              <- apgs/resampler.py:7-54
    apgs_b200.gmm / dmm / bmnist               <- apgs/{gmm,dmm,bmnist}/{objectives,models}.py (hot-path functions)
    apgs_b200.affine_transformer.Affine_Transformer <- apgs/bmnist/affine_transformer.py
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
