"""adaptive_b200 — B200-native (sm_100a) caption-decoder hot path of wzn0828/Adaptive.

Public surface mirrors ``code_src/models/adaptive_attention.py`` of the reference:
``Atten``, ``Sentinel``, ``AdaptiveBlock``, ``Decoder``, ``Encoder2Decoder`` (same
signatures and state_dict keys) on top of the C ABI in ``include/adaptive_b200.h``.
"""
from .modules import Atten, AdaptiveBlock, AttentiveCNN, Decoder, Encoder2Decoder, Sentinel  # noqa: F401
from . import functional  # noqa: F401
from ._lib import LIB_PATH, version  # noqa: F401

__all__ = ["Atten", "Sentinel", "AdaptiveBlock", "Decoder", "Encoder2Decoder", "AttentiveCNN", "functional", "version"]
