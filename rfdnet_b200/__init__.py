"""rfdnet_b200 -- B200 (sm_100a) implementation of RfD-Net's point-cloud hot path.

Importing the package does NOT require a GPU; calling any operator does, and raises if
librfdnet_b200.so has not been built (there is no CPU / PyTorch fallback).
"""
from . import _lib  # noqa: F401

__all__ = ["_ext", "pointnet2_utils", "pointnet2_modules", "detection", "onet", "dropin", "dist"]
__version__ = "0.1.0"
