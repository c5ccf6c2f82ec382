"""Mirrors models/backbones/__init__.py of the reference for the modules that are on the B200 path."""
from .swin_backbone import SwinTransformer3D as VQABackbone  # noqa: F401
from .swin_backbone import SwinTransformer3D, swin_3d_small, swin_3d_tiny  # noqa: F401

__all__ = ["VQABackbone", "SwinTransformer3D", "swin_3d_tiny", "swin_3d_small"]
