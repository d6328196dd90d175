"""Swin Transformer backbone (D2SwinTransformer of the reference) on the B200 kernels."""
from .backbone import SwinBackbone  # noqa: F401
