"""B200-native fused ray renderer for EvenNICER-SLAM (hot path only).

Sub-modules are imported lazily so that the numpy-only helpers (``synthetic``)
work without torch / the CUDA library.
"""
__version__ = "0.1.0"
