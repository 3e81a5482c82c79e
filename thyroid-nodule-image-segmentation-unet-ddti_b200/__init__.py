"""B200-native (sm_100a) UNet segmentation hot path.

Drop-in replacements for the reference's ``models/model.py`` (UNet), ``models/loss.py`` (losses) backed by
hand-written CUDA kernels in ``csrc/`` behind the C ABI of ``include/b2s.h``. There is no CPU fallback: the
compute entry points raise if ``lib/libb2s.so`` is missing.
"""
__all__ = ["_lib", "ops", "engine", "models"]
