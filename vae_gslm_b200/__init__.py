"""vae_gslm_b200 — B200-native (sm_100a) implementation of the VAE-GSLM hot path.

Drop-in for the reference's PyTorch module API (``models.speech.lvtr.LVTR`` and the ``modules.*`` it
is built from); the math runs in hand-written CUDA kernels behind the C ABI of ``include/vgslm.h``.
"""
__version__ = "0.1.0"
