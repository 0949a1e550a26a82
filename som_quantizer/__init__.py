"""Drop-in module for `from som_quantizer import ResidualQuantizer, tuple_checker`
(/root/reference/networks/vae.py:6): put the repo root on sys.path and the reference's
vae.py / training.py run unchanged on the B200 kernels."""
from audio_generation_b200.quantizer import ResidualQuantizer, tuple_checker  # noqa: F401

__all__ = ["ResidualQuantizer", "tuple_checker"]
