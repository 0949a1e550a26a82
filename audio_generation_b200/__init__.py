"""B200-native residual vector quantization: drop-in for the reference's `som_quantizer` import
(/root/reference/networks/vae.py:6).  Public surface: ResidualQuantizer, tuple_checker."""
from .quantizer import ResidualQuantizer, tuple_checker, approximate_square_root  # noqa: F401

__all__ = ["ResidualQuantizer", "tuple_checker", "approximate_square_root"]
