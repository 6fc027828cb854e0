"""B200-native Student-t entropy bottleneck + GDN (hot path of Dimitrinov74/Domain-Specific-Image-Compression, code/modelv2).

Python host mirroring the reference's model API over a C-ABI CUDA library (libsic.so, include/sic.h).  CUDA only.
"""
from ._lib import SicError, load as load_library  # noqa: F401
from .distributions import FactorizedGaussian, StudentT  # noqa: F401
from .layers import GDN, AnalysisTransform, HyperAnalysis, HyperSynthesis, SynthesisTransform  # noqa: F401
from .model import CompressionModel, rate_distortion_loss  # noqa: F401
from .codec_parallel import compress_sharded, decompress_sharded  # noqa: F401
from . import container  # noqa: F401

__all__ = ["CompressionModel", "rate_distortion_loss", "GDN", "StudentT", "FactorizedGaussian", "AnalysisTransform",
           "SynthesisTransform", "HyperAnalysis", "HyperSynthesis", "load_library", "SicError", "compress_sharded", "decompress_sharded"]
