"""dc_vic_b200 -- B200-native (sm_100a) implementation of DC-VIC's data-parallel hot path:
the VQGAN codebook quantizer and the CompressAI-style rate/entropy model, behind the
reference's own module interfaces.  Hand-written CUDA behind a C ABI (include/dcvic_b200.h);
PyTorch is plumbing (device memory, streams, autograd glue, torch.distributed)."""
from .quantize import (VectorQuantizer, VectorQuantizer2, codebook_lookup, onehot_feature, swap_quantizer,
                       decode_tokens, CrossEntropyLoss, FocalCrossEntropyLoss)
from . import tiling, bitstream, rans, parallel
from .entropy_models import (EntropyModel, EntropyBottleneck, GaussianConditional, LowerBound,
                             DcvicEntropyBottleneck, SteEntropyBottleneck, GaussianScaleConditional,
                             GaussianMeanScaleConditional, SteGaussianMeanScaleConditional, ste_round,
                             get_scale_table, pmf_to_quantized_cdf, likelihood_to_bit, batch_bits,
                             gaussian_rate_dual, gaussian_codec_step)
from .register import install_compressai_shim, register_entropy_models, ENTROPY_MODEL_CLASSES

__version__ = "0.1.0"
