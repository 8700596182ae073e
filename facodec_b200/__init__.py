"""facodec_b200 -- B200-native FAcodec encode -> quantize -> decode hot path (sm_100a CUDA
behind the reference's model.encoder / model.quantizer / model.decoder call surface)."""
from .modules import (Activation1d, CNNLSTM, Codec, CodecStream, Decoder, Encoder, Engine, FApredictors, FAquantizer, Munch,  # noqa: F401
                      Redecoder, ResidualVQ, VoiceConverter, build_model)
from ._lib import FacError  # noqa: F401

__all__ = ["build_model", "Encoder", "FAquantizer", "Decoder", "Redecoder", "Codec", "CodecStream", "VoiceConverter", "ResidualVQ", "Activation1d", "CNNLSTM", "FApredictors", "Engine",
           "Munch", "FacError"]
