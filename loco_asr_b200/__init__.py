"""loco_asr_b200 -- B200-native (sm_100a) SpeechT5 speech-encoder for LoCo-ASR embedding extraction.

Scope: the one data-parallel hot path of keya-dialog/LoCo-ASR, ``model.speecht5.encoder(**audios)``
(speech_text/extract_speecht5_{base,finetuned}_embeddings_slurp.py), as hand-written CUDA kernels behind a
C ABI (include/loco_asr.h) with this thin Python host mirroring the reference's call surface.
"""
from .config import LocoSpeechT5Config  # noqa: F401
from ._lib import LocoError, LIB_PATH, build as build_library  # noqa: F401


def __getattr__(name):
    if name in ("LocoSpeechT5Encoder", "LocoEncoderOutput"):
        from . import encoder
        return getattr(encoder, name)
    raise AttributeError(name)

__version__ = "0.1.0"
