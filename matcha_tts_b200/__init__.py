"""matcha_tts_b200 -- B200-native CFM decoder (Matcha-TTS inference hot path).

Package name note: the task names the package `matcha-tts_b200`; a hyphen is not importable in
Python, so the directory is `matcha_tts_b200`.
"""
from .model import (BASECFM, CFM, Decoder, MatchaTTS, denormalize, fix_len_compatibility, generate_path,
                    sequence_mask)

from .text_encoder import TextEncoder  # noqa: E402
from . import hifigan  # noqa: E402,F401  (HiFi-GAN Generator behind the reference's hifigan.models API)
from . import batching  # noqa: E402,F401  (length bucketing + utterance sharding front end)
from . import checkpoint  # noqa: E402,F401  (Lightning checkpoint load, flat weight file, mel writer)
from .checkpoint import load_lightning_checkpoint  # noqa: E402

__all__ = ["batching", "checkpoint", "hifigan", "load_lightning_checkpoint", "BASECFM", "CFM", "Decoder", "MatchaTTS", "TextEncoder", "denormalize", "fix_len_compatibility", "generate_path",
           "sequence_mask"]
