from oron_tts_b200.dit import DiT
from oron_tts_b200.f5tts import F5TTS
from oron_tts_b200.flow import CFM

__all__ = ["CFM", "DiT", "F5TTS"]
