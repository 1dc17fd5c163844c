from oron_tts_b200.audio import AudioProcessor
from oron_tts_b200.checkpoint import CheckpointManager
from oron_tts_b200.text import CyrillicTokenizer, NumberNormalizer, TextCleaner

__all__ = ["AudioProcessor", "CheckpointManager", "CyrillicTokenizer", "NumberNormalizer", "TextCleaner"]
