from oron_tts_b200.f5tts import (  # noqa: F401
    F5TTS,
    _concat_with_pause,
    _find_split_index,
    _normalise_synthesis_text,
    _stretch_text_to_len,
    split_text_for_synthesis,
)
