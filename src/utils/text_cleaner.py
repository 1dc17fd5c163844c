from oron_tts_b200.text import TextCleaner  # noqa: F401
