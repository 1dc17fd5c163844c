from oron_tts_b200.text import NumberNormalizer  # noqa: F401
