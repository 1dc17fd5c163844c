from oron_tts_b200.text import SPECIAL_TOKENS, CyrillicTokenizer, validate_language  # noqa: F401
