from oron_tts_b200.dit import DiT, InputEmbedding  # noqa: F401
