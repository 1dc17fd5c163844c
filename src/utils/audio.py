from oron_tts_b200.audio import AudioProcessor  # noqa: F401
