from oron_tts_b200.flow import CFM, _lens_to_mask  # noqa: F401
