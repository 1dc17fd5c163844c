from oron_tts_b200.checkpoint import CheckpointManager, adapt_state_dict_to_model  # noqa: F401
