from oron_tts_b200.checkpoint import CheckpointManager, adapt_state_dict_to_model, stale_remote_checkpoint_paths  # noqa: F401
