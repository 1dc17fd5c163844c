"""Import-path shim: the reference's scripts import ``src.models.*`` / ``src.utils.*`` (scripts/infer.py:9-10).
These modules re-export the B200-native implementations from ``oron_tts_b200`` under the same names."""
