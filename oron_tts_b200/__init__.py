"""oron_tts_b200 — B200-native (sm_100a) implementation of the OronTTS inference hot path.

Host code is Python/PyTorch (device memory, streams); all math on the path runs in the
hand-written CUDA kernels of liboron_b200.so through the C ABI in include/oron_b200.h.
"""

__version__ = "0.1.0"
