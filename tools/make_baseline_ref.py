"""baseline/_ref: an UNMODIFIED copy of the reference's importable sources (src/, scripts/infer.py, scripts/train.py,
configs/) plus a minimal `soundfile` stand-in (the real package is not installed in this image; the reference only uses
it for script-level file I/O). baseline/_ref is git-ignored -- reference sources never enter the history -- but travels
to the GPU box with the gpurun snapshot. Used by: bench.py --impl reference (the reference's own CFM.sample on the host
cores), bench.py's gpu_eager_baseline block (the reference in PyTorch eager on the same B200) and
tests/test_infer_dropin_gpu.py (the unmodified scripts/infer.py against this package).

    python tools/make_baseline_ref.py            # needs /root/reference; a no-op where it is absent
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ORON_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

SOUNDFILE_STUB = '''"""Stand-in for the `soundfile` package (not installed here): 16-bit PCM WAV through the standard library."""
import wave

import numpy as np


def write(file, data, samplerate, subtype=None):
    x = np.asarray(data, dtype=np.float32)
    if x.ndim == 1:
        x = x[:, None]
    pcm = (np.clip(x, -1.0, 1.0) * 32767.0).astype("<i2")
    with wave.open(str(file), "wb") as w:
        w.setnchannels(x.shape[1])
        w.setsampwidth(2)
        w.setframerate(int(samplerate))
        w.writeframes(pcm.tobytes())


def read(file, dtype="float32", always_2d=False):
    with wave.open(str(file), "rb") as w:
        n, ch, sr = w.getnframes(), w.getnchannels(), w.getframerate()
        x = np.frombuffer(w.readframes(n), dtype="<i2").astype(np.float32) / 32767.0
    x = x.reshape(-1, ch)
    if ch == 1 and not always_2d:
        x = x[:, 0]
    return x.astype(dtype), sr


def info(file):
    class _Info:
        pass
    with wave.open(str(file), "rb") as w:
        i = _Info()
        i.frames, i.samplerate, i.channels = w.getnframes(), w.getframerate(), w.getnchannels()
        i.duration = i.frames / float(i.samplerate)
    return i
'''


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF, "src")):
        if verbose:
            print(f"[baseline/_ref] {REF} is absent: keeping whatever baseline/_ref holds", file=sys.stderr)
        return os.path.isdir(os.path.join(DST, "src"))
    os.makedirs(DST, exist_ok=True)
    for sub in ("src", "configs"):
        if os.path.isdir(os.path.join(REF, sub)):
            shutil.copytree(os.path.join(REF, sub), os.path.join(DST, sub), dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DST, "scripts"), exist_ok=True)
    for f in ("infer.py", "train.py"):
        if os.path.exists(os.path.join(REF, "scripts", f)):
            shutil.copy2(os.path.join(REF, "scripts", f), os.path.join(DST, "scripts", f))
    os.makedirs(os.path.join(DST, "_stubs"), exist_ok=True)
    with open(os.path.join(DST, "_stubs", "soundfile.py"), "w") as f:
        f.write(SOUNDFILE_STUB)
    if verbose:
        print(f"[baseline/_ref] copied {REF}/{{src,configs,scripts/infer.py,scripts/train.py}} -> {DST}")
    return True


if __name__ == "__main__":
    make()
