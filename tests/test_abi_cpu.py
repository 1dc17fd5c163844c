"""The C-ABI library loads on a CPU-only box and exports every symbol include/oron_b200.h declares;
the ctypes mirror of oron_gemm_desc has the C compiler's size and field offsets."""

import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oron_b200.h")


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    from oron_tts_b200 import _lib

    return _lib


def test_header_symbols_exported(built):
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(oron_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in oron_b200.h but not exported"
    assert declared == set(built.EXPORTED_SYMBOLS)
    assert built.lib().oron_abi_version() == 1


def test_train_header_symbols_exported(built):
    """include/oron_b200_train.h (the training-step entry points) against the library and its ctypes mirror."""
    from oron_tts_b200 import _lib_train

    text = open(os.path.join(ROOT, "include", "oron_b200_train.h")).read()
    declared = set(re.findall(r"^int (oron_[a-z0-9_]+)\s*\(", text, flags=re.M))
    assert declared == set(_lib_train.TRAIN_SYMBOLS)
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in oron_b200_train.h but not exported"
    # every prototype has as many parameters as its ctypes argtypes
    for name, body in re.findall(r"^int (oron_[a-z0-9_]+)\s*\(([^;]*?)\);", text, flags=re.M | re.S):
        assert len(body.split(",")) == len(_lib_train._ARGTYPES[name]), name


def test_gemm_desc_layout_matches_c(built, tmp_path):
    fields = [f[0] for f in built.GemmDesc._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
            'printf("%zu\\n", sizeof(oron_gemm_desc));']
    prog += [f'printf("%zu\\n", offsetof(oron_gemm_desc, {f}));' for f in fields]
    prog += ["return 0;}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(built.GemmDesc)
    for f, off in zip(fields, vals[1:]):
        assert getattr(built.GemmDesc, f).offset == off, f


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from oron_tts_b200.audio import AudioProcessor
    from oron_tts_b200.f5tts import F5TTS

    model = F5TTS.from_config({"model": dict(dim=128, depth=1, heads=2, text_dim=64, conv_layers=1)}).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model.cfm.sample(torch.zeros(1, 60, 100), torch.zeros(1, 60, dtype=torch.long), 60, steps=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.cfm.backbone(torch.zeros(1, 60, 100), torch.zeros(1, 60, 100), torch.zeros(1, 60, dtype=torch.long),
                           torch.tensor([0.5]))
    with pytest.raises(RuntimeError, match="CUDA"):
        AudioProcessor().mel_spectrogram(torch.zeros(24000))
    # the training engine, the fp32 mode and the GPU data path have no CPU fallback either
    from oron_tts_b200.data import GpuBatcher
    from oron_tts_b200.precise import PreciseDiT
    from oron_tts_b200.train import TrainEngine

    with pytest.raises(RuntimeError, match="CUDA"):
        TrainEngine(model)
    with pytest.raises(RuntimeError, match="CUDA"):
        PreciseDiT(model.cfm.backbone)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.train()(torch.zeros(1, 100, 60), torch.zeros(1, 60, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        GpuBatcher(min_duration_s=0.1, device="cpu")([torch.zeros(24000)], ["сайн"])


def test_precise_header_symbols_exported(built):
    """include/oron_b200_precise.h (fp32-mode helpers) against the library and its ctypes mirror."""
    from oron_tts_b200 import precise

    text = open(os.path.join(ROOT, "include", "oron_b200_precise.h")).read()
    declared = set(re.findall(r"^int (oron_[a-z0-9_]+)\s*\(", text, flags=re.M))
    assert declared == set(precise.PRECISE_SYMBOLS)
    lib = ctypes.CDLL(built.LIB_PATH)
    for name, body in re.findall(r"^int (oron_[a-z0-9_]+)\s*\(([^;]*?)\);", text, flags=re.M | re.S):
        assert hasattr(lib, name), name
        assert len(body.split(",")) == len(precise._ARGTYPES[name]), name
