"""Vocos boundary pins that do not need the (unreachable) pretrained weights: the parameter tree must be exactly the
upstream `charactr/vocos-mel-24khz` one listed in SURVEY.md section 8(a) row a16 [upstream-memory] -- keys AND shapes -- so
that the real `pytorch_model.bin` loads with strict=True, and `Vocos.from_pretrained` must load such a file from a local
directory and from the Hugging Face cache layout, dropping the `feature_extractor.*` tensors the upstream file carries.
Numerical parity with the pretrained weights stays "unpinned" (DESIGN.md section 2). CPU only."""
import os

import pytest
import torch

from oron_tts_b200.vocos import Vocos


def _upstream_keys():
    dim, inter, n_mels, n_fft, layers = 512, 1536, 100, 1024, 8
    k = {"backbone.embed.weight": (dim, n_mels, 7), "backbone.embed.bias": (dim,),
         "backbone.norm.weight": (dim,), "backbone.norm.bias": (dim,),
         "backbone.final_layer_norm.weight": (dim,), "backbone.final_layer_norm.bias": (dim,),
         "head.out.weight": (n_fft + 2, dim), "head.out.bias": (n_fft + 2,), "head.istft.window": (n_fft,)}
    for i in range(layers):
        p = f"backbone.convnext.{i}."
        k.update({p + "dwconv.weight": (dim, 1, 7), p + "dwconv.bias": (dim,), p + "norm.weight": (dim,), p + "norm.bias": (dim,),
                  p + "pwconv1.weight": (inter, dim), p + "pwconv1.bias": (inter,), p + "pwconv2.weight": (dim, inter),
                  p + "pwconv2.bias": (dim,), p + "gamma": (dim,)})
    return k


def test_state_dict_is_the_upstream_layout():
    sd = Vocos().state_dict()
    want = _upstream_keys()
    assert set(sd) == set(want)
    for key, shape in want.items():
        assert tuple(sd[key].shape) == shape, key
    assert torch.equal(sd["head.istft.window"], torch.hann_window(1024))  # periodic Hann, as torch.istft is given upstream


def _upstream_file(path):
    torch.manual_seed(3)
    sd = {k: torch.randn(s) * 0.02 for k, s in _upstream_keys().items()}
    sd["feature_extractor.mel_spec.spectrogram.window"] = torch.hann_window(1024)
    sd["feature_extractor.mel_spec.mel_scale.fb"] = torch.rand(513, 100)
    torch.save(sd, path)
    return sd


def test_from_pretrained_local_directory(tmp_path):
    sd = _upstream_file(tmp_path / "pytorch_model.bin")
    voc = Vocos.from_pretrained(str(tmp_path))
    assert not voc.training
    got = voc.state_dict()
    assert not any(k.startswith("feature_extractor.") for k in got)
    for k, v in got.items():
        assert torch.equal(v, sd[k]), k


def test_from_pretrained_hf_cache_layout(tmp_path, monkeypatch):
    rev = "f" * 40
    repo = tmp_path / "models--charactr--vocos-mel-24khz"
    os.makedirs(repo / "snapshots" / rev)
    os.makedirs(repo / "refs")
    (repo / "refs" / "main").write_text(rev)
    sd = _upstream_file(repo / "snapshots" / rev / "pytorch_model.bin")
    import huggingface_hub.constants as hc

    monkeypatch.setenv("HF_HUB_CACHE", str(tmp_path))
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    monkeypatch.setattr(hc, "HF_HUB_CACHE", str(tmp_path), raising=False)
    monkeypatch.setattr(hc, "HF_HUB_OFFLINE", True, raising=False)
    voc = Vocos.from_pretrained("charactr/vocos-mel-24khz")
    for k, v in voc.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_wrong_geometry_and_missing_tensor_fail_loudly(tmp_path):
    with pytest.raises(NotImplementedError):
        Vocos(dim=384)
    sd = _upstream_file(tmp_path / "pytorch_model.bin")
    del sd["backbone.convnext.3.gamma"]
    torch.save(sd, tmp_path / "pytorch_model.bin")
    with pytest.raises(RuntimeError):
        Vocos.from_pretrained(str(tmp_path))


def test_decode_needs_cuda():
    with pytest.raises(RuntimeError):
        Vocos().decode(torch.zeros(1, 100, 8))
