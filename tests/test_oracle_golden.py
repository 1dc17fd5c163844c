"""Pin the CPU oracle (oracle/) to the live reference: every fixture under tests/golden/ was produced by
importing /root/reference (tests/golden/make_golden.py); the oracle restatement must reproduce it to fp32
round-off. The oracle is then the checker of the CUDA path in tests/test_parity_gpu.py."""

import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import weights as GW  # noqa: E402

from oracle import audio_oracle as AO  # noqa: E402
from oracle import dit_oracle as DO  # noqa: E402
from oron_tts_b200.f5tts import F5TTS  # noqa: E402


def _gold(name):
    return torch.load(os.path.join(HERE, "golden", name), weights_only=False)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def ref_state_dict(name: str) -> dict:
    """Same seeded weights the golden generator loaded into the reference model."""
    keys = _gold("state_keys.pt")[name]
    proto = {}
    for k, shape in keys.items():
        proto[k] = torch.empty(shape)
    # inv_freq is a buffer that keeps its constructor value (modules.py:73)
    sd = GW.fill_state_dict(proto, GW.SEEDS[name])
    mc = GW.CONFIGS[name]["model"]
    dh = mc["dim"] // mc["heads"]
    sd["cfm.backbone.rotary_embed.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, dh, 2).float() / dh))
    return sd


def test_state_dict_layout_matches_reference():
    """Checkpoint contract: identical keys and shapes for every BASELINE config."""
    keys = _gold("state_keys.pt")
    for name in ("tiny", "micro", "small"):
        mine = {k: tuple(v.shape) for k, v in F5TTS.from_config(GW.CONFIGS[name]).state_dict().items()}
        assert mine == keys[name], name
    with torch.device("meta"):
        base = F5TTS.from_config(GW.CONFIGS["base"])
    assert {k: tuple(v.shape) for k, v in base.state_dict().items()} == keys["base"]
    assert len(keys["base"]) == 364 and sum(torch.Size(s).numel() for k, s in keys["base"].items() if "inv_freq" not in k) == 428146788


def test_oracle_dit_forward_and_loss():
    g = _gold("dit_tiny.pt")
    sd = ref_state_dict("tiny")
    T = g["x"].shape[1]
    mask = torch.arange(T)[None, :] < g["lens"][:, None]
    out = DO.dit_forward(sd, g["x"], g["cond"], g["text"], g["time"], mask, cfg_infer=True)
    assert _rel(out, g["fwd_cfg"]) < 2e-5
    out = DO.dit_forward(sd, g["x"], g["cond"], g["text"], g["time"], mask, drop_audio_cond=True)
    assert _rel(out, g["fwd_drop"]) < 2e-5
    out = DO.dit_forward(sd, g["x"][:1], g["cond"][:1], g["text"][:1], torch.tensor(0.5))
    assert _rel(out, g["fwd_nomask_scalar_t"]) < 2e-5


def test_oracle_micro_reference_test_config():
    """The reference's own test configuration (tests/test_checkpoint.py:9-24): 2 heads of 32, dim 64, text_dim 32, ff_mult 2."""
    g = _gold("dit_micro.pt")
    sd = ref_state_dict("micro")
    T = g["x"].shape[1]
    mask = torch.arange(T)[None, :] < g["lens"][:, None]
    out = DO.dit_forward(sd, g["x"], g["cond"], g["text"], g["time"], mask, cfg_infer=True)
    assert _rel(out, g["fwd_cfg"]) < 2e-5
    mel, traj = DO.cfm_sample(sd, g["s_ref"], g["s_ids"], torch.tensor([120]), lens=torch.tensor([40]), steps=3,
                              cfg_strength=2.0, sway_sampling_coef=-1.0, seed=7)
    assert _rel(torch.stack(traj), g["s_traj"]) < 2e-5 and _rel(mel, g["s_mel"]) < 2e-5


def test_oracle_sample_tiny():
    g = _gold("dit_tiny.pt")
    sd = ref_state_dict("tiny")
    mel, traj = DO.cfm_sample(sd, torch.zeros(1, 143, 100), g["s1_ids"], torch.tensor([143]), lens=torch.tensor([0]),
                              steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0)
    assert torch.equal(traj[0], g["s1_traj"][0])  # seeded noise, bit-exact (flow.py:270-283)
    assert _rel(torch.stack(traj), g["s1_traj"]) < 2e-5 and _rel(mel, g["s1_mel"]) < 2e-5
    mel, traj = DO.cfm_sample(sd, g["s2_ref"], g["s2_ids"], torch.tensor([150]), lens=torch.tensor([60]), steps=3,
                              cfg_strength=0.0, sway_sampling_coef=None, seed=5)
    assert _rel(torch.stack(traj), g["s2_traj"]) < 2e-5 and _rel(mel, g["s2_mel"]) < 2e-5
    assert torch.equal(mel[:, :60], g["s2_ref"])  # conditioning region is spliced back (flow.py:304)


def test_oracle_sample_small_config1():
    """BASELINE config 1: Small DiT, 'Сайн байна уу', T=143, 32 NFE, CFG 1.5, sway -1, seed 0."""
    g = _gold("sample_small.pt")
    assert g["T"] == 143 and g["ids"] == [4, 30, 11, 21, 25, 53, 12, 11, 21, 25, 11, 53, 32, 32]
    assert torch.allclose(g["y0"][0, 0, :4], torch.tensor([-1.12584, -1.15236, -0.25058, -0.43388]), atol=1e-5)
    sd = ref_state_dict("small")
    mel, traj = DO.cfm_sample(sd, torch.zeros(1, 143, 100), g["full_ids"], torch.tensor([143]), lens=torch.tensor([0]),
                              steps=32, cfg_strength=1.5, sway_sampling_coef=-1.0, seed=0)
    got = torch.stack([traj[i] for i in g["traj_steps"]])
    assert _rel(got, g["traj"]) < 1e-4 and _rel(mel, g["mel"]) < 1e-4


def test_oracle_logmel_bit_exact():
    g = _gold("mel.pt")
    assert torch.equal(AO.htk_filterbank(), g["fb"])
    gen = torch.Generator().manual_seed(11)
    for case in g["cases"]:
        S = case["n"]
        wav = (torch.rand(S, generator=gen) * 2 - 1) * 0.3
        if S == 48000:
            wav = 0.5 * torch.sin(2 * torch.pi * 220 * torch.arange(S) / 24000)
        mel = AO.log_mel(wav)
        assert mel.shape == case["mel"].shape == (100, 1 + S // 256)
        assert float((mel - case["mel"]).abs().max()) < 1e-4
        if case["norm"] is not None:
            assert torch.equal(AO.peak_normalize(wav * 0.37), case["norm"])


def test_oracle_istft_against_reference_decoder():
    g = _gold("istft.pt")
    h = g["head"]  # [B, T, 1026] interleaved (re, im), normalized=True (src/models/decoder.py:86-102)
    ri = h.view(h.shape[0], h.shape[1], 513, 2)
    spec = torch.complex(ri[..., 0], ri[..., 1]).transpose(1, 2)
    wav = AO.istft_center(spec, normalized=True)
    assert wav.shape == g["wav"].shape
    assert _rel(wav, g["wav"]) < 1e-5


def test_oracle_training_loss_grads_and_optimizer():
    """Gradient oracle (autograd over the oracle's CFM loss) and the clip + AdamW restatement against the fixture
    recorded from the live reference (tests/golden/make_golden_train.py)."""
    g = _gold("train_tiny.pt")
    sd = ref_state_dict("tiny")
    x1 = g["mel"].transpose(1, 2)
    draws = DO.cfm_eval_draws(x1, g["lens"])
    loss, grads = DO.cfm_loss_and_grads(sd, draws, g["text"], g["lens"])
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    assert set(grads) == set(g["grads"])
    for k, ref in g["grads"].items():
        assert _rel(grads[k], ref) < 1e-3 or float((grads[k] - ref).abs().max()) < 1e-7, k
    params = {k: sd[k].clone() for k in grads}
    state: dict = {}
    for step in range(2):
        full = dict(sd)
        full.update(params)
        _, gr = DO.cfm_loss_and_grads(full, draws, g["text"], g["lens"])
        norm = DO.adamw_clip_step(params, gr, state, lr=g[f"lr_step{step}"], step=step + 1)
        if step == 0:
            assert abs(norm - float(g["grad_norm"])) < 1e-4 * norm
    for k, ref in g["params_after_2_steps"].items():
        assert float((params[k] - ref).abs().max()) < 2e-6, k
