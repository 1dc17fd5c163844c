"""Parity of the CUDA path with the reference, through the public Python API (DiT.forward, CFM.sample,
AudioProcessor.mel_spectrogram, Vocos.decode), on the same seeded weights / text / noise:

  * against the golden fixtures generated from the live reference (tests/golden/*.pt), and
  * against the CPU oracle (oracle/), itself pinned to those fixtures by tests/test_oracle_golden.py.

Tolerances are the north-star ones: per-NFE-step velocity <= 2e-2 relative L2 (bf16 tensor-core
arithmetic, fp32 state), final mel <= 1e-2 relative L2; frame counts / ids exact (CPU tests).
"""

import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import weights as GW  # noqa: E402

from oracle import audio_oracle as AO  # noqa: E402
from oracle import dit_oracle as DO  # noqa: E402
from oron_tts_b200.audio import AudioProcessor  # noqa: E402
from oron_tts_b200.f5tts import F5TTS  # noqa: E402
from oron_tts_b200.vocos import Vocos  # noqa: E402

VEL_TOL = 2e-2
MEL_TOL = 1e-2
DEV = "cuda"


def _gold(name):
    return torch.load(os.path.join(HERE, "golden", name), weights_only=False)


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def ref_state_dict(name):
    keys = _gold("state_keys.pt")[name]
    sd = GW.fill_state_dict({k: torch.empty(s) for k, s in keys.items()}, GW.SEEDS[name])
    mc = GW.CONFIGS[name]["model"]
    dh = mc["dim"] // mc["heads"]
    sd["cfm.backbone.rotary_embed.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, dh, 2).float() / dh))
    return sd


_MODELS = {}


def model_for(name):
    if name not in _MODELS:
        m = F5TTS.from_config(GW.CONFIGS[name])
        m.load_state_dict(ref_state_dict(name), strict=True)
        _MODELS[name] = m.to(DEV).eval()
    return _MODELS[name]


def sway(steps, coef):
    t = torch.linspace(0, 1, steps + 1)
    return t if coef is None else t + coef * (torch.cos(torch.pi / 2 * t) - 1 + t)


# --------------------------------------------------------------------------------------------------
def test_dit_forward_tiny_vs_golden():
    g = _gold("dit_tiny.pt")
    bb = model_for("tiny").cfm.backbone
    T = g["x"].shape[1]
    mask = (torch.arange(T)[None, :] < g["lens"][:, None]).to(DEV)
    args = [g[k].to(DEV) for k in ("x", "cond", "text", "time")]
    out = bb(*args, mask=mask, cfg_infer=True)
    assert out.shape == g["fwd_cfg"].shape
    lens = g["lens"].tolist()
    for i in range(4):
        n = lens[i % 2]
        assert _rel(out[i, :n], g["fwd_cfg"][i, :n]) < VEL_TOL, i
    out = bb(*args, mask=mask, drop_audio_cond=True)
    for i in range(2):
        assert _rel(out[i, : lens[i]], g["fwd_drop"][i, : lens[i]]) < VEL_TOL
    out = bb(args[0][:1], args[1][:1], args[2][:1], torch.tensor(0.5, device=DEV))
    assert _rel(out, g["fwd_nomask_scalar_t"]) < VEL_TOL


def test_reference_test_config_head_dim_32():
    """The reference's own test configuration (tests/test_checkpoint.py:9-24: dim 64, 2 heads of 32, text_dim 32, ff_mult 2):
    heads are zero-padded to the kernels' 64-wide layout when the weights are packed (engine.DiTWeights). Batched CFG forward
    with ragged lengths and a 3-step CFG sample against fixtures recorded from the live reference (dit_micro.pt)."""
    g = _gold("dit_micro.pt")
    m = model_for("micro")
    bb = m.cfm.backbone
    T = g["x"].shape[1]
    mask = (torch.arange(T)[None, :] < g["lens"][:, None]).to(DEV)
    out = bb(*[g[k].to(DEV) for k in ("x", "cond", "text", "time")], mask=mask, cfg_infer=True)
    assert out.shape == g["fwd_cfg"].shape
    lens = g["lens"].tolist()
    for i in range(4):
        n = lens[i % 2]
        assert _rel(out[i, :n], g["fwd_cfg"][i, :n]) < VEL_TOL, i
    mel, traj = m.cfm.sample(g["s_ref"].to(DEV), g["s_ids"].to(DEV), torch.tensor([120], device=DEV),
                             lens=torch.tensor([40], device=DEV), steps=3, cfg_strength=2.0, sway_sampling_coef=-1.0,
                             y0=g["s_traj"][0])
    assert _rel(mel, g["s_mel"]) < MEL_TOL
    assert torch.equal(mel[:, :40].cpu(), g["s_ref"])


def test_eval_loss_tiny():
    g = _gold("dit_tiny.pt")
    m = model_for("tiny")
    loss = m(g["loss_mel"].to(DEV), g["text"].to(DEV), g["lens"].to(DEV))
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 2e-2


def test_sample_tiny_vs_golden_and_oracle():
    g = _gold("dit_tiny.pt")
    cfm = model_for("tiny").cfm
    y0 = g["s1_traj"][0]
    mel, traj = cfm.sample(torch.zeros(1, 143, 100, device=DEV), g["s1_ids"].to(DEV), torch.tensor([143], device=DEV),
                           lens=torch.tensor([0], device=DEV), steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0,
                           seed=0, y0=y0)
    assert len(traj) == 5 and mel.shape == (1, 143, 100) and mel.dtype == torch.float32
    assert torch.equal(traj[0].cpu(), y0)
    assert _rel(mel, g["s1_mel"]) < MEL_TOL
    t = sway(4, -1.0)
    for i in range(4):  # free-running velocities
        v = (traj[i + 1] - traj[i]).cpu() / (t[i + 1] - t[i])
        vr = (g["s1_traj"][i + 1] - g["s1_traj"][i]) / (t[i + 1] - t[i])
        assert _rel(v, vr) < VEL_TOL, i
    # the oracle agrees with the CUDA path on the second scenario (reference region, no CFG, uniform schedule)
    sd = ref_state_dict("tiny")
    o_mel, o_traj = DO.cfm_sample(sd, g["s2_ref"], g["s2_ids"], torch.tensor([150]), lens=torch.tensor([60]), steps=3,
                                  cfg_strength=0.0, sway_sampling_coef=None, seed=5)
    mel2, traj2 = cfm.sample(g["s2_ref"].to(DEV), g["s2_ids"].to(DEV), torch.tensor([150], device=DEV),
                             lens=torch.tensor([60], device=DEV), steps=3, cfg_strength=0.0, sway_sampling_coef=None,
                             y0=o_traj[0])
    assert _rel(mel2, o_mel) < MEL_TOL and _rel(mel2, g["s2_mel"]) < MEL_TOL
    assert torch.equal(mel2[:, :60].cpu(), g["s2_ref"])


def test_sample_batched_matches_per_utterance():
    """Two utterances of different length in one call == each alone (varlen handling; canonical B=1 semantics)."""
    g = _gold("dit_tiny.pt")
    cfm = model_for("tiny").cfm
    gen = torch.Generator().manual_seed(77)
    durs = [143, 90]
    ids = torch.full((2, 143), -1, dtype=torch.long)
    ids[0] = g["s1_ids"][0]
    ids[1, :90] = torch.randint(4, 65, (90,), generator=gen)
    y0 = torch.zeros(2, 143, 100)
    y0[0] = torch.randn(143, 100, generator=gen)
    y0[1, :90] = torch.randn(90, 100, generator=gen)
    cond = torch.zeros(2, 143, 100)
    kw = dict(steps=3, cfg_strength=2.0, sway_sampling_coef=-1.0)
    mel, _ = cfm.sample(cond.to(DEV), ids.to(DEV), torch.tensor(durs, device=DEV), lens=torch.tensor([0, 0], device=DEV),
                        y0=y0, **kw)
    for b, d in enumerate(durs):
        one, _ = cfm.sample(cond[b:b + 1, :d].to(DEV), ids[b:b + 1, :d].to(DEV), torch.tensor([d], device=DEV),
                            lens=torch.tensor([0], device=DEV), y0=y0[b:b + 1, :d], **kw)
        assert _rel(mel[b, :d], one[0]) < 2e-3, b


def test_synthesize_batch_matches_one_by_one():
    """F5TTS.synthesize_batch == [synthesize(t) for t in texts]: chunking, seeds (seed + chunk index), frame counts
    and waveform lengths exact; waveforms equal up to bf16 arithmetic on differently padded batches."""
    m = model_for("tiny")
    voc = Vocos()
    voc.load_state_dict(GW.fill_state_dict(voc.state_dict(), 4321), strict=True)
    m.set_vocoder(voc.to(DEV).eval())
    gen = torch.Generator().manual_seed(5)
    ref = (torch.rand(24000, generator=gen) * 2 - 1) * 0.3
    texts = ["Сайн байна уу", "Өнөөдөр цаг агаар сайхан байна. Бид хамтдаа уул руу алхаж, голын эрэг дээр амарна.", "Баярлалаа"]
    kw = dict(lang="mn", ref_audio_path=ref, ref_text="Энэ бол жишээ", n_steps=3, cfg_strength=2.0, max_chars_per_chunk=40,
              device=DEV)
    seeds = [11, 22, 33]
    got = m.synthesize_batch(texts, seeds=seeds, max_rows_per_batch=512, **kw)
    assert len(got) == len(texts)
    for t, s, w in zip(texts, seeds, got):
        one = m.synthesize(t, seed=s, **kw)
        assert w.shape == one.shape and not w.is_cuda
        assert _rel(w, one) < 2e-2, t
    with pytest.raises(ValueError, match="text must not be empty"):
        m.synthesize_batch(["ok", "   "], **kw)


def test_sample_midpoint_method_vs_oracle():
    """CFM.sample(method="midpoint"): the explicit midpoint rule (two evaluations per step) against the oracle's
    restatement, with CFG and the sway schedule; "euler" stays the default and an unknown method raises."""
    g = _gold("dit_tiny.pt")
    sd = ref_state_dict("tiny")
    cfm = model_for("tiny").cfm
    y0 = g["s1_traj"][0]
    kw = dict(steps=3, cfg_strength=2.0, sway_sampling_coef=-1.0)
    o_mel, o_traj = DO.cfm_sample(sd, torch.zeros(1, 143, 100), g["s1_ids"], torch.tensor([143]), lens=torch.tensor([0]),
                                  y0=y0, method="midpoint", **kw)
    mel, traj = cfm.sample(torch.zeros(1, 143, 100, device=DEV), g["s1_ids"].to(DEV), torch.tensor([143], device=DEV),
                           lens=torch.tensor([0], device=DEV), y0=y0, method="midpoint", **kw)
    assert len(traj) == 4 and torch.equal(traj[0].cpu(), y0)
    for a, b in zip(traj[1:], o_traj[1:]):
        assert _rel(a, b) < MEL_TOL
    assert _rel(mel, o_mel) < MEL_TOL
    e_mel, _ = DO.cfm_sample(sd, torch.zeros(1, 143, 100), g["s1_ids"], torch.tensor([143]), lens=torch.tensor([0]), y0=y0, **kw)
    assert _rel(o_mel, e_mel) > 3 * _rel(mel, o_mel)  # the CUDA result follows the midpoint rule, not Euler's
    with pytest.raises(ValueError, match="method"):
        cfm.sample(torch.zeros(1, 143, 100, device=DEV), g["s1_ids"].to(DEV), 143, method="rk4")


def test_sample_small_config1():
    g = _gold("sample_small.pt")
    cfm = model_for("small").cfm
    mel, traj = cfm.sample(torch.zeros(1, 143, 100, device=DEV), g["full_ids"].to(DEV), torch.tensor([143], device=DEV),
                           lens=torch.tensor([0], device=DEV), steps=32, cfg_strength=1.5, sway_sampling_coef=-1.0,
                           seed=0, y0=g["y0"])
    assert _rel(mel, g["mel"]) < MEL_TOL
    # teacher-forced velocity at the recorded reference states (steps 0 and 31 have both neighbours stored)
    t = sway(32, -1.0)
    bb = cfm.backbone
    steps = g["traj_steps"]
    ids = g["full_ids"].to(DEV)
    for a, b_ in ((0, 1), (31, 32)):
        xa, xb = g["traj"][steps.index(a)], g["traj"][steps.index(b_)]
        v_ref = (xb - xa) / (t[b_] - t[a])
        both = bb(xa.to(DEV), torch.zeros(1, 143, 100, device=DEV), ids, t[a].to(DEV), mask=torch.ones(1, 143, dtype=torch.bool, device=DEV),
                  cfg_infer=True)
        v = both[:1] + (both[:1] - both[1:]) * 1.5
        assert _rel(v, v_ref) < VEL_TOL, a


def test_base_config2_velocity_and_mel():
    path = os.path.join(HERE, "golden", "sample_base.pt")
    if not os.path.exists(path):
        pytest.skip("sample_base.pt not generated")
    g = torch.load(path, weights_only=False)
    m = model_for("base")
    bb = m.cfm.backbone
    T, ref_len = g["T"], g["ref_len"]
    cond = torch.nn.functional.pad(g["ref_mel"], (0, 0, 0, T - ref_len)).to(DEV)
    ids = g["full_ids"].to(DEV)
    mask = torch.ones(1, T, dtype=torch.bool, device=DEV)
    for name, tv in (("v_t0", 0.0), ("v_t05", 0.5)):
        both = bb(g["y0"].to(DEV), cond, ids, torch.tensor([tv], device=DEV), mask=mask, cfg_infer=True)
        assert both.shape == (2, T, 100)
        assert _rel(both[0], g[name][0]) < VEL_TOL and _rel(both[1], g[name][1]) < VEL_TOL, name
    # fp32 mode at the BASELINE config-2 size (Base, T = 1406, CFG): 1e-4 on the per-NFE velocity
    both = bb(g["y0"].to(DEV), cond, ids, torch.tensor([0.5], device=DEV), mask=mask, cfg_infer=True, precision="fp32")
    assert _rel(both[0], g["v_t05"][0]) < 1e-4 and _rel(both[1], g["v_t05"][1]) < 1e-4
    bb.__dict__["_precise"] = None
    if "mel" in g:
        mel, traj = m.cfm.sample(g["ref_mel"].to(DEV), ids, torch.tensor([T], device=DEV), lens=torch.tensor([ref_len], device=DEV),
                                 steps=32, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0, y0=g["y0"])
        assert _rel(traj[16], g["x16"]) < MEL_TOL
        assert _rel(mel[:, ref_len:], g["mel"][:, ref_len:]) < MEL_TOL
        assert torch.equal(mel[:, :ref_len].cpu(), g["ref_mel"])
    del _MODELS["base"]
    torch.cuda.empty_cache()


def test_sample_validation_messages():
    cfm = model_for("tiny").cfm
    c = torch.zeros(1, 60, 100, device=DEV)
    ids = torch.zeros(1, 60, dtype=torch.long, device=DEV)
    with pytest.raises(ValueError, match="steps must be >= 1"):
        cfm.sample(c, ids, 60, steps=0)
    with pytest.raises(ValueError, match="cfg_strength must be >= 0"):
        cfm.sample(c, ids, 60, cfg_strength=-1.0)
    with pytest.raises(ValueError, match="duration values must be > 0"):
        cfm.sample(c, ids, 0)
    with pytest.raises(ValueError, match="conditioning lens must be <= duration"):
        cfm.sample(c, ids, 50, lens=torch.tensor([60]))
    with pytest.raises(ValueError, match="conditioning sequence length must be <= max duration"):
        cfm.sample(c, ids, 50, lens=torch.tensor([10]))
    with pytest.raises(ValueError, match="exceeds max_duration"):
        cfm.sample(c, ids, 100, max_duration=64)
    with pytest.raises(ValueError, match="lens must have 1 values"):
        cfm.sample(c, ids, 60, lens=torch.tensor([1, 2]))


# --------------------------------------------------------------------------------------------------
def test_logmel_vs_golden():
    g = _gold("mel.pt")
    ap = AudioProcessor()
    assert torch.equal(ap._fb_cpu, g["fb"]) and torch.equal(ap._win_cpu, g["window"])
    gen = torch.Generator().manual_seed(11)
    for case in g["cases"]:
        S = case["n"]
        wav = (torch.rand(S, generator=gen) * 2 - 1) * 0.3
        if S == 48000:
            wav = 0.5 * torch.sin(2 * torch.pi * 220 * torch.arange(S) / 24000)
        mel = ap.mel_spectrogram(wav.to(DEV))
        assert mel.shape == case["mel"].shape
        # log-mel: relative L2; linear mel: absolute error relative to the loudest bin (fp32 FFT round-off is
        # relative to the frame's peak, so near-silent bins of a pure tone only agree in the linear domain)
        ref = case["mel"]
        assert _rel(mel, ref) < 1e-3
        lin, lin_ref = mel.cpu().exp(), ref.exp()
        assert float((lin - lin_ref).abs().max() / lin_ref.max()) < 1e-5
        if case["norm"] is not None:
            out = ap.normalize_audio((wav * 0.37).to(DEV))
            assert torch.allclose(out.cpu(), case["norm"], atol=1e-7)
    batch = torch.stack([(torch.rand(30000, generator=gen) * 2 - 1) * 0.3 for _ in range(3)])
    mb = ap.mel_spectrogram(batch.to(DEV))
    assert mb.shape == (3, 100, 118) and _rel(mb, AO.log_mel(batch)) < 1e-4


def test_istft_vs_reference_decoder_golden():
    from oron_tts_b200 import _lib as L

    g = _gold("istft.pt")
    B, T, _ = g["head"].shape
    h = torch.zeros(B * T, 1056, device=DEV)
    h[:, :1026] = g["head"].reshape(B * T, 1026).to(DEV)
    out = torch.empty(B, (T - 1) * 256, device=DEV)
    L.istft_head(h, torch.hann_window(1024, device=DEV), out, rows_per_batch=T, nb=B, n_frames=T, mode=1)
    assert _rel(out, g["wav"]) < 1e-5


def test_vocos_decode_vs_oracle():
    voc = Vocos()
    sd = GW.fill_state_dict(voc.state_dict(), 4321)
    voc.load_state_dict(sd, strict=True)
    voc = voc.to(DEV).eval()
    gen = torch.Generator().manual_seed(9)
    mel = torch.randn(2, 100, 150, generator=gen) * 1.5 - 3.0
    wav = voc.decode(mel.to(DEV))
    ref = AO.vocos_decode(sd, mel)
    assert wav.shape == ref.shape == (2, 149 * 256)
    assert _rel(wav, ref) < 2e-2
    one = voc.decode(mel[0].to(DEV))
    assert one.shape == (1, 149 * 256)


def test_vocos_decode_large_batch_tiles():
    """Batches with >= 8192 frames take the 256-wide / 2-SM tiles (vocos.py `big`); they must agree with the
    short-utterance tile path clip by clip, and with the CPU oracle."""
    voc = Vocos()
    sd = GW.fill_state_dict(voc.state_dict(), 4321)
    voc.load_state_dict(sd, strict=True)
    voc = voc.to(DEV).eval()
    gen = torch.Generator().manual_seed(10)
    mel = torch.randn(12, 100, 701, generator=gen) * 1.5 - 3.0
    wav = voc.decode(mel.to(DEV))
    assert wav.shape == (12, 700 * 256)
    for b in (0, 5, 11):
        one = voc.decode(mel[b].to(DEV))
        assert _rel(wav[b], one[0]) < 2e-3
    ref = AO.vocos_decode(sd, mel[:1])
    assert _rel(wav[:1], ref) < 2e-2


def test_gpu_training_batch_vs_reference_dataset_and_collator():
    """data.GpuBatcher == TTSDataset.__getitem__ + TTSCollator of the live reference (fixture): ids, lengths, masks and
    padding exact; log-mel within fp32 FFT rounding; and the batch feeds the training engine."""
    from oron_tts_b200.data import GpuBatcher

    g = _gold("data_batch.pt")
    out = GpuBatcher(min_duration_s=0.5, device=DEV)(g["waves"], g["texts"], g["langs"], g["attrs"])
    assert torch.equal(out["text_ids"].cpu(), g["text_ids"]) and torch.equal(out["mel_lengths"].cpu(), g["mel_lengths"])
    assert torch.equal(out["mask"].cpu(), g["mask"])
    mel = out["mel"].cpu()
    assert mel.shape == g["mel"].shape
    for i, t in enumerate(g["mel_lengths"].tolist()):
        assert float(mel[i, :, t:].abs().max()) == 0.0 if t < mel.shape[-1] else True
        assert float((mel[i, :, :t] - g["mel"][i, :, :t]).abs().max()) < 2e-3, i
    from oron_tts_b200.train import TrainEngine

    m = F5TTS.from_config(GW.CONFIGS["tiny"])
    m.load_state_dict(ref_state_dict("tiny"), strict=True)
    eng = TrainEngine(m.to(DEV).train())
    loss = eng.train_step(out["mel"], out["text_ids"], out["mel_lengths"])
    assert bool(torch.isfinite(loss)) and int(eng.skipped) == 0
    with pytest.raises(ValueError, match="too short"):
        GpuBatcher(min_duration_s=1.0, device=DEV)([g["waves"][0][:1000]], ["a"])


def test_precise_fp32_mode_velocity_within_1e4():
    """north_star: per-NFE-step velocity within 1e-4 relative L2 "in fp32 mode". PreciseDiT (fp32 activations, 3-way bf16
    split GEMMs, fp32 attention / row-wise kernels) against the fixtures recorded from the live fp32 reference: batched
    CFG forward with ragged lengths and fillers, a forward with the audio condition dropped, and an unmasked one."""
    from oron_tts_b200.precise import PreciseDiT

    g = _gold("dit_tiny.pt")
    bb = model_for("tiny").cfm.backbone
    pd = PreciseDiT(bb)
    T_ = g["x"].shape[1]
    mask = (torch.arange(T_)[None, :] < g["lens"][:, None]).to(DEV)
    args = (g["x"].to(DEV), g["cond"].to(DEV), g["text"].to(DEV), g["time"].to(DEV))
    valid = torch.cat([mask, mask], 0).cpu()
    out = pd.forward(*args, mask=mask, cfg_infer=True).cpu()
    assert _rel(out[valid], g["fwd_cfg"][valid]) < 1e-4, _rel(out[valid], g["fwd_cfg"][valid])
    out = pd.forward(*args, mask=mask, drop_audio_cond=True).cpu()
    assert _rel(out[mask.cpu()], g["fwd_drop"][mask.cpu()]) < 1e-4
    out = pd.forward(g["x"][:1].to(DEV), g["cond"][:1].to(DEV), g["text"][:1].to(DEV), torch.tensor(0.5, device=DEV)).cpu()
    assert _rel(out, g["fwd_nomask_scalar_t"]) < 1e-4
    # the bf16 production path on the same inputs, for scale
    ref = bb(*args, mask=mask, cfg_infer=True).cpu()
    assert 1e-4 < _rel(ref[valid], g["fwd_cfg"][valid]) < VEL_TOL
    # the same through the public kwargs: DiT.forward(precision="fp32") and a 4-step CFM.sample(precision="fp32")
    out = bb(*args, mask=mask, cfg_infer=True, precision="fp32").cpu()
    assert _rel(out[valid], g["fwd_cfg"][valid]) < 1e-4
    cfm = model_for("tiny").cfm
    mel, traj = cfm.sample(torch.zeros(1, 143, 100, device=DEV), g["s1_ids"].to(DEV), torch.tensor([143], device=DEV),
                           lens=torch.tensor([0], device=DEV), steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0,
                           y0=g["s1_traj"][0], precision="fp32")
    assert _rel(torch.stack(traj), g["s1_traj"]) < 1e-4 and _rel(mel, g["s1_mel"]) < 1e-4
    with pytest.raises(ValueError, match="precision"):
        bb(*args, mask=mask, precision="fp16")


def test_seeded_sample_is_bit_reproducible():
    """The reference's seeded CFM.sample is bit-reproducible run to run. Here the only reduction whose
    order is not fixed by the program is the stream-K split of the FFN down-projection, whose partial sums are added to
    the residual stream in contributor order (gemm_tcgen05.cuh: sk_turn_wait): two seeded runs must agree bit for bit,
    including the whole trajectory, with CFG (batched cond + uncond) and with ragged batches."""
    import weights as GW
    from oron_tts_b200.f5tts import F5TTS

    model = F5TTS.from_config(GW.CONFIGS["small"])
    model.load_state_dict(GW.fill_state_dict(model.state_dict(), GW.SEEDS["small"]), strict=True)
    model = model.to(DEV).eval()
    g = torch.Generator().manual_seed(11)
    for B, T, lens, durs in ((1, 1406, [469], [1406]), (3, 700, [100, 0, 300], [700, 333, 512])):
        cond = (torch.randn(B, T, 100, generator=g) * 1.5 - 3.0).to(DEV)
        ids = torch.randint(4, 65, (B, T), generator=g).to(DEV)
        kw = dict(lens=torch.tensor(lens, device=DEV), steps=6, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=123)
        a, ta = model.cfm.sample(cond, ids, torch.tensor(durs, device=DEV), **kw)
        a, ta = a.clone(), [t.clone() for t in ta]
        for _ in range(2):
            b, tb = model.cfm.sample(cond, ids, torch.tensor(durs, device=DEV), **kw)
            assert torch.equal(a, b)
            assert all(torch.equal(x, y) for x, y in zip(ta, tb))
