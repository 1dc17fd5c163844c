"""Generate the golden fixtures in this directory from the LIVE reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py [--base]

The reference's ``src`` package shadows this repository's import shim of the same name, so this script
never puts the repository root on sys.path; it loads ``weights.py`` by file path.
Outputs (all torch.save files, fp32 / int64):
    text.pt        cleaned strings + token ids, chunking, stretching, segment plans (ids, frame counts)
    mel.pt         log-mel of seeded waveforms through AudioProcessor.mel_spectrogram
    istft.pt       in-repo VocosDecoder (src/models/decoder.py) head activations -> waveform
    dit_tiny.pt    tiny DiT: single forward (cfg_infer), eval loss, 4-step sample with trajectory
    dit_micro.pt   (--micro) the reference's own test configuration (head_dim 32, dim 64): CFG forward + 3-step sample
    sample_small.pt  BASELINE config 1 (Small, "Сайн байна уу", 32 NFE, CFG 1.5, seed 0)
    sample_base.pt   BASELINE config 2 (Base, 469 + 937 frames, CFG 2.0): teacher-forced velocities and,
                     with --base-full, the free-running final mel
"""

from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

# --- import the reference with a stub soundfile (src/utils/audio.py:14 imports it at module top) ---------
_stub = types.ModuleType("soundfile")
_stub.write = lambda *a, **k: None
sys.modules["soundfile"] = _stub
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(HERE))]
os.chdir(tempfile.gettempdir())

import torch  # noqa: E402

from src.models.decoder import VocosDecoder  # noqa: E402
from src.models.f5tts import F5TTS, _stretch_text_to_len, split_text_for_synthesis  # noqa: E402
from src.utils.audio import AudioProcessor  # noqa: E402
from src.utils.text_cleaner import TextCleaner  # noqa: E402

assert os.path.realpath(sys.modules["src"].__path__[0]).startswith(REF), "must import the reference's src package"

_spec = importlib.util.spec_from_file_location("golden_weights", os.path.join(HERE, "weights.py"))
W = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(W)

torch.set_grad_enabled(False)


def build(name: str) -> F5TTS:
    model = F5TTS.from_config(W.CONFIGS[name]).eval()
    model.load_state_dict(W.fill_state_dict(model.state_dict(), W.SEEDS[name]), strict=True)
    return model


# ------------------------------------------------------------------------------------------------------
TEXTS_MN = [
    "Сайн байна уу", "сайн байна уу", "Монгол улсын нийслэл Улаанбаатар хот.", "2024 онд", "10", "25", "100", "1-р", "2024",
    "3/4", "1/2", "2024-ны 5 сарын 17", "Өнөөдөр 2024/03/15 өдөр 12:30 цагт уулзъя!", "Үнэ нь 1,234,567₮ буюу $100 байна...",
    "Температур -15°C, чийгшил 45% байв.", "3.14 бол пи тоо", "Утас: +976 9911 2233", "10-20 хүн ирнэ", "20-р зуун, XV зуун, IV бүлэг",
    "5 км зам, 3 кг алим, 250 мл ус, 7 м өндөр", "5-ын 3, 1-ний өдөр, 3-ийн даваа", "1000000 төгрөг, 2500000000 хүн, 1000000000000",
    "«Сайн уу?» гэж тэр хэлэв — тийм ээ… үнэхээр!!!", "Hello world 123 abc", "   олон    зай   ", "А + Б = В, 5 > 3, 2 < 4 ~ 7", "11 12 13 19 20 21 99 101 110 111 999 1001 2000 10000",
    "100 EUR, 50 USD, 7 KRW, €20, £5", "15:45:30", "17.05.2023", "1 234 ширхэг", "тов. Бат 1990 г. төрсөн", "ж. нь", "0 хэм, 0.5, 007",
    "Ёс суртахуун, ьъ тэмдэг; (хаалт) \"ишлэл\" 'дан'", "3-дугаар байр, 4-дүгээр анги, 2-ахь удаа", "1.5 цаг. 2,5 минут", "123456789",
]
TEXTS_KZ = [
    "Сәлем, қалайсың?", "Қазақстан Республикасының астанасы", "2024 жылы 5-ші сынып", "3/4 және 1/2", "25% өсім, 100₸, 15°C",
    "12:30 сағат, 10-20 адам", "1000000 теңге 3.14", "ж. басы, 5 км, 10 сағ. кейін", "ұлы һәм ғажайып іңкәр", "+7 701 123 4567", "XIX ғасыр", "7 м 3 кг",
]
CHUNK_CASES = [
    ("", 120), ("   ", 10), ("Сайн байна уу", 120), ("Сайн байна уу. " * 20, 120), ("а" * 300, 120), ("үг " * 100, 50),
    ("Нэг, хоёр; гурав: дөрөв. Тав! Зургаа? Долоо… найм " * 6, 60), ("богино", 0), ("Энэ бол урт өгүүлбэр бөгөөд таслалгүй үргэлжилсээр л байна " * 4, 100),
    ("нэг.хоёр.гурав " * 30, 33), ("x" * 119 + ". " + "y" * 10, 120),
]


def gen_text() -> dict:
    tc = TextCleaner()
    out = {"clean": [], "chunks": [], "stretch": [], "plans": []}
    for lang, texts in (("mn", TEXTS_MN), ("kz", TEXTS_KZ)):
        for t in texts:
            out["clean"].append(dict(lang=lang, text=t, cleaned=tc.clean(t, lang=lang), ids=tc.text_to_sequence(t, lang=lang)))
    out["clean"].append(dict(lang="mn", text="Сайн", cleaned=tc.clean("Сайн"), attr=["[FEMALE]", "[YOUNG]", "[BOGUS]"],
                             ids=tc.text_to_sequence("Сайн", lang="mn", attr_tokens=["[FEMALE]", "[YOUNG]", "[BOGUS]"])))
    for text, mc in CHUNK_CASES:
        out["chunks"].append(dict(text=text, max_chars=mc, chunks=split_text_for_synthesis(text, mc)))
    for n, T in [(0, 5), (3, 10), (14, 143), (10, 10), (12, 7), (1, 50), (7, 469), (60, 937)]:
        ids = list(range(4, 4 + n))
        out["stretch"].append(dict(ids=ids, T=T, out=_stretch_text_to_len(ids, T)))

    # segment plans: run the reference's own _synthesize_segment with the sampler/vocoder intercepted
    model = F5TTS.from_config(W.CONFIGS["tiny"]).eval()
    captured = {}

    def fake_sample(cond, text_ids, duration, lens, **kw):
        captured.update(text_ids=text_ids.clone(), duration=duration.clone(), lens=lens.clone(), cond_shape=tuple(cond.shape), kw=kw)
        return torch.zeros(1, int(duration[0]), model.n_mels), []

    class _FakeVocos:
        def decode(self, mel):
            captured["mel_shape"] = tuple(mel.shape)
            return torch.zeros(1, (mel.shape[-1] - 1) * 256)

    model.cfm.sample = fake_sample
    model._get_vocos = lambda device: _FakeVocos()
    gen = torch.Generator().manual_seed(7)
    cases = [
        dict(text="Сайн байна уу", lang="mn"),
        dict(text="Сайн байна уу", lang="mn", speed=0.8),
        dict(text="Сайн байна уу", lang="mn", target_duration_s=10.0),
        dict(text="Өнөөдөр цаг агаар сайхан байна, 25 хэм дулаан.", lang="mn", ref_samples=120000, ref_text="Энэ бол жишээ өгүүлбэр юм."),
        dict(text="Өнөөдөр цаг агаар сайхан байна.", lang="mn", ref_samples=72000, ref_text=None),
        dict(text="Сәлем әлем", lang="kz", ref_samples=50000, ref_text="Қайырлы таң", speed=1.3),
        dict(text="Сайн", lang="mn", ref_samples=120000, ref_text="Энэ бол жишээ өгүүлбэр юм.", target_duration_s=10.0),
        dict(text="а", lang="mn", target_duration_s=0.001),
    ]
    for c in cases:
        captured.clear()
        ref_samples = c.get("ref_samples")
        if ref_samples:
            wav = (torch.rand(ref_samples, generator=gen) * 2 - 1) * 0.3
            model._audio_processor.load_audio = lambda path, _w=wav: (_w, 24000)
        model._synthesize_segment(text=c["text"], lang=c["lang"], ref_audio_path="ref.wav" if ref_samples else None,
                                  ref_text=c.get("ref_text"), n_steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0,
                                  speed=c.get("speed", 1.0), target_duration_s=c.get("target_duration_s"), seed=0, device="cpu")
        out["plans"].append(dict(case=c, text_ids=captured["text_ids"][0].tolist(), duration=int(captured["duration"][0]),
                                 ref_len=int(captured["lens"][0]), mel_frames=captured["mel_shape"][-1]))
    return out


def gen_mel() -> dict:
    ap = AudioProcessor()
    gen = torch.Generator().manual_seed(11)
    out = []
    for S in (48000, 7777, 120000):
        wav = (torch.rand(S, generator=gen) * 2 - 1) * 0.3
        if S == 48000:  # the reference smoke test's 220 Hz sine (scripts/test_pipeline.py:41-44)
            wav = 0.5 * torch.sin(2 * torch.pi * 220 * torch.arange(S) / 24000)
        out.append(dict(n=S, seed_note="rand*0.3 from Generator(11) in order 48000(sine),7777,120000", mel=ap.mel_spectrogram(wav),
                        norm=ap.normalize_audio(wav * 0.37) if S == 7777 else None))
    return {"cases": out, "fb": ap._mel_transform.mel_scale.fb.clone(), "window": ap._mel_transform.spectrogram.window.clone()}


def gen_istft() -> dict:
    torch.manual_seed(5)
    dec = VocosDecoder().eval()
    gen = torch.Generator().manual_seed(5)
    mel = torch.randn(2, 100, 40, generator=gen)
    wav = dec(mel)
    # replay the head input so the iSTFT stage can be checked in isolation
    x = dec.input_proj(mel).transpose(1, 2)
    x = dec.norm_pre(x).transpose(1, 2)
    for layer in dec.layers:
        x = layer(x)
    head = dec.istft_head(dec.norm_post(x.transpose(1, 2)))
    return dict(head=head, wav=wav)


def gen_dit_tiny() -> dict:
    model = build("tiny")
    gen = torch.Generator().manual_seed(3)
    B, T = 2, 150
    lens = torch.tensor([150, 97])
    x = torch.randn(B, T, 100, generator=gen)
    cond = torch.randn(B, T, 100, generator=gen) * (torch.arange(T)[None, :, None] < 40)
    text = torch.randint(4, 65, (B, T), generator=gen)
    text[0, 120:] = -1
    text[1, 97:] = -1
    text[1, 30:35] = -1
    time = torch.tensor([0.3, 0.8])
    mask = torch.arange(T)[None, :] < lens[:, None]
    bb = model.cfm.backbone
    out = dict(x=x, cond=cond, text=text, time=time, lens=lens)
    out["fwd_cfg"] = bb(x, cond, text, time, mask=mask, cfg_infer=True)
    out["fwd_drop"] = bb(x, cond, text, time, mask=mask, drop_audio_cond=True, drop_text=False)
    out["fwd_nomask_scalar_t"] = bb(x[:1], cond[:1], text[:1], torch.tensor(0.5))
    # eval-mode deterministic loss (flow.py:113-128, 136-138)
    mel = torch.randn(B, 100, T, generator=gen)
    out["loss_mel"] = mel
    out["loss"] = model(mel, text, lens)
    # B=1 sample, 4 steps, with and without CFG; y0 recorded (CPU generator stream)
    ids = torch.tensor([_stretch_text_to_len([4, 30, 11, 21, 25, 53, 12, 11, 21, 25, 11, 53, 32, 32], 143)])
    cond1 = torch.zeros(1, 143, 100)
    mel1, traj1 = model.cfm.sample(cond1, ids, torch.tensor([143]), lens=torch.tensor([0]), steps=4, cfg_strength=2.0,
                                   sway_sampling_coef=-1.0, seed=0)
    out["s1_ids"], out["s1_traj"], out["s1_mel"] = ids, torch.stack(traj1), mel1
    # B=1 voice-cloning shaped sample: 60-frame "reference" + 90 target, uniform schedule, no CFG
    refmel = torch.randn(1, 60, 100, generator=gen) * 1.5 - 3
    ids2 = torch.randint(4, 65, (1, 150), generator=gen)
    mel2, traj2 = model.cfm.sample(refmel, ids2, torch.tensor([150]), lens=torch.tensor([60]), steps=3, cfg_strength=0.0,
                                   sway_sampling_coef=None, seed=5)
    out["s2_ref"], out["s2_ids"], out["s2_traj"], out["s2_mel"] = refmel, ids2, torch.stack(traj2), mel2
    return out


def gen_dit_micro() -> dict:
    """The reference's own test configuration (tests/test_checkpoint.py:9-24: dim 64, 2 heads of 32, text_dim 32, ff_mult 2):
    batched CFG forward with ragged lengths and a 3-step CFG sample."""
    model = build("micro")
    gen = torch.Generator().manual_seed(11)
    B, T = 2, 140
    lens = torch.tensor([140, 77])
    x = torch.randn(B, T, 100, generator=gen)
    cond = torch.randn(B, T, 100, generator=gen) * (torch.arange(T)[None, :, None] < 30)
    text = torch.randint(4, 65, (B, T), generator=gen)
    text[0, 100:] = -1
    text[1, 77:] = -1
    time = torch.tensor([0.2, 0.7])
    mask = torch.arange(T)[None, :] < lens[:, None]
    bb = model.cfm.backbone
    out = dict(x=x, cond=cond, text=text, time=time, lens=lens)
    out["fwd_cfg"] = bb(x, cond, text, time, mask=mask, cfg_infer=True)
    ids = torch.randint(4, 65, (1, 120), generator=gen)
    refmel = torch.randn(1, 40, 100, generator=gen) * 1.5 - 3
    mel, traj = model.cfm.sample(refmel, ids, torch.tensor([120]), lens=torch.tensor([40]), steps=3, cfg_strength=2.0,
                                 sway_sampling_coef=-1.0, seed=7)
    out["s_ref"], out["s_ids"], out["s_traj"], out["s_mel"] = refmel, ids, torch.stack(traj), mel
    return out


def gen_sample_small() -> dict:
    model = build("small")
    tc = TextCleaner()
    ids = tc.text_to_sequence("Сайн байна уу", lang="mn")
    T = max(50, int(len("Сайнбайнауу") * 13 / 1.0))
    full = torch.tensor([_stretch_text_to_len(ids, T)])
    cond = torch.zeros(1, T, 100)
    mel, traj = model.cfm.sample(cond, full, torch.tensor([T]), lens=torch.tensor([0]), steps=32, cfg_strength=1.5,
                                 sway_sampling_coef=-1.0, seed=0)
    traj = torch.stack(traj)
    # teacher-forced velocities at a few steps: v_i = (x_{i+1} - x_i) / dt_i is implied by the trajectory
    return dict(ids=ids, T=T, full_ids=full, y0=traj[0].clone(), traj_steps=[0, 1, 8, 16, 24, 31, 32],
                traj=traj[[0, 1, 8, 16, 24, 31, 32]].clone(), mel=mel.clone())


def gen_sample_base(full_run: bool) -> dict:
    model = build("base")
    gen = torch.Generator().manual_seed(2)
    ref_len, tgt_len = 469, 937
    T = ref_len + tgt_len
    refmel = torch.randn(1, ref_len, 100, generator=gen) * 1.5 - 3.0
    ref_ids = torch.randint(11, 65, (60,), generator=gen).tolist()
    tgt_ids = torch.randint(11, 65, (120,), generator=gen).tolist()
    full = torch.tensor([_stretch_text_to_len(ref_ids, ref_len) + _stretch_text_to_len(tgt_ids, tgt_len)])
    y0 = torch.randn(T, 100, generator=torch.Generator().manual_seed(0))[None]
    cond = torch.nn.functional.pad(refmel, (0, 0, 0, tgt_len))
    mask = torch.ones(1, T, dtype=torch.bool)
    bb = model.cfm.backbone
    out = dict(ref_mel=refmel, full_ids=full, ref_len=ref_len, T=T, y0=y0)
    # teacher-forced CFG velocity at t = 0 and t = 0.5 on x = y0 (one NFE each, ~4 s on 8 cores)
    for name, t in (("v_t0", 0.0), ("v_t05", 0.5)):
        both = bb(y0, cond, full, torch.tensor([t]), mask=mask, cfg_infer=True)
        out[name] = both
        bb.clear_cache()
    if full_run:
        mel, traj = model.cfm.sample(refmel, full, torch.tensor([T]), lens=torch.tensor([ref_len]), steps=32, cfg_strength=2.0,
                                     sway_sampling_coef=-1.0, seed=0)
        assert torch.equal(traj[0], y0)
        out["mel"] = mel.clone()
        out["x16"] = traj[16].clone()
    return out


def gen_state_keys() -> dict:
    """state_dict key -> shape of the reference F5TTS for every BASELINE config."""
    out = {}
    for name, cfg in W.CONFIGS.items():
        m = F5TTS.from_config(cfg)
        out[name] = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    return out


if __name__ == "__main__":
    if "--micro" in sys.argv:
        torch.save(gen_dit_micro(), os.path.join(HERE, "dit_micro.pt"))
        print("dit_micro written")
        sys.exit(0)
    if "--keys" in sys.argv:
        torch.save(gen_state_keys(), os.path.join(HERE, "state_keys.pt"))
        print("state_keys written")
        sys.exit(0)
    torch.save(gen_text(), os.path.join(HERE, "text.pt"))
    torch.save(gen_mel(), os.path.join(HERE, "mel.pt"))
    torch.save(gen_istft(), os.path.join(HERE, "istft.pt"))
    torch.save(gen_dit_tiny(), os.path.join(HERE, "dit_tiny.pt"))
    print("text / mel / istft / dit_tiny written", flush=True)
    if "--small" in sys.argv or "--all" in sys.argv:
        torch.save(gen_sample_small(), os.path.join(HERE, "sample_small.pt"))
        print("sample_small written", flush=True)
    if "--base" in sys.argv or "--base-full" in sys.argv or "--all" in sys.argv:
        torch.save(gen_sample_base("--base-full" in sys.argv or "--all" in sys.argv), os.path.join(HERE, "sample_base.pt"))
        print("sample_base written", flush=True)
