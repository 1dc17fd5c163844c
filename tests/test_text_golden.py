"""Bit-exact integer/string contracts of the host-side front end against vectors produced by the live
reference (tests/golden/make_golden.py -> text.pt): cleaning, ids, chunking, token stretching and the
per-segment plan (stretched ids, T_total, ref_len, target frame count)."""

import os

import pytest
import torch

from oron_tts_b200.f5tts import F5TTS, _stretch_text_to_len, split_text_for_synthesis
from oron_tts_b200.text import CyrillicTokenizer, NumberNormalizer, TextCleaner, validate_language

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "text.pt"), weights_only=False)


def test_clean_and_ids_bit_exact():
    tc = TextCleaner()
    assert len(GOLD["clean"]) >= 50
    for case in GOLD["clean"]:
        assert tc.clean(case["text"], lang=case["lang"]) == case["cleaned"], case["text"]
        assert tc.text_to_sequence(case["text"], lang=case["lang"], attr_tokens=case.get("attr")) == case["ids"], case["text"]


def test_reference_pinned_facts():
    # scripts/test_pipeline.py:87-136 and README number table
    tok, tc, nn_ = CyrillicTokenizer(), TextCleaner(), NumberNormalizer("mn")
    assert tok.vocab_size == 65 == tc.vocab_size
    assert tok.encode("а", attr_tokens=["[FEMALE]"])[1] == 6
    assert 3 not in tc.text_to_sequence("сайн байна уу") and 3 not in tc.text_to_sequence("сәлем", lang="kz")
    assert not any(ch.isdigit() for ch in tc.clean("2024 онд"))
    assert tc.text_to_sequence("Сайн байна уу") == [4, 30, 11, 21, 25, 53, 12, 11, 21, 25, 11, 53, 32, 32]
    assert nn_.convert(10) == "арав" and nn_.convert(25) == "хорин тав" and nn_.convert(100) == "зуу"
    assert nn_.convert_ordinal(1) == "нэгдүгээр" and nn_.convert(2024) == "хоёр мянга хорин дөрөв"
    assert nn_.normalize_text("3/4") == "дөрөвдүгээрийн гурав" and nn_.normalize_text("2024-ны") == "хоёр мянга хорин дөрвөн"
    with pytest.raises(ValueError):
        validate_language("en")
    with pytest.raises(ValueError):
        tok.encode("а", lang="ru")


def test_chunking_and_stretching():
    for case in GOLD["chunks"]:
        assert split_text_for_synthesis(case["text"], case["max_chars"]) == case["chunks"], case["text"][:30]
    for case in GOLD["stretch"]:
        assert _stretch_text_to_len(case["ids"], case["T"]) == case["out"]


def test_segment_plans_match_reference_synthesize():
    """ids / frame counts the reference's _synthesize_segment hands to CFM.sample and Vocos."""
    model = F5TTS.from_config({"model": dict(dim=128, depth=1, heads=2, text_dim=64, conv_layers=1)})
    for plan in GOLD["plans"]:
        c = plan["case"]
        ref_mel = None
        if c.get("ref_samples"):
            ref_mel = torch.zeros(100, 1 + c["ref_samples"] // 256)  # frame count rule of the STFT front end
        got = model.prepare_segment(c["text"], c["lang"], ref_mel, c.get("ref_text"), c.get("speed", 1.0),
                                    c.get("target_duration_s"))
        assert got["ref_len"] == plan["ref_len"]
        assert got["T_total"] == plan["duration"]
        assert got["full_ids"] == plan["text_ids"]
        assert got["target_len"] == plan["mel_frames"]
        assert (got["target_len"] - 1) * 256 >= 0  # waveform length rule S = (target_len - 1) * hop


def test_synthesize_argument_validation():
    model = F5TTS.from_config({"model": dict(dim=128, depth=1, heads=2, text_dim=64, conv_layers=1)})
    for kw in (dict(lang="en"), dict(n_steps=0), dict(cfg_strength=-1), dict(speed=0), dict(target_duration_s=0),
               dict(max_chars_per_chunk=-1), dict(pause_s=-0.1)):
        with pytest.raises(ValueError):
            model.synthesize("сайн", device="cpu", **kw)


def test_dynamic_batch_plan_matches_reference_sampler():
    """data.dynamic_batches / epoch_order against DynamicBatchSampler (fixture recorded from the live reference)."""
    import os

    import torch

    from oron_tts_b200.data import dynamic_batches, epoch_order

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_batch.pt"), weights_only=False)
    for (thr, mx), ref in g["plans"].items():
        plan = dynamic_batches(g["durations"], frames_threshold=thr, max_samples=mx)
        assert plan == ref["batches"], (thr, mx)
        assert [plan[i] for i in epoch_order(len(plan), 3)] == ref["epoch3"]
    assert sorted(i for b in dynamic_batches(g["durations"], 3000) for i in b) == list(range(len(g["durations"])))
