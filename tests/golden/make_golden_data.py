"""Golden fixture for the training data path, generated from the LIVE reference.

Run in the build container only:  python tests/golden/make_golden_data.py
Output: data_batch.pt — TTSDataset.__getitem__ (normalize_audio -> log-mel -> ids stretched to the mel length,
dataset.py:187-222) + TTSCollator (zero / -1 padding, dataset.py:334-362) on seeded waveforms, and the batch plan of
DynamicBatchSampler (dataset.py:375-423) with its seeded epoch order.
"""

from __future__ import annotations

import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
_stub = types.ModuleType("soundfile")
_stub.write = lambda *a, **k: None
_stub.info = lambda *a, **k: None
_stub.read = lambda *a, **k: None
sys.modules["soundfile"] = _stub
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(HERE))]
os.chdir(tempfile.gettempdir())

import torch  # noqa: E402

from src.data.dataset import DynamicBatchSampler, TTSCollator, TTSDataset  # noqa: E402

assert os.path.realpath(sys.modules["src"].__path__[0]).startswith(REF)


def main() -> None:
    gen = torch.Generator().manual_seed(31)
    lengths = [36000, 51234, 24000, 47999]
    waves = [((torch.rand(n, generator=gen) * 2 - 1) * (0.05 + 0.2 * i)).numpy() for i, n in enumerate(lengths)]
    waves[2] = waves[2] * 0.0  # a silent clip (normalize_audio passes it through, audio.py:73-77)
    texts = ["Сайн байна уу", "Өнөөдөр цаг агаар сайхан байна.", "Баярлалаа", "2024 онд 25 хүн ирсэн"]
    langs = ["mn", "mn", "mn", "mn"]
    attrs = [[], ["[FEMALE]"], [], ["[MALE]", "[AGE_20S]"]]
    ds = TTSDataset(audio_arrays=waves, texts=texts, langs=langs, attr_tokens_list=attrs, min_duration_s=0.5)
    items = [ds[i] for i in range(len(waves))]
    batch = TTSCollator()(items)
    out = dict(waves=[torch.from_numpy(w) for w in waves], texts=texts, langs=langs, attrs=attrs,
               mel=batch["mel"], text_ids=batch["text_ids"], mask=batch["mask"], mel_lengths=batch["mel_lengths"])
    rnd = torch.Generator().manual_seed(5)
    durations = (torch.rand(57, generator=rnd) * 29 + 1).tolist()
    plans = {}
    for thr, mx in ((3000, 0), (8192, 6), (1200, 0)):
        smp = DynamicBatchSampler(durations, frames_threshold=thr, max_samples=mx)
        smp.set_epoch(3)
        plans[(thr, mx)] = dict(batches=smp.batches, epoch3=list(iter(smp)))
    out["durations"], out["plans"] = durations, plans
    torch.save(out, os.path.join(HERE, "data_batch.pt"))
    print("mel", tuple(batch["mel"].shape), "lengths", batch["mel_lengths"].tolist(), "ids0", batch["text_ids"][0, :8].tolist())


if __name__ == "__main__":
    main()
