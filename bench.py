#!/usr/bin/env python
"""Headline benchmark: target-audio-seconds per second at 32 NFE, Base DiT (BASELINE.json config 2:
5 s synthetic reference mel + 10 s target, CFG 2.0, sway -1, bf16 tensor-core arithmetic, fp32 state).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference --gpus N ...             # the UNMODIFIED reference (baseline/_ref) on the host cores
  python bench.py --impl reference-gpu                      # internal: the reference in PyTorch eager on cuda:0 (JSON dict)

A "step" is one utterance through the hot path: CFM.sample (32 x [DiT forward on cond+uncond] + CFG/Euler)
followed by the Vocos decode of the target region. `value` = whole-job target-audio seconds / device time
with inputs resident in HBM; `e2e` = the same through F5TTS.synthesize with the reference waveform in pinned
host memory and the result copied back to the host. Multi-GPU: independent utterances per rank, no data-path
collective (weak scaling); only the timing barrier / max-over-ranks uses NCCL.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

METRIC = "target_audio_seconds_per_second_32nfe_base_dit"
UNIT = "audio-s/s"
REF_SAMPLES, REF_LEN, TGT_LEN, STEPS_NFE, CFG, SWAY = 120000, 469, 937, 32, 2.0, -1.0
T_TOTAL = REF_LEN + TGT_LEN
AUDIO_S = (TGT_LEN - 1) * 256 / 24000.0  # seconds of waveform the vocoder emits for the target region
BENCH_TEXT = ("Өнөөдөр цаг агаар сайхан байна, бид хамтдаа уул руу алхаж, голын эрэг дээр амарч, "
              "орой нь гэртээ харина.")
BENCH_REF_TEXT = "Энэ бол жишээ өгүүлбэр бөгөөд дуу хоолойг дуурайхад ашиглагдана."


def algorithmic_flops_per_nfe(T: int, D: int = 1024, depth: int = 22, nb: int = 2) -> dict:
    """per row 22*(8D^2 + 16D^2) GEMM + attention 4*T*D per layer (rows = nb*T)."""
    rows = nb * T
    gemm = rows * depth * 24 * D * D
    attn = rows * depth * 4 * T * D
    io = rows * (2 * 128 * D + 2 * 2 * D * 64 * 31 + 2 * D * 100)
    return dict(gemm=gemm, attn=attn, other=io, total=gemm + attn + io)


# ----------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 200 ms during the timed region (pynvml)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.2)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------
def build_workload(device, seed: int):
    """Synthetic config-2 inputs + random-init Base model (zero-init tensors re-randomised)."""
    import weights as GW

    from oron_tts_b200.f5tts import F5TTS, _stretch_text_to_len
    from oron_tts_b200.vocos import Vocos

    model = F5TTS.from_config(GW.CONFIGS["base"])
    model.load_state_dict(GW.fill_state_dict(model.state_dict(), GW.SEEDS["base"]), strict=True)
    model = model.to(device).eval()
    voc = Vocos()
    voc.load_state_dict(GW.fill_state_dict(voc.state_dict(), 4321), strict=True)
    voc = voc.to(device).eval()
    model.set_vocoder(voc)
    g = torch.Generator().manual_seed(seed)
    ref_mel = torch.randn(1, REF_LEN, 100, generator=g) * 1.5 - 3.0
    ref_ids = torch.randint(11, 65, (60,), generator=g).tolist()
    tgt_ids = torch.randint(11, 65, (120,), generator=g).tolist()
    full = torch.tensor([_stretch_text_to_len(ref_ids, REF_LEN) + _stretch_text_to_len(tgt_ids, TGT_LEN)])
    ref_wav = ((torch.rand(REF_SAMPLES, generator=g) * 2 - 1) * 0.3).pin_memory()
    return model, voc, ref_mel.to(device), full.to(device), ref_wav


def run_ours(args) -> dict:
    import torch.distributed as dist

    from oron_tts_b200 import _lib as L

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.lib()
    model, voc, ref_mel, ids, ref_wav = build_workload(dev, seed=100 + rank)
    cfm = model.cfm
    # The mel / waveform payload is device resident for the `value` leg; token ids, durations and lengths are host-side
    # metadata, as in F5TTS.synthesize: CFM.sample validates them without a device read-back (a read-back is a host sync that
    # would wait for the previous step's whole ODE loop and keep consecutive utterances from pipelining)
    ids = ids.cpu()
    dur = T_TOTAL
    lens = torch.tensor([REF_LEN])

    # The timed legs draw fresh noise every step (seed=None: what scripts/infer.py does unless --seed is given). A pinned
    # seed switches CFM.sample to its bit-reproducible mode (no stream-K split of the FFN down-projection): timed
    # separately below and reported as `deterministic`.
    def step_device(seed, pinned=False):
        mel, _ = cfm.sample(ref_mel, ids, dur, lens=lens, steps=STEPS_NFE, cfg_strength=CFG, sway_sampling_coef=SWAY,
                            seed=seed if pinned else None)
        return voc.decode(mel[:, REF_LEN:, :].transpose(1, 2))

    def step_e2e(seed):
        return model.synthesize(BENCH_TEXT, lang="mn", ref_audio_path=ref_wav, ref_text=BENCH_REF_TEXT, n_steps=STEPS_NFE,
                                cfg_strength=CFG, sway_sampling_coef=SWAY, target_duration_s=10.0, seed=None, device=str(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for i in range(k):
            out = fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for i in range(args.warmup):
        wav = step_device(1000 + i)
    eng = cfm.backbone.engine()
    sampler = ClockSampler(local)
    sampler.start()
    n0, r0 = L.launch_count(), eng.replayed_launches
    ms, wav = timed(step_device, args.steps)
    launches = (L.launch_count() - n0) + (eng.replayed_launches - r0)
    assert wav.shape[-1] == (TGT_LEN - 1) * 256 and bool(torch.isfinite(wav).all())
    # end to end through the public API, host buffers in, host waveform out
    for i in range(min(args.warmup, 2)):
        step_e2e(2000 + i)
    ms_e2e, wav_host = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert not wav_host.is_cuda and wav_host.numel() == (TGT_LEN - 1) * 256

    # bit-reproducible (seeded) mode: same workload, seeds pinned; two runs from one seed must agree bit for bit
    step_device(1, pinned=True)
    ms_det, wav_a = timed(lambda i: step_device(1, pinned=True), 3)
    wav_b = step_device(1, pinned=True)
    deterministic = {"ms_per_nfe": round(ms_det / 3 / STEPS_NFE, 4), "audio_s_per_s": round(world * 3 * AUDIO_S / (ms_det / 1e3), 3),
                     "bit_identical_reruns": bool(torch.equal(wav_a, wav_b)),
                     "note": "CFM.sample(seed=...) : fixed-order reductions (stream-K of the FFN down-projection off)"}
    value = world * args.steps * AUDIO_S / (ms / 1e3)
    value_e2e = world * args.steps * AUDIO_S / (ms_e2e / 1e3)
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 3), "ms_per_nfe": round(ms / args.steps / STEPS_NFE, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg2: Base DiT (dim 1024, depth 22), 469 ref + 937 target frames, 32 NFE, CFG 2.0, sway -1, "
                               "+ Vocos decode of the target; 1 utterance per GPU per step",
                   "l2": "weights (856 MB bf16) exceed the 126 MB L2 and are re-streamed every NFE; no explicit flush",
                   "inputs": "value: reference mel resident in HBM, token ids / lengths host-side metadata (11 KB H2D per step inside the "
                             "timed region); e2e: pinned-host waveform + strings in, host waveform out",
                   "parallelism": f"utterance-sharded x{world} (no data-path collective)"},
        "e2e": {"value": round(value_e2e, 3), "unit": UNIT, "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": int(ref_wav.numel() * 4 + T_TOTAL * 8 + 16),
                "d2h_bytes_per_step": int(wav_host.numel() * 4 + 16)},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "deterministic": deterministic,
    }
    if not args.no_secondary:
        cfg3 = secondary_cfg3(model, cfm, voc, dev, rank, world, args.cfg3_utterances or 32 * world, barrier)
        # the named 256-utterance job at EVERY N (strong scaling: total work fixed), device-resident leg only
        cfg3s = None if args.no_cfg3_strong else secondary_cfg3(model, cfm, voc, dev, rank, world, 256, barrier, e2e=False)
        cfg5 = secondary_cfg5(model, dev, rank, world, barrier)
        if rank == 0:
            with torch.inference_mode():
                out["secondary"] = secondary_cfg4(model, voc, dev)
                out["secondary_cfg1"] = secondary_cfg1(voc, dev)
            out["secondary_cfg3"] = cfg3
            if cfg3s is not None:
                cfg3s["scaling"] = "strong"
                out["secondary_cfg3_strong"] = cfg3s
            out["secondary_cfg5"] = cfg5
    if rank == 0:
        print("[bench] main legs done: " + json.dumps({k: out[k] for k in ("value", "ms_per_step", "ms_per_nfe", "e2e")}),
              file=sys.stderr, flush=True)
        with torch.inference_mode():
            out["roofline"] = roofline(eng, cfm, ref_mel, ids, dur, lens)
        if not args.no_cpu_baseline and world == 1:  # contract: the CPU baseline is reported on rank 0 at N=1 only
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
            out["cpu_baseline"] = cpu_baseline(nfe=4, reps=3)
            try:
                out["cpu_baseline"]["other_configs"] = cpu_side_figures()
            except Exception as e:  # pragma: no cover
                out["cpu_baseline"]["other_configs"] = {"unavailable": f"{type(e).__name__}: {e}"}
        if not args.no_gpu_baseline and world == 1:
            out["gpu_eager_baseline"] = gpu_eager_baseline()
            ge = out["gpu_eager_baseline"]
            for k in ("eager_fp32", "eager_tf32", "eager_bf16_autocast", "compiled_bf16_autocast"):
                if isinstance(ge.get(k), dict) and "audio_s_per_s" in ge[k]:
                    ge[k]["ours_speedup"] = round(out["value"] / ge[k]["audio_s_per_s"], 2)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else {}


def roofline(eng, cfm, ref_mel, ids, dur, lens) -> dict:
    """Dominant kernel = gemm_bf16_tcgen05 (all Linear / conv launches of an NFE). achieved = algorithmic GEMM
    FLOPs of one NFE / CUDA-event time of a graph replaying exactly those launches."""
    from oron_tts_b200 import _lib as L

    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak, which = (peaks["bf16_tflops_sustained"], "measured (sustained cuBLAS bf16)") if "bf16_tflops_sustained" in peaks \
        else (1400.0, "fallback (B200_PROFILING.md sustained)")
    # (re)populate the config-2 workspace (the side measurements may have evicted or replaced it) and take it by key
    cfm.sample(ref_mel, ids, dur, lens=lens, steps=STEPS_NFE, cfg_strength=CFG, sway_sampling_coef=SWAY, seed=None)
    eng.deterministic = False  # the kernel families are timed as the headline legs run them
    ws = eng.workspace(1, 2, (T_TOTAL + 127) // 128 * 128, STEPS_NFE, True)
    kinds = ("gemm", "attention", "ln_modulate")
    orig = {name: getattr(L, name) for name in kinds}
    t, launches = {}, {}
    ws.step.zero_()
    for kind in kinds:
        # one CUDA graph holding ONLY this kernel family's launches of one NFE (same operands, same order, back to
        # back as in the real step): replay time / launches = average launch duration under the timed conditions.
        # Bracketing every launch with its own event pair instead costs several microseconds of bubble per kernel.
        count = [0]

        def counted(*a, _f=orig[kind], **k):
            count[0] += 1
            return _f(*a, **k)

        try:
            for name in kinds:
                setattr(L, name, counted if name == kind else (lambda *a, **k: None))
            eng.velocity(ws, mod_nb=1, use_step=True)  # warm (attributes, tables)
            torch.cuda.synchronize()
            count[0] = 0
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                eng.velocity(ws, mod_nb=1, use_step=True)
        finally:
            for name, fn in orig.items():
                setattr(L, name, fn)
        launches[kind] = count[0]
        for _ in range(2):
            g.replay()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t[kind] = e0.elapsed_time(e1) / reps
    fl = algorithmic_flops_per_nfe(T_TOTAL)
    # memory-bound family of the NFE: LayerNorm + AdaLN modulation, fp32 row in (4 B/elt) -> bf16 row out (2 B/elt).
    # Its operands (17 MB per launch) stay L2-resident between the producing GEMM and this kernel, so the figure is
    # an on-chip bandwidth; the HBM-streaming row-wise kernels are measured at config-4 size in `secondary`.
    hbm = peaks.get("hbm_gbs", 6650.0)
    ln_bytes = 6.0 * ws.nbp * ws.tpad * eng.w.dim
    ln_gbs = ln_bytes * launches["ln_modulate"] / (t["ln_modulate"] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum per launch of the FFN-up GEMM, from one ncu --set full capture
        traffic = json.load(open(tp)).get("gemm2_ffn_up_dram_bytes_per_launch")
    gemm_flops = fl["gemm"] + fl["other"]
    achieved = gemm_flops / (t["gemm"] * 1e-3) / 1e12
    total = sum(t.values())
    return {
        "bound": "tensor", "kernel": "gemm2_bf16_tcgen05_kernel (all Linear/conv launches of one NFE)", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
        "frac": round(achieved / peak, 4), "peak_source": which, "traffic": traffic,
        "traffic_source": "static: profiles/r01_ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one FFN-up launch, ncu --set full; not re-measured in this run)" if traffic is not None else None,
        "memory_bound": {"ln_modulate": {"bytes_per_launch": int(ln_bytes), "gbs": round(ln_gbs, 1), "frac_of_hbm_peak": round(ln_gbs / hbm, 3),
                                         "note": "operands L2-resident between kernels"}},
        "flops_per_nfe": fl, "kernel_ms_per_nfe": {k: round(v, 4) for k, v in t.items()},
        "share_of_step": {k: round(v / total, 4) for k, v in t.items()}, "launches_per_nfe": launches,
        "avg_launch_us": {k: round(1e3 * v / max(launches[k], 1), 2) for k, v in t.items()},
        "attention_tflops": round(fl["attn"] / (t["attention"] * 1e-3) / 1e12, 1),
    }


def secondary_cfg3(model, cfm, voc, dev, rank, world, n_utt, barrier, e2e: bool = True) -> dict:
    """BASELINE config 3 (not the headline): Base DiT batched synthesis of mixed-length utterances (durations
    ~ U(1, 30) s, seed 0 -> T = int(d * 93.75) frames, reference-free, 32 NFE, CFG 2.0), sharded across the ranks by
    the longest-first cost-model assignment (shard.assign_utterances, no data-path collective) and, inside a rank,
    packed into length-sorted batches of <= 8192 padded rows (shard.plan_batches). n_utt = 32 per GPU by default, i.e.
    exactly the named 256-utterance job at 8 GPUs. One untimed pass first (workspaces + CUDA graphs per batch shape)."""
    import random

    import torch.distributed as dist

    from oron_tts_b200.shard import assign_utterances, imbalance, padding_waste, plan_batches

    rnd = random.Random(0)
    frames = [int(rnd.uniform(1.0, 30.0) * 93.75) for _ in range(n_utt)]
    plan = assign_utterances(frames, world)
    mine = plan[rank]
    my_frames = [frames[i] for i in mine]
    batches = plan_batches(my_frames, max_rows=8192)
    g = torch.Generator().manual_seed(3)
    ids_all = [torch.randint(11, 65, (max(4, f // 8),), generator=g) for f in frames]

    def one_pass():
        outs = 0
        for b in batches:
            fr = [my_frames[j] for j in b]
            tmax, B = max(fr), len(b)
            ids = torch.full((B, tmax), -1, dtype=torch.long)
            for r, j in enumerate(b):
                src = ids_all[mine[j]]
                ids[r, : fr[r]] = src[(torch.arange(fr[r]) * src.numel() // fr[r])]
            mel, _ = cfm.sample(torch.zeros(B, tmax, 100, device=dev), ids, torch.tensor(fr),
                                lens=torch.zeros(B, dtype=torch.long), steps=STEPS_NFE, cfg_strength=CFG,
                                sway_sampling_coef=SWAY, seed=None)
            for r in range(B):
                outs += voc.decode(mel[r:r + 1, : fr[r]].transpose(1, 2)).shape[-1]
        return outs

    one_pass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    samples = one_pass()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    tot = torch.tensor([float(samples)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    audio_s = float(tot.item()) / 24000.0
    fl = sum(2 * f * (563.4e6 + 90112.0 * f) for f in frames) * STEPS_NFE
    if not e2e:
        return {"workload": f"cfg3: the named {n_utt} reference-free utterances, durations U(1,30) s (seed 0), 32 NFE, CFG 2.0, + Vocos; "
                            f"sharded over {world} GPU(s), length-sorted batches of <= 8192 padded rows",
                "utterances": n_utt, "audio_s": round(audio_s, 1), "ms": round(float(ms.item()), 1),
                "audio_s_per_s": round(audio_s / (float(ms.item()) / 1e3), 1),
                "algorithmic_tflops": round(fl / (float(ms.item()) * 1e-3) / 1e12, 1),
                "rank_imbalance": round(imbalance(frames, plan), 4), "padding_waste_rank0": round(padding_waste(my_frames, batches), 4),
                "batches_rank0": len(batches)}
    # the same job end to end through the public API: host strings in, host waveforms out (F5TTS.synthesize_batch)
    words = BENCH_TEXT.replace(",", "").replace(".", "").split()
    texts = [" ".join(words[(i + k) % len(words)] for k in range(3 + i % 9)) for i in mine]
    durs = [(frames[i] + 0.5) * 256 / 24000.0 for i in mine]  # int(d * 24000 / 256) == frames[i]
    kw = dict(lang="mn", n_steps=STEPS_NFE, cfg_strength=CFG, sway_sampling_coef=SWAY, target_durations_s=durs,
              seeds=None, max_chars_per_chunk=0, device=str(dev))
    model.synthesize_batch(texts, **kw)  # warm: the target lengths (hence batch shapes) equal those of the device leg
    barrier()
    e2 = torch.cuda.Event(enable_timing=True)
    e3 = torch.cuda.Event(enable_timing=True)
    e2.record()
    wavs = model.synthesize_batch(texts, **kw)
    e3.record()
    barrier()
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev)  # the call ends with the device-to-host copies: fully drained
    tot2 = torch.tensor([float(sum(w.numel() for w in wavs))], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot2, op=dist.ReduceOp.SUM)
    e2e_audio = float(tot2.item()) / 24000.0
    fl = sum(2 * f * (563.4e6 + 90112.0 * f) for f in frames) * STEPS_NFE
    return {"workload": f"cfg3: {n_utt} reference-free utterances, durations U(1,30) s (seed 0), 32 NFE, CFG 2.0, + Vocos; "
                        f"sharded over {world} GPU(s), length-sorted batches of <= 8192 padded rows",
            "utterances": n_utt, "audio_s": round(audio_s, 1), "ms": round(float(ms.item()), 1),
            "audio_s_per_s": round(audio_s / (float(ms.item()) / 1e3), 1),
            "e2e_audio_s_per_s": round(e2e_audio / (float(ms2.item()) / 1e3), 1), "e2e_ms": round(float(ms2.item()), 1),
            "algorithmic_tflops": round(fl / (float(ms.item()) * 1e-3) / 1e12, 1),
            "rank_imbalance": round(imbalance(frames, plan), 4), "padding_waste_rank0": round(padding_waste(my_frames, batches), 4),
            "batches_rank0": len(batches)}


def secondary_cfg5(model, dev, rank, world, barrier, steps: int = 3) -> dict:
    """BASELINE config 5 (not the headline): Base DiT OT-CFM training step -- CFM.forward (random t / span / noise / CFG
    drops, flow.py:101-138) + backward + NCCL gradient all-reduce + clip + AdamW (trainer.py:191-262) -- on the sm_100a
    kernels, bf16 tensor-core operands with fp32 master weights / gradients / moments. Per-GPU batch 8 x 1024 frames
    (runpod.yaml batch_size), data-parallel over the ranks (weak scaling). Algorithmic FLOPs = 3 x forward."""
    import torch.distributed as dist
    import weights as GW

    from oron_tts_b200.f5tts import F5TTS
    from oron_tts_b200.train import TrainEngine

    B, Tn = 8, 1024
    with torch.device(dev):
        tm = F5TTS.from_config(GW.CONFIGS["base"])
    tm.load_state_dict(model.state_dict(), strict=True)  # same seeded weights as the inference legs, separate storage
    tm = tm.to(dev).train()
    eng = TrainEngine(tm, lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01, max_grad_norm=1.0)
    g = torch.Generator(device=dev).manual_seed(500 + rank)
    mel = torch.randn(B, 100, Tn, device=dev, generator=g) * 1.5 - 3.0
    text = torch.randint(4, 65, (B, Tn), device=dev, generator=g)
    lens = torch.full((B,), Tn, device=dev, dtype=torch.long)
    losses = []
    for _ in range(2):
        losses.append(eng.train_step(mel, text, lens, lr=1e-4))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses.append(eng.train_step(mel, text, lens, lr=1e-4))
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    ls = [float(x) for x in losses]
    assert all(v == v and v < 1e6 for v in ls), ls
    fwd = B * Tn * (563.4e6 + 90112.0 * Tn)
    comm = ("no collective" if world == 1 else
            f"sharded optimizer: NCCL reduce-scatter of {eng.arena.numel * 4 / 1e9:.2f} GB fp32 gradients, AdamW on 1/{world} of the arenas, "
            f"all-gather of the {eng.arena.numel * 2 / 1e9:.2f} GB bf16 operand copy" if eng.sharded is not None else
            f"NCCL all-reduce of {eng.arena.numel * 4 / 1e9:.2f} GB fp32 gradients per step, per-block buckets overlapped with backward")
    out = {"workload": f"cfg5: Base DiT OT-CFM training step, per-GPU batch {B} x {Tn} frames, bf16 operands / fp32 master+grads, "
                       f"data-parallel x{world} ({comm})",
           "ms_per_step": round(ms, 2), "samples_per_s": round(world * B / (ms / 1e3), 2),
           "frames_per_s": round(world * B * Tn / (ms / 1e3), 1),
           "algorithmic_tflops_per_gpu": round(3 * fwd / (ms * 1e-3) / 1e12, 1), "loss_first_last": [round(ls[0], 4), round(ls[-1], 4)],
           "optimizer_steps_skipped": int(eng.skipped.item())}
    del eng, tm
    torch.cuda.empty_cache()
    return out


def secondary_cfg4(model, voc, dev) -> dict:
    """BASELINE config 4 (not the headline): log-mel STFT and Vocos decode on 64 x 30 s clips, inputs resident in HBM.
    Reports the Vocos real-time factor (decode seconds / audio seconds) and the achieved HBM bandwidth of the fused
    log-mel kernel against its algorithmic bytes (4 B/sample read + 400 B/frame written)."""
    nb, S = 64, 720000
    g = torch.Generator(device=dev).manual_seed(4)
    wav = (torch.rand(nb, S, device=dev, generator=g) * 2 - 1) * 0.3
    ap = model._audio_processor

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    ms_mel, mel = timeit(lambda: ap.mel_spectrogram(wav))
    frames = mel.shape[-1]
    mel_bytes = nb * S * 4 + nb * 100 * frames * 4
    ms_voc, w = timeit(lambda: voc.decode(mel), reps=3)
    audio_s = nb * w.shape[-1] / 24000.0
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    hbm = peaks.get("hbm_gbs", 6650.0)
    # the HBM-streaming row-wise kernels of the decode, each timed alone at this size (algorithmic bytes)
    from oron_tts_b200 import _lib as L

    pk = voc._pack()
    R, D = nb * frames, 512
    x = torch.randn(R, D, device=dev, dtype=torch.float32)
    n_ = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    blk = pk["blocks"][0]
    ms_dw, _ = timeit(lambda: L.dwconv7_ln(x, rows_per_batch=frames, nbatch=nb, seq_lens=None, w=blk["dw_w"], wb=blk["dw_b"],
                                           ln_w=blk["ln_w"], ln_b=blk["ln_b"], eps=1e-6, out=n_))
    dw_bytes = R * D * 6
    nh = pk["head_w"].shape[0]
    hs = torch.randn(R, (nh + 31) // 32 * 32, device=dev, dtype=torch.float32) * 0.5
    wv = torch.empty(nb, (frames - 1) * 256, device=dev, dtype=torch.float32)
    ms_is, _ = timeit(lambda: L.istft_head(hs, pk["window"], wv, rows_per_batch=frames, nb=nb, n_frames=frames, mode=0))
    is_bytes = R * nh * 4 + nb * (frames - 1) * 256 * 4
    del x, n_, hs, wv
    return {
        "workload": "cfg4: 64 x 30 s clips (24 kHz, n_fft 1024, hop 256, 100 mels)",
        "dwconv7_ln_ms": round(ms_dw, 3), "dwconv7_ln_gbs": round(dw_bytes / ms_dw / 1e6, 1), "dwconv7_ln_hbm_frac": round(dw_bytes / ms_dw / 1e6 / hbm, 4),
        "istft_head_ms": round(ms_is, 3), "istft_head_gbs": round(is_bytes / ms_is / 1e6, 1), "istft_head_hbm_frac": round(is_bytes / ms_is / 1e6 / hbm, 4),
        "logmel_roofline_class": "fp32 issue: a 1024-point FFT per frame pair is ~2.9 k instructions per lane against 4 KB of HBM traffic (floor ~0.125 ms at this size, HBM floor 0.039 ms)",
        "logmel_ms": round(ms_mel, 3), "logmel_gbs": round(mel_bytes / ms_mel / 1e6, 1), "logmel_hbm_frac": round(mel_bytes / ms_mel / 1e6 / hbm, 4),
        "vocos_ms": round(ms_voc, 3), "vocos_rtf": round(ms_voc / 1e3 / audio_s, 7), "vocos_audio_s_per_s": round(audio_s / (ms_voc / 1e3), 1),
        "vocos_tflops": round(nb * frames * 27.0e6 / (ms_voc * 1e-3) / 1e12, 1),
    }


def secondary_cfg1(voc, dev) -> dict:
    """BASELINE config 1: Small DiT (dim 512, depth 12, heads 8, text_dim 256), reference-free "Сайн байна уу" (T = 143),
    32 Euler steps, CFG 1.5, batch 1 -- the reference's own CPU-runnable case, here on the GPU path."""
    import weights as GW

    from oron_tts_b200.f5tts import F5TTS

    m = F5TTS.from_config(GW.CONFIGS["small"])
    m.load_state_dict(GW.fill_state_dict(m.state_dict(), GW.SEEDS["small"]), strict=True)
    m = m.to(dev).eval()
    m.set_vocoder(voc)
    ids = torch.tensor([[4, 30, 11, 21, 25, 53, 12, 11, 21, 25, 11, 53, 32, 32]])  # host-side metadata, as in synthesize
    cond, lens = torch.zeros(1, 143, 100, device=dev), torch.tensor([0])

    # unseeded, like the headline legs: a seeded call selects the bit-reproducible mode (no stream-K in the FFN
    # down-projection), which costs ~16 % at this launch-bound size and is reported for config 2 under "deterministic"
    def dev_step(seed):
        mel, _ = m.cfm.sample(cond, ids, 143, lens=lens, steps=32, cfg_strength=1.5, sway_sampling_coef=SWAY)
        return voc.decode(mel.transpose(1, 2))

    def e2e_step(seed):
        return m.synthesize("Сайн байна уу", lang="mn", n_steps=32, cfg_strength=1.5, sway_sampling_coef=SWAY, device=str(dev))

    res = {}
    for name, fn in (("device", dev_step), ("e2e", e2e_step)):
        for i in range(3):
            w = fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            w = fn(10 + i)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
        assert w.shape[-1] == (143 - 1) * 256
    audio = (143 - 1) * 256 / 24000.0
    fl = 2 * 143 * (78.1e6 + 24576.0 * 143)
    return {"workload": "cfg1: Small DiT, reference-free 'Сайн байна уу' (T = 143), 32 NFE, CFG 1.5, batch 1, + Vocos",
            "ms_per_utterance": round(res["device"], 3), "ms_per_nfe": round(res["device"] / 32, 4),
            "audio_s_per_s": round(audio / (res["device"] / 1e3), 1), "e2e_audio_s_per_s": round(audio / (res["e2e"] / 1e3), 1),
            "algorithmic_tflops": round(fl * 32 / (res["device"] * 1e-3) / 1e12, 2),
            "note": "23 GFLOP per NFE on 286 rows: launch-latency bound (CUDA-graph replay of ~200 kernels per NFE)"}


# ----------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def import_reference():
    """The unmodified reference package from baseline/_ref (tools/make_baseline_ref.py) as `src`, or None. Must run before
    anything imports this repository's own `src` shim (bench.py never does: it uses oron_tts_b200 directly)."""
    if not os.path.isdir(os.path.join(REF_DIR, "src")):
        return None
    if "src" in sys.modules and not os.path.abspath(getattr(sys.modules["src"], "__file__", "") or "").startswith(REF_DIR):
        return None
    for pth in (os.path.join(REF_DIR, "_stubs"), REF_DIR):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    try:
        from src.models.f5tts import F5TTS as RefF5TTS  # noqa: PLC0415

        return RefF5TTS
    except Exception as e:  # pragma: no cover
        print(f"[bench] reference import failed: {type(e).__name__}: {e}", file=sys.stderr)
        return None


def reference_model(name: str, device: str):
    """Reference F5TTS with the same seeded weights as this repo's legs (tests/golden/weights.py)."""
    import weights as GW

    RefF5TTS = import_reference()
    if RefF5TTS is None:
        return None
    m = RefF5TTS.from_config(GW.CONFIGS[name])
    m.load_state_dict(GW.fill_state_dict(m.state_dict(), GW.SEEDS[name]), strict=True)
    return m.to(device).eval()


def cfg2_inputs(device, seed: int = 100):
    g = torch.Generator().manual_seed(seed)
    ref_mel = torch.randn(1, REF_LEN, 100, generator=g) * 1.5 - 3.0
    ids = torch.randint(11, 65, (1, T_TOTAL), generator=g)
    return ref_mel.to(device), ids.to(device), torch.tensor([T_TOTAL], device=device), torch.tensor([REF_LEN], device=device)


def cpu_baseline(nfe: int = 4, reps: int = 3) -> dict:
    """The reference on the host cores, fp32 eager, all torch threads: `reps` samples of `nfe` CFG NFE steps of config 2
    through the reference's own CFM.sample (best sample kept: one sample is noisy, 1.4-2.4 s per NFE on 16 cores), extrapolated
    to 32 NFE. baseline/_ref absent: the oracle port of the same algorithm (kind "port"). The vocoder is excluded in both
    (the `vocos` package the reference calls is not installed here); it is <1 % of an utterance."""
    ref = reference_model("base", "cpu")
    if ref is not None:
        ref_mel, ids, dur, lens = cfg2_inputs("cpu")
        cfm = ref.cfm
        with torch.inference_mode():
            cfm.sample(ref_mel[:, :64], ids[:, :128], 128, lens=torch.tensor([64]), steps=1, cfg_strength=CFG, sway_sampling_coef=SWAY, seed=0)  # warm
            ts = []
            for r in range(reps):
                t0 = time.perf_counter()
                cfm.sample(ref_mel, ids, dur, lens=lens, steps=nfe, cfg_strength=CFG, sway_sampling_coef=SWAY, seed=r)
                ts.append((time.perf_counter() - t0) / nfe)
        dt, kind = min(ts), "reference"
        what = f"best of {reps} x {nfe}-NFE CFM.sample calls of cfg2 (T=1406, rows 2812) on the unmodified reference (baseline/_ref), x{STEPS_NFE // nfe} to one utterance; Vocos excluded"
    else:
        import weights as GW

        from oracle import dit_oracle as DO
        from oron_tts_b200.f5tts import F5TTS

        with torch.device("meta"):
            proto = F5TTS.from_config(GW.CONFIGS["base"]).state_dict()
        sd = GW.fill_state_dict({k: torch.empty(v.shape) for k, v in proto.items()}, GW.SEEDS["base"])
        sd["cfm.backbone.rotary_embed.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
        g = torch.Generator().manual_seed(100)
        x = torch.randn(1, T_TOTAL, 100, generator=g)
        cond = torch.zeros(1, T_TOTAL, 100)
        cond[:, :REF_LEN] = torch.randn(1, REF_LEN, 100, generator=g) * 1.5 - 3.0
        ids = torch.randint(11, 65, (1, T_TOTAL), generator=g)
        mask = torch.ones(1, T_TOTAL, dtype=torch.bool)
        cache: dict = {}
        ts = []
        with torch.inference_mode():
            DO.dit_forward(sd, x[:, :256], cond[:, :256], ids[:, :256], torch.tensor([0.1]), mask[:, :256], cfg_infer=True)  # warm
            for r in range(reps):
                t0 = time.perf_counter()
                for i in range(nfe):
                    DO.dit_forward(sd, x, cond, ids, torch.tensor([0.1 * (i + 1)]), mask, cfg_infer=True, text_cache=cache)
                ts.append((time.perf_counter() - t0) / nfe)
        dt, kind = min(ts), "port"
        what = f"best of {reps} x {nfe} CFG NFE of cfg2 (T=1406, rows 2812) on the oracle port (baseline/_ref absent), x{STEPS_NFE // nfe} to one utterance"
    return {"value": round(AUDIO_S / (dt * STEPS_NFE), 5), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "s_per_nfe": round(dt, 3), "s_per_nfe_samples": [round(t, 3) for t in ts], "host_cpus": os.cpu_count(), "sample": what}


def cpu_side_figures() -> dict:
    """CPU figures of the other configs on the unmodified reference (SURVEY 8d "CPU baseline timing"), bounded samples:
    cfg 1 full 32-NFE CFM.sample (Small); cfg 4 log-mel on 8 of the 64 clips and the in-repo VocosDecoder on 2 clips
    (scaled, stated); cfg 5 one fwd + bwd + AdamW step at B = 1, T = 1024 (Base)."""
    out: dict = {}
    small = reference_model("small", "cpu")
    if small is None:
        return {"unavailable": "baseline/_ref absent"}
    ids = torch.tensor([[4, 30, 11, 21, 25, 53, 12, 11, 21, 25, 11, 53, 32, 32]])
    with torch.inference_mode():
        ts = []
        for r in range(3):
            t0 = time.perf_counter()
            small.cfm.sample(torch.zeros(1, 143, 100), ids, 143, lens=torch.tensor([0]), steps=32, cfg_strength=1.5, sway_sampling_coef=SWAY, seed=r)
            ts.append(time.perf_counter() - t0)
    out["cfg1"] = {"s_per_utterance": round(min(ts), 3), "ms_per_nfe": round(min(ts) / 32 * 1e3, 1),
                   "audio_s_per_s": round((143 - 1) * 256 / 24000.0 / min(ts), 3), "sample": "best of 3 full 32-NFE CFM.sample calls (Small, T = 143)"}
    del small
    from src.models.decoder import VocosDecoder  # noqa: PLC0415  (the reference's in-repo analogue of pretrained Vocos)
    from src.utils.audio import AudioProcessor as RefAP  # noqa: PLC0415

    ap_ref = RefAP()
    g = torch.Generator().manual_seed(4)
    wav = (torch.rand(8, 720000, generator=g) * 2 - 1) * 0.3
    with torch.inference_mode():
        ap_ref.mel_spectrogram(wav[:1])
        t0 = time.perf_counter()
        mel = ap_ref.mel_spectrogram(wav)
        t_mel = (time.perf_counter() - t0) * 8
        dec = VocosDecoder().eval()
        dec(mel[:1, :, :64])
        t0 = time.perf_counter()
        dec(mel[:2])
        t_voc = (time.perf_counter() - t0) * 32
    out["cfg4"] = {"logmel_ms": round(t_mel * 1e3, 1), "vocos_ms": round(t_voc * 1e3, 1), "vocos_rtf": round(t_voc / (64 * 719872 / 24000.0), 6),
                   "sample": "log-mel on 8 of the 64 clips x8; the reference's VocosDecoder (decoder.py, architecture analogue) on 2 clips x32"}
    base = reference_model("base", "cpu")
    base.train()
    opt = torch.optim.AdamW(base.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01)
    g = torch.Generator().manual_seed(500)
    mel1 = torch.randn(1, 100, 1024, generator=g) * 1.5 - 3.0
    text1 = torch.randint(4, 65, (1, 1024), generator=g)
    t0 = time.perf_counter()
    loss = base(mel1, text1, torch.tensor([1024]))
    loss.backward()
    torch.nn.utils.clip_grad_norm_(base.parameters(), 1.0)
    opt.step()
    dt = time.perf_counter() - t0
    out["cfg5"] = {"s_per_step_b1": round(dt, 2), "samples_per_s": round(1.0 / dt, 4),
                   "sample": "one fwd + bwd + clip + AdamW step, B = 1, T = 1024, single process (no DDP), first (cold) step"}
    return out


def run_reference(args) -> dict:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return {}
    world = int(os.environ.get("WORLD_SIZE", 1))
    # all host threads (torchrun exports OMP_NUM_THREADS=1 for its workers)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    # a "step" of this arm = one 4-NFE CFM.sample call (about 6-9 s on 16 cores); steps + warm-up are bounded to 6 calls
    cb = cpu_baseline(nfe=4, reps=max(3, min(args.steps + min(args.warmup, 1), 6)))
    return {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(cb["s_per_nfe"] * STEPS_NFE * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: Base DiT (dim 1024, depth 22), 469 ref + 937 target frames, 32 NFE, CFG 2.0, sway -1 "
                               f"(host cores, {cb['kind']}: {cb['sample']})"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference_gpu(args) -> dict:
    """The reference on the SAME B200 in PyTorch (SURVEY 8d / BASELINE.md section 3: "the real bar to beat"): config 2 through
    the reference's own CFM.sample, (a) eager fp32 -- what scripts/infer.py runs --, (b) eager under bf16 autocast, (c) with
    torch.compile(backbone, dynamic=True) as scripts/train.py:206-214 sets it up, under bf16 autocast. cuBLAS / SDPA / Inductor
    do the work here: none of this repository's kernels are on this path. Vocos is excluded (package not installed)."""
    dev = "cuda:0"
    out: dict = {"workload": "cfg2 through the reference's CFM.sample (32 NFE, CFG 2.0, sway -1), DiT only; best of 3 after 1 warm call",
                 "torch": torch.__version__}
    ref = reference_model("base", dev)
    if ref is None:
        return {"unavailable": "baseline/_ref absent"}
    ref_mel, ids, dur, lens = cfg2_inputs(dev)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def sample():
        return ref.cfm.sample(ref_mel, ids, dur, lens=lens, steps=STEPS_NFE, cfg_strength=CFG, sway_sampling_coef=SWAY, seed=1)

    def entry(ms):
        return {"ms_per_utterance": round(ms, 2), "ms_per_nfe": round(ms / STEPS_NFE, 3), "audio_s_per_s": round(AUDIO_S / (ms / 1e3), 2)}

    out["eager_fp32"] = entry(timed(sample))
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    out["eager_tf32"] = entry(timed(sample))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def sample_bf16():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return sample()

    out["eager_bf16_autocast"] = entry(timed(sample_bf16))
    if not args.no_compile:
        try:
            t0 = time.perf_counter()
            ref.cfm.backbone = torch.compile(ref.cfm.backbone, dynamic=True)
            ms = timed(sample_bf16)
            out["compiled_bf16_autocast"] = dict(entry(ms), compile_s=round(time.perf_counter() - t0 - 4 * ms / 1e3, 1))
        except Exception as e:  # pragma: no cover
            out["compiled_bf16_autocast"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    return out


def gpu_eager_baseline(timeout_s: int = 420) -> dict:
    """run_reference_gpu in a child process (its own CUDA context, bounded time: Inductor compilation of a 22-block DiT)."""
    import subprocess

    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-gpu"]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env, cwd=ROOT)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"unavailable": f"rc {r.returncode}: {(r.stderr or r.stdout)[-300:]}"}
    except subprocess.TimeoutExpired:
        try:  # eager numbers only
            r = subprocess.run(cmd + ["--no-compile"], capture_output=True, text=True, timeout=180, env=env, cwd=ROOT)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and lines:
                d = json.loads(lines[-1])
                d["compiled_bf16_autocast"] = {"unavailable": f"torch.compile did not finish within {timeout_s} s"}
                return d
        except Exception:  # pragma: no cover
            pass
        return {"unavailable": f"timed out after {timeout_s} s"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--no-compile", action="store_true", help="reference-gpu: skip the torch.compile variant")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-on-GPU (PyTorch eager / compile) block")
    ap.add_argument("--no-cfg3-strong", action="store_true", help="skip the 256-utterance strong-scaling leg of config 3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg3-utterances", type=int, default=0, help="utterances of the config-3 side measurement (default 32 per GPU)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-4 (log-mel / Vocos) side measurement")
    a = ap.parse_args()
    res = run_reference(a) if a.impl == "reference" else run_reference_gpu(a) if a.impl == "reference-gpu" else run_ours(a)
    if res:
        print(json.dumps(res, ensure_ascii=False), flush=True)
