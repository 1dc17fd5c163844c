"""CFM sampler (Euler ODE + classifier-free guidance + sway schedule) on the sm_100a engine.

Drop-in for src/models/flow.py:49-306: same constructor, ``sample`` signature, return value
(``(mel [B, T, n_mels], trajectory list of steps+1 tensors)``) and ValueError messages. The ODE loop
itself is one CUDA graph replayed ``steps`` times: text embedding, the step-invariant part of the
input projection and all AdaLN modulation vectors are hoisted out of the loop, and the conditional /
unconditional passes are batched in every kernel.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .dit import DiT
from .engine import Branch, TILE, _rup


def _lens_to_mask(lens: torch.Tensor, length: int | None = None) -> torch.Tensor:
    """Boolean [B, N] prefix mask from lengths (flow.py:22-27)."""
    n = int(lens.amax().item()) if length is None else length
    return torch.arange(n, device=lens.device)[None, :] < lens[:, None]


class CFM(nn.Module):
    def __init__(self, backbone: DiT, sigma: float = 0.0, audio_drop_prob: float = 0.3, cond_drop_prob: float = 0.2,
                 frac_lengths_mask: tuple[float, float] = (0.7, 1.0), n_mels: int = 100) -> None:
        super().__init__()
        self.backbone = backbone
        self.sigma = sigma
        self.audio_drop_prob = audio_drop_prob
        self.cond_drop_prob = cond_drop_prob
        self.frac_lengths_mask = frac_lengths_mask
        self.n_mels = n_mels

    # ---- training objective (flow.py:69-159) -----------------------------------------------------
    def forward(self, inp: torch.Tensor, text_ids: torch.Tensor, *, lens: torch.Tensor | None = None) -> torch.Tensor:
        """OT-CFM loss (flow.py:69-159). In training mode with autograd enabled the loss is an autograd node backed by
        the sm_100a training engine (oron_tts_b200/train.py): ``loss.backward()`` delivers the parameter gradients
        exactly as the reference's autograd graph would. Eval mode: the deterministic validation objective."""
        if self.training and torch.is_grad_enabled():
            from .train import OTCFMLoss, TrainEngine

            eng = self.__dict__.get("_train_engine")
            if eng is None:
                eng = TrainEngine(self)
                self.__dict__["_train_engine"] = eng
            params = [eng.arena.named[k] for k in eng.arena.order]
            return OTCFMLoss.apply(eng, inp, text_ids, lens, *params)
        if self.training:
            raise RuntimeError("CFM.forward in training mode needs autograd enabled (or call .eval() for the validation loss)")
        if inp.ndim == 3 and inp.shape[1] == self.n_mels:
            inp = inp.transpose(1, 2)
        B, T, dev = inp.shape[0], inp.shape[1], inp.device
        if lens is None:
            lens = torch.full((B,), T, device=dev, dtype=torch.long)
        mask = _lens_to_mask(lens, length=T)
        mid = sum(self.frac_lengths_mask) / 2
        span = (torch.full((B,), mid, device=dev).float() * lens).long()
        start = ((lens - span) // 2).clamp(min=0)
        pos = torch.arange(T, device=dev)
        span_mask = (pos[None, :] >= start[:, None]) & (pos[None, :] < (start + span)[:, None]) & mask
        time = torch.full((B,), 0.5, dtype=inp.dtype, device=dev)
        x1 = inp
        cond = torch.where(span_mask[..., None], torch.zeros_like(x1), x1)
        gen = torch.Generator(device=dev).manual_seed(0)
        x0 = torch.randn(x1.shape, generator=gen, device=dev, dtype=inp.dtype)
        t = time[:, None, None]
        phi = (1 - t) * x0 + t * x1
        pred = self.backbone(x=phi, cond=cond, text=text_ids, time=time, mask=mask)
        return F.mse_loss(pred, x1 - x0, reduction="none")[span_mask].mean()

    # ---- sampling (flow.py:161-306) -----------------------------------------------------------------
    @torch.inference_mode()
    def sample(self, cond: torch.Tensor, text_ids: torch.Tensor, duration: torch.Tensor | int, *,
               lens: torch.Tensor | None = None, steps: int = 32, cfg_strength: float = 1.0,
               sway_sampling_coef: float | None = None, seed: int | None = None, max_duration: int = 65536,
               y0: torch.Tensor | None = None, method: str = "euler",
               precision: str = "bf16", deterministic: bool | None = None) -> tuple[torch.Tensor, list[torch.Tensor]]:
        """``y0`` (extra, optional): inject the initial noise [B, max_dur, n_mels] instead of drawing it —
        needed for cross-device parity because CPU and CUDA generators produce different streams.
        ``method`` (extra): "euler" (the reference, flow.py:290-299) or "midpoint" (explicit midpoint rule on the same
        schedule: two DiT evaluations per step, second-order accurate).
        ``deterministic`` (extra): bit-reproducible run to run, like the reference's seeded sample. Default
        (None): on whenever the noise is pinned (``seed`` or ``y0`` given), off for unseeded calls -- whose noise differs
        from call to call anyway -- which then use the stream-K split of the FFN down-projection (about 4 % faster at
        config 2; its fp32 partial sums land in arrival order, i.e. differences in the last bits)."""
        if deterministic is None:
            deterministic = seed is not None or y0 is not None
        if method not in ("euler", "midpoint"):
            raise ValueError(f"method must be 'euler' or 'midpoint', got {method!r}")
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        evals = 2 if method == "midpoint" else 1
        if steps < 1:
            raise ValueError(f"steps must be >= 1, got {steps}")
        if cfg_strength < 0:
            raise ValueError(f"cfg_strength must be >= 0, got {cfg_strength}")
        if self.training:  # eval() is a sweep over every submodule (host milliseconds per call with the GPU idle)
            self.eval()
        batch, cond_seq_len, device = cond.shape[0], cond.shape[1], cond.device
        if not cond.is_cuda:
            raise RuntimeError("CFM.sample runs only on a CUDA device (oron_tts_b200 has no CPU fallback)")

        # Host copies of duration / lens for the reference's checks (flow.py:219-230) and the workspace shape. Reading a CUDA
        # tensor back is a host sync that waits for everything already enqueued (the previous utterance's whole ODE loop when
        # calls are pipelined): ints and CPU tensors (what F5TTS.synthesize passes) need none, CUDA tensors cost one.
        if lens is not None and lens.numel() != batch:
            raise ValueError(f"lens must have {batch} values, got {lens.numel()}")
        if not isinstance(duration, int) and duration.numel() != batch:
            raise ValueError(f"duration must have {batch} values, got {duration.numel()}")
        dur_h = [int(duration)] * batch if isinstance(duration, int) else None
        lens_h = [cond_seq_len] * batch if lens is None else None
        if dur_h is None and not duration.is_cuda:
            dur_h = [int(v) for v in duration.reshape(-1).tolist()]
        if lens_h is None and not lens.is_cuda:
            lens_h = [int(v) for v in lens.reshape(-1).tolist()]
        if dur_h is None or lens_h is None:  # at least one CUDA tensor: one read-back for both
            dev_vals = [v.to(device=device, dtype=torch.long).reshape(-1) for v, h in ((duration, dur_h), (lens, lens_h)) if h is None]
            host = torch.stack(dev_vals).tolist()
            if dur_h is None:
                dur_h = host.pop(0)
            if lens_h is None:
                lens_h = host.pop(0)
        lens = torch.tensor(lens_h, dtype=torch.long).to(device, non_blocking=True) if lens is None or not lens.is_cuda \
            else lens.to(device=device, dtype=torch.long)
        if any(d <= 0 for d in dur_h):
            raise ValueError("duration values must be > 0")
        if any(v < 0 for v in lens_h):
            raise ValueError("lens values must be >= 0")
        if any(v > d for v, d in zip(lens_h, dur_h)):
            raise ValueError("conditioning lens must be <= duration for every sample")
        if any(d > max_duration for d in dur_h):
            raise ValueError(f"duration exceeds max_duration={max_duration}")
        max_dur = max(dur_h)
        if cond_seq_len > max_dur:
            raise ValueError("conditioning sequence length must be <= max duration")

        cond = cond.to(torch.float32)
        cond_mask = _lens_to_mask(lens)
        cond = F.pad(cond, (0, 0, 0, max_dur - cond_seq_len), value=0.0)
        cond_mask = F.pad(cond_mask, (0, max_dur - cond_mask.shape[-1]), value=False)
        cond_mask_3d = cond_mask.unsqueeze(-1)
        step_cond = torch.where(cond_mask_3d, cond, torch.zeros_like(cond))

        if precision == "fp32":
            return self._sample_fp32(cond, cond_mask_3d, step_cond, text_ids.to(device), dur_h, max_dur, steps, cfg_strength,
                                     sway_sampling_coef, seed, y0, method)
        eng = self.backbone.engine()
        w = eng.w
        use_cfg = cfg_strength >= 1e-5
        branches = [Branch(False, False), Branch(True, True)] if use_cfg else [Branch(False, False)]
        tpad = _rup(max_dur, TILE)
        ws = eng.workspace(batch, batch * len(branches), tpad, steps * evals, True)

        # initial noise: per-sample sequential draws from one generator, zero padded (flow.py:270-283)
        if y0 is None:
            generator = None
            if seed is not None:
                generator = torch.Generator(device=device).manual_seed(seed)
            ys = [torch.randn(d, self.n_mels, device=device, dtype=step_cond.dtype, generator=generator) for d in dur_h]
            y0 = torch.nn.utils.rnn.pad_sequence(ys, padding_value=0.0, batch_first=True)
        else:
            y0 = y0.to(device=device, dtype=torch.float32)

        # sway-sampled schedule (flow.py:286-288)
        t = torch.linspace(0, 1, steps + 1, device=device, dtype=step_cond.dtype)
        if sway_sampling_coef is not None:
            t = t + sway_sampling_coef * (torch.cos(torch.pi / 2 * t) - 1 + t)
        ws.dt[:steps].copy_(t[1:] - t[:-1])

        try:
            eng.load_sequences(ws, text=text_ids, durations=dur_h, seq_len=max_dur, branches=branches)
            eng.text_embed(ws)
            eng.static_embed(ws, step_cond, branches)
            # one modulation row per velocity evaluation: t_i, and t_i + dt_i / 2 for the midpoint rule
            t_eval = t[:-1] if evals == 1 else torch.stack([t[:-1], 0.5 * (t[:-1] + t[1:])], dim=1).reshape(-1)
            eng.modulation_table(ws, t_eval)
            xv = ws.x.view(batch, tpad, self.n_mels)
            xv.zero_()
            xv[:, :max_dur].copy_(y0)
            from . import _lib as L
            L.cast_rows_bf16(ws.x, ws.xb[:, : self.n_mels], reps=len(branches))
            ws.traj[0].copy_(ws.x)
            ws.step.zero_()
            eng.deterministic = bool(deterministic)
            eng.run_ode(ws, steps=steps * evals, cfg=float(cfg_strength), has_uncond=use_cfg, method=evals - 1)
        finally:
            self.backbone.clear_cache()

        tr = ws.traj.view(steps * evals + 1, batch, tpad, self.n_mels)[: steps + 1, :, :max_dur]
        trajectory = list(tr.clone().unbind(0))  # one copy out of the (reused) workspace instead of steps + 1
        out = torch.where(cond_mask_3d, cond, trajectory[-1])
        return out, trajectory

    def _sample_fp32(self, cond, cond_mask_3d, step_cond, text_ids, dur_h, max_dur, steps, cfg_strength, sway, seed, y0, method):
        """``sample(precision="fp32")``: the reference loop (flow.py:244-306) around the fp32-mode DiT forward
        (precise.PreciseDiT). A parity / debugging path: no CUDA graph, text embedding recomputed per evaluation; the
        three-term state update per step is plain torch on [B, T, n_mels]."""
        device, batch = cond.device, cond.shape[0]
        pd = self.backbone.precise()
        if y0 is None:
            generator = None if seed is None else torch.Generator(device=device).manual_seed(seed)
            ys = [torch.randn(d, self.n_mels, device=device, dtype=step_cond.dtype, generator=generator) for d in dur_h]
            y0 = torch.nn.utils.rnn.pad_sequence(ys, padding_value=0.0, batch_first=True)
        else:
            y0 = y0.to(device=device, dtype=torch.float32)
        t = torch.linspace(0, 1, steps + 1, device=device, dtype=step_cond.dtype)
        if sway is not None:
            t = t + sway * (torch.cos(torch.pi / 2 * t) - 1 + t)
        mask = torch.arange(max_dur, device=device)[None, :] < torch.tensor(dur_h, device=device)[:, None]

        def velocity(x, tt):
            tb = tt.expand(batch)
            if cfg_strength < 1e-5:
                return pd.forward(x, step_cond, text_ids, tb, mask=mask)
            both = pd.forward(x, step_cond, text_ids, tb, mask=mask, cfg_infer=True)
            return both[:batch] + (both[:batch] - both[batch:]) * cfg_strength

        x = y0
        trajectory = [y0]
        for i in range(steps):
            dt = t[i + 1] - t[i]
            v = velocity(x, t[i])
            if method == "midpoint":
                v = velocity(x + v * (0.5 * dt), 0.5 * (t[i] + t[i + 1]))
            x = x + v * dt
            trajectory.append(x)
        self.backbone.clear_cache()
        return torch.where(cond_mask_3d, cond, x), trajectory
