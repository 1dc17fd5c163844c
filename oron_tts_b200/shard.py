"""Utterance sharding across the GPUs of one box: the inference path has independent
units (utterances / text chunks) and therefore NO data-path collective. Every rank computes the same
deterministic assignment locally; results are gathered on the host side only.

Cost model per utterance with T frames: linear GEMM work plus quadratic attention work,
cost(T) = T * (563.4e6 + 90112 * T) FLOP per NFE and CFG branch.
"""

from __future__ import annotations


def utterance_cost(frames: int) -> float:
    return frames * (563.4e6 + 90112.0 * frames)


def assign_utterances(frames: list[int], world_size: int) -> list[list[int]]:
    """Greedy longest-processing-time-first bin packing; returns per-rank lists of utterance indices.

    Deterministic (ties broken by index) so that every rank derives the same plan without communicating.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(frames)), key=lambda i: (-utterance_cost(frames[i]), i))
    load = [0.0] * world_size
    plan: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += utterance_cost(frames[i])
    return plan


def imbalance(frames: list[int], plan: list[list[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = [sum(utterance_cost(frames[i]) for i in part) for part in plan]
    mean = sum(loads) / len(loads)
    return max(loads) / mean if mean > 0 else 1.0


def plan_batches(frames: list[int], max_rows: int = 8192, tile: int = 128) -> list[list[int]]:
    """Length-sorted batches for one GPU: every batch is padded to ceil(longest / tile) * tile rows per utterance
    (engine layout), so utterances are taken longest first and a batch is closed when one more
    utterance would exceed ``max_rows`` padded rows. Returns lists of indices into ``frames``; an utterance longer
    than ``max_rows`` gets a batch of its own."""
    if max_rows < 1:
        raise ValueError("max_rows must be >= 1")
    order = sorted(range(len(frames)), key=lambda i: (-frames[i], i))
    batches: list[list[int]] = []
    cur: list[int] = []
    tpad = 0
    for i in order:
        if frames[i] <= 0:
            raise ValueError("frame counts must be > 0")
        if not cur:
            cur, tpad = [i], (frames[i] + tile - 1) // tile * tile
        elif (len(cur) + 1) * tpad <= max_rows:
            cur.append(i)
        else:
            batches.append(cur)
            cur, tpad = [i], (frames[i] + tile - 1) // tile * tile
    if cur:
        batches.append(cur)
    return batches


def padding_waste(frames: list[int], batches: list[list[int]], tile: int = 128) -> float:
    """padded rows / valid rows over a batch plan (1.0 = no padding)."""
    padded = sum(len(b) * ((max(frames[i] for i in b) + tile - 1) // tile * tile) for b in batches)
    return padded / max(1, sum(frames[i] for b in batches for i in b))
