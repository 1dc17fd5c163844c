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
