"""Multi-rank host logic on CPU (gloo, world_size 2): the sharding plan is identical on every rank,
covers every utterance exactly once without any data-path collective, and the max-over-ranks timing
reduction used by bench.py works."""

import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oron_tts_b200.shard import assign_utterances, imbalance


def test_assignment_covers_everything_and_balances():
    rng = random.Random(0)
    frames = [int(rng.uniform(1, 30) * 93.75) for _ in range(256)]  # BASELINE config 3 length draw
    for world in (1, 2, 4, 8):
        plan = assign_utterances(frames, world)
        flat = sorted(i for part in plan for i in part)
        assert flat == list(range(256))
        assert imbalance(frames, plan) < 1.02
    assert assign_utterances([], 4) == [[], [], [], []]
    with pytest.raises(ValueError):
        assign_utterances([10], 0)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(0)
    frames = [int(rng.uniform(1, 30) * 93.75) for _ in range(64)]
    plan = assign_utterances(frames, world)
    mine = plan[rank]
    # each rank "synthesises" its own utterances: no collective involved; host-side gather of the results
    results = {i: frames[i] * 256 for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, (plan, results))
    ms = torch.tensor([10.0 + rank])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((gathered, float(ms)))
    dist.destroy_process_group()


def test_plan_batches_rows_budget_and_cover():
    from oron_tts_b200.shard import padding_waste, plan_batches

    import random

    rnd = random.Random(0)
    frames = [int(rnd.uniform(1, 30) * 93.75) for _ in range(256)]  # BASELINE config 3 lengths
    plan = plan_batches(frames, max_rows=8192)
    assert sorted(i for b in plan for i in b) == list(range(256))
    for b in plan:
        tpad = (max(frames[i] for i in b) + 127) // 128 * 128
        assert len(b) == 1 or len(b) * tpad <= 8192
        assert frames[b[0]] == max(frames[i] for i in b)
    assert padding_waste(frames, plan) < 1.12
    assert plan_batches([5000], max_rows=1024) == [[0]]
    assert plan_batches([], max_rows=1024) == []
    import pytest

    with pytest.raises(ValueError):
        plan_batches([0], max_rows=1024)


def test_two_rank_plan_agreement_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gathered[0][0] == gathered[1][0]  # identical plan on both ranks
    merged = {}
    for _, res in gathered:
        assert not (set(res) & set(merged))
        merged.update(res)
    assert sorted(merged) == list(range(64))
    assert ms == 11.0  # max over ranks
