"""CPU, world_size 2 over gloo: the N>1 host logic — round-robin scene sharding, fixed-shape gather, restoring scene
order. The per-frame-pair "result" is a deterministic stand-in (the CUDA path cannot run on the CPU box); the GPU
scaling run exercises the same functions over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shasta_b200 import sharding


def test_round_robin_assignment_is_a_partition():
    lengths = [3, 1, 4, 1, 5, 9, 2]
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            fps = sharding.frame_pairs_for_rank(lengths, world, r)
            # scene-major, time order inside a scene
            assert fps == sorted(fps, key=lambda x: (x[0], x[1]))
            assert all(s % world == r for s, _ in fps)
            seen += fps
        assert sorted(seen) == [(s, f) for s, n in enumerate(lengths) for f in range(n)]
        assert sharding.padded_count(lengths, world) == max(
            sum(lengths[s] for s in range(r, len(lengths), world)) for r in range(world))


def _fake_result(scene, frame, width):
    return torch.arange(width, dtype=torch.float32) + 1000.0 * scene + 10.0 * frame


def _worker(rank, world, port, lengths, width, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.frame_pairs_for_rank(lengths, world, rank)
        pad = sharding.padded_count(lengths, world)
        block = torch.zeros((pad, width))
        for chunk_start, chunk in zip(range(0, len(mine), 2), sharding.batches(mine, 2)):   # batches of 2 frame pairs
            for j, (s, f) in enumerate(chunk):
                block[chunk_start + j] = _fake_result(s, f, width)
        gathered = sharding.gather_rank_blocks(block)
        per_scene = sharding.scatter_to_scene_order(gathered, lengths, world)
        ok = all(torch.equal(per_scene[s][f], _fake_result(s, f, width))
                 for s, n in enumerate(lengths) for f in range(n))
        ret[rank] = bool(ok) and tuple(gathered.shape) == (world, pad, width)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_restores_scene_order():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    lengths = [3, 2, 4, 1, 2]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, 6, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] is True and ret[1] is True


def test_single_process_gather_is_identity():
    x = torch.randn(3, 4)
    assert torch.equal(sharding.gather_rank_blocks(x)[0], x)
