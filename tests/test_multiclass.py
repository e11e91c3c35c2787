"""Host logic of the per-class sequence batch (shasta_b200/multiclass.py; BASELINE.json configs[2]) on CPU: scene
sharding over a world of 2 (gloo), batching with a ragged last batch, gather back into scene order. The class lanes
are stubs here (the CUDA head is exercised by tests/test_gpu_multiclass.py)."""
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shasta_b200 import multiclass


class _StubLane:
    """Decode block of a frame pair = a function of (class, scene, frame) only."""

    def __init__(self, name, M, salt):
        self.name, self.salt = name, salt
        self.model = types.SimpleNamespace(max_obj=M, aff=[types.SimpleNamespace(weight=torch.zeros(1))])
        self.batch_sizes = []

    def step(self, batch):
        ids = batch["ids"]                                  # (B, 2) scene, frame
        B, M = ids.shape[0], self.model.max_obj
        self.batch_sizes.append(B)
        base = (ids[:, 0] * 1000 + ids[:, 1] * 10 + self.salt).to(torch.int32)
        return (base.view(1, B, 1) + torch.arange(6, dtype=torch.int32).view(6, 1, 1) * 100000
                + torch.arange(M, dtype=torch.int32).view(1, 1, M) * 0).contiguous()


def _provider(name, items):
    return {"ids": torch.tensor(items, dtype=torch.int64).view(-1, 2)}


def _expected(scene_lengths, M, salt):
    out = []
    for s, n in enumerate(scene_lengths):
        t = torch.empty((n, 6, M), dtype=torch.int32)
        for f in range(n):
            for k in range(6):
                t[f, k] = s * 1000 + f * 10 + salt + k * 100000
        out.append(t)
    return out


def _run(world, rank, lengths):
    lanes = [_StubLane("car", 9, 1), _StubLane("bus", 4, 2)]
    res = multiclass.run_sequence_batch(lanes, lengths, _provider, {"car": 4, "bus": 3}, world_size=world, rank=rank)
    return lanes, res


def _worker(rank, world, port, lengths, ret):
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, world_size=world, rank=rank)
    try:
        _, res = _run(world, rank, lengths)
        ret[rank] = {k: [t.clone() for t in v] for k, v in res.items()}
    finally:
        dist.destroy_process_group()


def test_single_rank_batches_and_scene_order():
    lengths = [3, 5, 1, 4]
    lanes, res = _run(1, 0, lengths)
    assert lanes[0].batch_sizes == [4, 4, 4, 1] and lanes[1].batch_sizes == [3, 3, 3, 3, 1]
    for lane in lanes:
        want = _expected(lengths, lane.model.max_obj, lane.salt)
        assert all(torch.equal(a, b) for a, b in zip(res[lane.name], want))


def test_gloo_world2_matches_single_rank():
    lengths = [3, 5, 1, 4, 2]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29611, lengths, ret), nprocs=2, join=True)
    for rank in range(2):
        for name, M, salt in (("car", 9, 1), ("bus", 4, 2)):
            want = _expected(lengths, M, salt)
            got = ret[rank][name]
            assert len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want))


def test_decode_fields_bitcast():
    block = torch.zeros((2, 6, 3), dtype=torch.int32)
    block[:, 2] = torch.tensor([0.25, 0.5, 1.0]).view(torch.int32)
    f = multiclass.decode_fields(block)
    assert f["fn_dead_prob"].dtype == torch.float32 and torch.equal(f["fn_dead_prob"][1], torch.tensor([0.25, 0.5, 1.0]))
    assert f["prev_state"].dtype == torch.int32 and set(f) == set(multiclass.DECODE_FIELDS)


def test_class_table_matches_the_shipped_configs():
    assert dict(multiclass.NUSC_CLASS_MAX_OBJ) == {"car": 90, "pedestrian": 90, "bus": 20, "truck": 60, "trailer": 60,
                                                   "bicycle": 50, "motorcycle": 50}
    assert len(multiclass.NUSC_CLASS_MAX_OBJ) + len(multiclass.SYNTHETIC_CLASS_MAX_OBJ) == 10
    assert max(m for _, m in multiclass.SYNTHETIC_CLASS_MAX_OBJ) == 500


def test_rank_without_scenes_contributes_an_empty_block():
    """More ranks than scenes: the idle rank runs no batch and hands back a zero-row block (gather=False view)."""
    lanes = [_StubLane("car", 5, 1)]
    res = multiclass.run_sequence_batch(lanes, [2, 3], _provider, 4, world_size=3, rank=2, gather=False)
    assert lanes[0].batch_sizes == [] and tuple(res["car"].shape) == (6, 0, 5)
    res = multiclass.run_sequence_batch(lanes, [2, 3], _provider, 4, world_size=3, rank=1, gather=False)
    assert lanes[0].batch_sizes == [3] and tuple(res["car"].shape) == (6, 3, 5)
