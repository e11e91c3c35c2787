"""BASELINE.json configs[2] / SURVEY.md §8d C3 as a throughput run: the 7 shipped class models + 3 synthetic ones
(max_obj up to 500), `--scenes` x `--pairs` frame pairs per class, scenes dealt round-robin to the ranks, decode blocks
gathered once per class. One JSON line from rank 0; launch like bench.py (python, or torch.distributed.run for N > 1).
Not the headline bench (bench.py is): a measured data point for the sharded multi-class configuration."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from shasta_b200 import build_track, multiclass  # noqa: E402


def add_args(ap):
    ap.add_argument("--scenes", type=int, default=150)
    ap.add_argument("--pairs", type=int, default=40)
    ap.add_argument("--ring", type=int, default=2)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--no-step-graphs", action="store_true")
    ap.add_argument("--classes", type=str, default="")
    ap.add_argument("--big-batch", type=int, default=128,
                    help="frame pairs per step of the class models with max_obj >= 200: their aug_shape weight stream "
                         "(1.03 / 6.4 GB per step) is amortised over twice as many frame pairs as with 64")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--hw", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64)
    add_args(ap)
    run(ap.parse_args())


def run(a):

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    classes = list(multiclass.NUSC_CLASS_MAX_OBJ + multiclass.SYNTHETIC_CLASS_MAX_OBJ)
    if a.classes:
        classes = [c for c in classes if c[0] in a.classes.split(",")]
    lengths = [a.pairs] * a.scenes
    pc_start = (-a.hw * 0.3, -a.hw * 0.3)
    lanes = []
    for i, (name, M) in enumerate(classes):
        cfg = dict(type="Shasta", reader=None, backbone=None, neck=None,
                   bev_extractor=dict(type="BEVFeatureExtractor", pc_start=list(pc_start), voxel_size=[0.075, 0.075],
                                      out_stride=8), max_obj=M, num_feats=3)
        torch.manual_seed(i)
        with torch.device(device):
            model = build_track(cfg)
        model.eval()
        lanes.append(multiclass.ClassLane(name, model, step_graphs=not a.no_step_graphs))
    batch_pairs = {name: (a.big_batch if M >= 200 else a.batch) for name, M in classes}
    prov = multiclass.SyntheticProvider(classes, lengths, a.hw, device, seed=1000 * 3 + rank, ring=a.ring,
                                        batch_pairs=batch_pairs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(marks=None):
        def done(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        return multiclass.run_sequence_batch(lanes, lengths, prov, batch_pairs, world_size=world, rank=rank,
                                             on_class_done=done)

    t_setup = time.time()
    one_pass()                      # warm-up pass: packs the weights, captures the step graphs
    barrier()
    t_setup = time.time() - t_setup
    sampler = B.ClockSampler(local)
    sampler.start()
    times, per_class = [], None
    t0w = time.time()
    for _ in range(a.passes):
        marks = []
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        res = one_pass(marks)
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1))
        prev, pc = e0, {}
        for name, e in marks:
            pc[name] = round(prev.elapsed_time(e), 3)
            prev = e
        per_class = pc
    t1w = time.time()
    clocks = sampler.stop(t0w, t1w)
    ms = torch.tensor([min(times)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    total_pairs = len(classes) * a.scenes * a.pairs
    if rank == 0:
        assert all(len(res[name]) == a.scenes for name, _ in classes)
        states = multiclass.decode_fields(res[classes[0][0]][0])["prev_state"]
        print(json.dumps({
            "metric": "frame-pairs/sec, %d-class sequence batch (per-class models, scenes sharded over ranks)" % len(classes),
            "value": total_pairs / (ms * 1e-3), "unit": "frame-pairs/s", "n_gpus": world, "ms_per_pass": ms,
            "passes": a.passes, "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[2]", "classes": dict(classes), "scenes": a.scenes,
                       "frame_pairs_per_scene": a.pairs, "frame_pairs_total": total_pairs, "bev_hw": a.hw,
                       "batch_pairs": batch_pairs, "input_ring": a.ring, "step_graphs": not a.no_step_graphs,
                       "timed_region": "all class lanes over the rank's frame pairs (box refresh, forward, decode) + "
                                       "the gather of the decode blocks; max over ranks, best of the passes"},
            "per_class_ms_rank0": per_class, "setup_s": round(t_setup, 1), "clocks": clocks,
            "sanity": {"prev_state_values": sorted(set(states.flatten().tolist()))},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
