"""End to end at the tool level (tools/nusc_shasta/eval.py + pub_tracker_merged.py): detection JSON of a scene ->
frame-pair packing -> Shasta.forward (CUDA) -> device decode -> annotations -> greedy ID tracker (device assignment),
against the same chain on the CPU oracles (O.forward, O.decode, tracker_oracle). The discrete outcome - which
detections survive, which are newborn / dead / FN, every tracking id - must be identical."""
import copy

import numpy as np
import pytest
import torch

from oracle import shasta_oracle as O
from oracle import tracker_oracle as TO
from shasta_b200 import formats, pipeline, synthetic
from shasta_b200.tracker import PubTrackerMerged
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _scene(seed, nframes, M, extent):
    rng = np.random.default_rng(seed)
    objs = [{"p": rng.uniform(-extent, extent, 2), "v": rng.normal(0, 2, 2), "yaw": rng.uniform(-3, 3),
             "size": rng.uniform(1.5, 4.5, 3)} for _ in range(int(M * 0.6))]
    frames = []
    for f in range(nframes):
        dets, cls = [], []
        for o in objs:
            if rng.random() < 0.12:
                continue
            p = o["p"] + rng.normal(0, 0.15, 2)
            q = [float(np.cos(o["yaw"] / 2)), 0.0, 0.0, float(np.sin(o["yaw"] / 2))]
            dets.append([float(p[0]), float(p[1]), -1.0] + [float(s) for s in o["size"]] + q +
                        [float(o["v"][0]), float(o["v"][1])])
            cls.append({"detection_name": "car", "detection_score": float(rng.uniform(0.2, 1)),
                        "translation": [float(p[0]), float(p[1]), -1.0], "velocity": [float(o["v"][0]), float(o["v"][1])],
                        "sample_token": "t%d" % f})
        for o in objs:
            o["p"] = o["p"] + o["v"] * 0.5
        if rng.random() < 0.5:
            objs.append({"p": rng.uniform(-extent, extent, 2), "v": rng.normal(0, 2, 2), "yaw": rng.uniform(-3, 3),
                         "size": rng.uniform(1.5, 4.5, 3)})
        frames.append({"token": "t%d" % f, "prev_token": "" if f == 0 else "t%d" % (f - 1),
                       "timestamp": 500_000 * (f + 1), "prev_timestamp": 500_000 * f, "dets": dets, "cls": cls})
    return frames


def _cpu_chain(weights, frames, maps, M, pc_start):
    wt = O.weights_to_torch(weights)
    by_token = {f["token"]: f for f in frames}
    results, dead_tracker = {}, {}
    for f in frames:
        token, prev_token = f["token"], f["prev_token"]
        dead_tracker.setdefault(token, {"dead_idx": [], "keep_idx": []})
        prev = by_token.get(prev_token) if prev_token else None
        ex = formats.frame_pair_example(None if prev is None else copy.deepcopy(prev["dets"]),
                                        None if prev is None else copy.deepcopy(prev["cls"]), copy.deepcopy(f["dets"]),
                                        copy.deepcopy(f["cls"]), M, 1e-6 * f["timestamp"] - 1e-6 * f["prev_timestamp"], ["car"])
        bev, prev_bev = maps[token]
        m1, m2 = O.forward(wt, bev, prev_bev, torch.from_numpy(ex["det_boxes"].copy()),
                           torch.from_numpy(ex["prev_det_boxes"]), pc_start=pc_start)
        n_prev, n_det = len(ex["prev_cls_det_boxes"]), len(ex["cls_det_boxes"])
        d = O.decode(m1[0], m2[0], n_prev, n_det)
        prev_state = [1 if n in d["dead"] else 2 if n in d["fn"] else 0 for n in range(n_prev)]
        fn_score = np.zeros(max(n_prev, 1), np.float32)
        for n in d["fn"]:
            fn_score[n] = float(m1[0][n, -2])     # raw value: annos_from_decode forms 1 - value in double
        det_state = [2] * n_det
        det_score = np.zeros(max(n_det, 1), np.float32)
        for k, nb in zip(d["keep_dets"], d["newborn"]):
            det_state[k] = 1 if nb else 0
            det_score[k] = float(m2[0][-1, k])
        annos, dead_idx, keep = formats.annos_from_decode(ex["prev_cls_det_boxes"], ex["cls_det_boxes"], prev_state,
                                                          fn_score, det_state, det_score, token,
                                                          float(ex["prev_det_boxes"][0, 0, 9]))
        if n_prev > 0:
            dead_tracker.setdefault(prev_token, {"dead_idx": [], "keep_idx": []})["dead_idx"].extend(dead_idx)
        if n_det > 0:
            dead_tracker[token]["keep_idx"] = keep
        results[token] = annos
    return formats.mark_dead(results, dead_tracker)


@pytest.mark.parametrize("seed,peaky", [(3, 400.0), (4, 900.0)])
def test_scene_through_head_decode_and_tracker(seed, peaky):
    M, H, W = 20, 32, 32
    pc_start = (-W * 0.3, -H * 0.3)
    weights = synthetic.make_weights(M, seed=seed, peaky=peaky)
    model = G.make_model(M, pc_start, weights)
    frames = _scene(seed, 9, M, extent=8.0)
    g = torch.Generator().manual_seed(seed)
    maps = {}
    for f in frames:   # the previous map of a frame is the current map of the frame before
        cur = torch.relu(torch.randn((1, H, W, 64), generator=g))
        prev = maps[f["prev_token"]][0] if f["prev_token"] else torch.relu(torch.randn((1, H, W, 64), generator=g))
        maps[f["token"]] = (cur, prev)
    got = pipeline.run_class_sequence(model, copy.deepcopy(frames), lambda tok: tuple(m.to(G.DEV) for m in maps[tok]),
                                      det_type=["car"])
    want = _cpu_chain(weights, copy.deepcopy(frames), maps, M, pc_start)
    flags = 0
    for f in frames:
        a, b = got[f["token"]], want[f["token"]]
        assert len(a) == len(b), f["token"]
        for x, y in zip(a, b):
            assert x["translation"] == y["translation"]
            assert x.get("newborn") == y.get("newborn") and x.get("dead") == y.get("dead") and x.get("FN") == y.get("FN")
            assert abs(x["ref_detection_score"] - y["ref_detection_score"]) < 1e-4
            flags += bool(x.get("newborn")) + bool(x.get("dead")) + bool(x.get("FN"))
    assert flags > 0, "weights not peaky enough: the scene exercised no newborn / dead / FN decision"
    # downstream ID tracker: device assignment vs numpy oracle, on the two annotation streams
    ta, tb = PubTrackerMerged(max_age=3), TO.Tracker(max_age=3)
    for f in frames:
        ra = ta.step_centertrack(copy.deepcopy(got[f["token"]]), 0.5)
        rb = tb.step(copy.deepcopy(want[f["token"]]), 0.5)
        sa, sb = TO.summarize(ra), TO.summarize(rb)
        assert [(t["id"], t["age"], t["active"], t["x"]) for t in sa] == [(t["id"], t["age"], t["active"], t["x"]) for t in sb]
        assert all(abs(p["score"] - q["score"]) < 1e-4 for p, q in zip(sa, sb))
