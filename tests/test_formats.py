"""Wire formats either side of the path: shasta_b200.formats against fixtures written by the
UNMODIFIED reference NuScenesDataset.get_sensor_data (oracle/make_formats_golden.py, seeded ``random``)."""
import glob
import json
import os
import random

import numpy as np
import pytest

from oracle import make_formats_golden as MG
from shasta_b200 import formats as F

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "formats_seed*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_pack_and_gt_match_reference_fixture(path):
    g = np.load(path, allow_pickle=False)
    c = json.loads(str(g["case"]))
    rng = np.random.default_rng(c["seed"])
    prev_dets, prev_cls = MG.synth_frame(rng, c["n_prev"])
    cur_dets, cur_cls = MG.synth_frame(rng, c["n_cur"])
    matched, newborn = MG.synth_labels(rng, c["n_prev"], c["n_cur"])
    random.seed(c["seed"])
    dt = 1e-6 * 1_500_000 - 1e-6 * 1_000_000
    M = c["max_objects"]
    if c["first"]:
        pb, pk, pc, pn = np.zeros((M, 11)), list(range(M)), [], 0
    else:
        pb, pk, pc, pn = F.pack_detections(prev_dets, prev_cls, M, dt, c["det_type"], random)
    cb, ck, cc, cn = F.pack_detections(cur_dets, cur_cls, M, dt, c["det_type"], random)
    assert np.array_equal(pb, g["prev_det_boxes"]) and np.array_equal(cb, g["det_boxes"])
    assert len(pc) == int(g["n_prev_cls"]) and len(cc) == int(g["n_cls"])
    assert np.array_equal(np.array([x["detection_score"] for x in pc]), g["prev_scores"])
    assert np.array_equal(np.array([x["detection_score"] for x in cc]), g["scores"])
    if c["test_mode"]:
        assert pn == int(g["num_prev_det_boxes"]) and cn == int(g["num_det_boxes"])
        return
    gt, num_prev, num_det = F.build_gt_affinity(matched, newborn, pk, ck, M, not c["first"], 0.5, 0.5, random)
    assert np.array_equal(gt, g["gt"])
    assert num_det == int(g["num_det_boxes"])
    assert (pn if num_prev is None else num_prev) == int(g["num_prev_det_boxes"])


def test_fixtures_exist():
    assert len(GOLDEN) >= 6


def test_quaternion_yaw_known_values():
    for yaw in (-3.0, -1.2, 0.0, 0.7, 3.1):
        q = [np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2)]
        assert abs(F.quaternion_yaw(q)[0] - yaw) < 1e-12
        assert abs(F.quaternion_yaw([2 * v for v in q])[0] - yaw) < 1e-12   # normalisation


def test_label_affinity_rows_sum_to_one_and_first_frame():
    # previous frame: dets 0,1,2 are TPs of GT 0,1,2 (ids a,b,c), det 3 is a FP; current: det 0 <-> id b, det 1 <-> id d
    prev = ({0: 0, 1: 1, 2: 2}, ["a", "b", "c"], 4)
    matched, newborn = F.label_affinity({0: 0, 1: 1}, ["b", "d", "c"], fn_inds=[2], num_dets=3, prev=prev)
    assert matched.shape == (4, 5) and np.allclose(matched.sum(axis=1), 1)
    assert matched[1, 0] == 1            # id b continues
    assert matched[2, -1] == 1           # id c is a GT of the current frame without a detection: FN track
    assert matched[0, -2] == 1 and matched[3, -2] == 1   # id a left, det 3 was a FP: dead
    assert newborn.tolist() == [0.0, 1.0, 0.0]
    m0, n0 = F.label_affinity({0: 0, 2: 1}, ["x", "y"], fn_inds=[], num_dets=3, prev=None)
    assert m0 is None and n0.tolist() == [1.0, 0.0, 1.0]


def test_frame_pair_example_feeds_the_head_layout():
    rng = np.random.default_rng(0)
    pd, pc = MG.synth_frame(rng, 7)
    cd, cc = MG.synth_frame(rng, 5)
    ex = F.frame_pair_example(pd, pc, cd, cc, max_objects=10, time_diff=0.5)
    assert ex["det_boxes"].shape == (1, 10, 11) and ex["det_boxes"].dtype == np.float32
    assert ex["num_prev_det_boxes"] == 7 and ex["num_det_boxes"] == 5
    assert np.all(ex["det_boxes"][0, 5:] == 0) and np.all(ex["prev_det_boxes"][0, :7, 9] == 0.5)
    first = F.frame_pair_example(None, None, cd, cc, max_objects=10, time_diff=0.5)
    assert first["num_prev_det_boxes"] == 0 and not first["prev_det_boxes"].any()


def test_annos_from_decode_follows_the_eval_loop():
    """Against the oracle's decode (pinned restatement of eval.py:126-181) on planted matrices."""
    import copy
    import torch
    from oracle import shasta_oracle as O
    M, n_prev, n_det = 6, 4, 5
    m1 = torch.full((M, M + 2), 0.01)
    m2 = torch.full((M + 2, M), 0.01)
    m1[0, M] = 0.9           # prev 0 dead
    m1[1, M + 1] = 0.8       # prev 1 FN
    m1[2, 3] = 0.95          # prev 2 -> det 3
    m2[M + 1, 0] = 0.75      # det 0 dropped as FP
    m2[M, 1] = 0.6           # det 1 newborn
    m2[2, 3] = 0.9
    want = O.decode(m1, m2, n_prev, n_det)
    prev_state = [1 if n in want["dead"] else 2 if n in want["fn"] else 0 for n in range(n_prev)]
    fn_score = [float(m1[n, -2]) for n in range(n_prev)]       # raw matched1[n,-2]: the helper forms 1 - value
    det_state = [0 if k in want["keep_dets"] else 2 for k in range(n_det)]
    for k, nb in zip(want["keep_dets"], want["newborn"]):
        if nb:
            det_state[k] = 1
    det_score = [float(m2[-1, k]) for k in range(n_det)]      # raw matched2[-1,k]
    mk = lambda i: {"translation": [float(i), 2.0 * i, 0.0], "velocity": [1.0, -1.0], "detection_name": "car",  # noqa: E731
                    "detection_score": 0.5}
    prev_cls, cur_cls = [mk(i) for i in range(n_prev)], [mk(10 + i) for i in range(n_det)]
    annos, dead_idx, keep = F.annos_from_decode(copy.deepcopy(prev_cls), cur_cls, prev_state, fn_score, det_state,
                                                det_score, "tok", 0.5)
    assert dead_idx == want["dead"] == [0] and keep == want["keep_dets"]
    assert [a.get("FN", False) for a in annos] == [False] * len(keep) + [True]
    assert annos[-1]["translation"][:2] == [1.0 + 0.5, 2.0 - 0.5] and annos[-1]["token"] == "tok"
    assert cur_cls[1].get("newborn") is True and "newborn" not in cur_cls[2]
    res = F.mark_dead({"t0": [{"a": 1}, {"a": 2}]}, {"t0": {"dead_idx": [3], "keep_idx": [1, 3]}})
    assert res["t0"][1].get("dead") is True and "dead" not in res["t0"][0]


@pytest.mark.parametrize("path", sorted(__import__("glob").glob(os.path.join(os.path.dirname(__file__), "golden", "labelaff_seed*.json"))),
                         ids=lambda p: os.path.basename(p))
def test_label_affinity_matches_reference_script(path):
    """Fixtures written by the UNMODIFIED preprocessing/make_gt_shasta.py (oracle/make_labelaff_golden.py runs its
    main() on a synthetic scene with the devkit stubbed): matched / newborn per frame must be identical."""
    import json
    g = json.load(open(path))
    scene = g["scene"]

    def assoc(f):
        tp = {d: gi for d, gi in enumerate(f["det_gt"]) if gi >= 0}
        fn = [gi for gi in range(len(f["gt_ids"])) if gi not in tp.values()]
        return tp, fn

    for i, (f, want) in enumerate(zip(scene, g["outputs"])):
        tp, fn = assoc(f)
        prev = None
        if i > 0:
            ptp, _ = assoc(scene[i - 1])
            prev = (ptp, scene[i - 1]["gt_ids"], len(scene[i - 1]["det_gt"]))
        matched, newborn = F.label_affinity(tp, f["gt_ids"], fn, len(f["det_gt"]), prev)
        assert newborn.tolist() == want["newborn"], i
        if want["matched"] is None:
            assert matched is None
        else:
            assert matched.tolist() == want["matched"], i
            assert np.all(matched.sum(axis=1) == 1)
