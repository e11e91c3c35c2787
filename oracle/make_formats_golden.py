"""Runs the UNMODIFIED reference ``NuScenesDataset.get_sensor_data`` (nuscenes.py:198-349) on synthetic per-frame
detection / label files and writes tests/golden/formats_*.npz. Run in the authoring container (needs /root/reference)."""
import contextlib
import io
import json
import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_dataset_loader  # noqa: E402

NAMES = ["car", "pedestrian", "truck"]


def synth_frame(rng, n):
    dets, cls = [], []
    for _ in range(n):
        yaw = rng.uniform(-np.pi, np.pi)
        q = [float(np.cos(yaw / 2)), 0.0, 0.0, float(np.sin(yaw / 2))]
        if rng.random() < 0.3:   # a not-quite-normalised, slightly tilted quaternion
            q = [q[0] * 1.01, 0.02, -0.01, q[3] * 1.01]
        dets.append([float(v) for v in rng.uniform(-50, 50, 2)] + [float(rng.normal(-1, 1))] +
                    [float(v) for v in rng.uniform(0.5, 4.5, 3)] + q + [float(v) for v in rng.normal(0, 3, 2)])
        cls.append({"detection_name": NAMES[int(rng.integers(len(NAMES)))], "detection_score": float(rng.uniform(0.1, 1)),
                    "sample_token": "x"})
    return dets, cls


def synth_labels(rng, n_prev, k):
    matched = np.zeros((n_prev, k + 2))
    used = set()
    for i in range(n_prev):
        r = rng.random()
        free = [j for j in range(k) if j not in used]
        if r < 0.55 and free:
            j = free[int(rng.integers(len(free)))]
            used.add(j)
            matched[i, j] = 1
        elif r < 0.7:
            matched[i, -1] = 1
    matched[:, -2] = 1 - matched.sum(axis=1)
    newborn = np.array([1.0 if (j not in used and rng.random() < 0.4) else 0.0 for j in range(k)])
    return matched, newborn


def run_case(mod, seed, n_prev, n_cur, max_objects, det_type, test_mode, first):
    rng = np.random.default_rng(seed)
    with tempfile.TemporaryDirectory() as d:
        det_dir, cls_dir, lab_dir = (os.path.join(d, x) for x in ("det", "cls", "lab"))
        for x in (det_dir, cls_dir, lab_dir):
            os.makedirs(x)
        frames = {"prev": synth_frame(rng, n_prev), "cur": synth_frame(rng, n_cur)}
        for tok, (dets, cls) in frames.items():
            json.dump(dets, open(os.path.join(det_dir, tok + ".json"), "w"))
            json.dump(cls, open(os.path.join(cls_dir, tok + ".json"), "w"))
        matched, newborn = synth_labels(rng, n_prev, n_cur)
        np.savez_compressed(os.path.join(lab_dir, "cur.npz"), matched=matched, newborn=newborn)
        self = types.SimpleNamespace(
            _nusc_infos=[{"token": "prev"}, {"token": "cur"}],
            _frame_info={"cur": {"prev": "" if first else "prev", "timestamp": 1_500_000, "prev_timestamp": 1_000_000}},
            get_frame_idx=lambda tok: {"prev": 0, "cur": 1}.get(tok), _max_objects=max_objects, _det_path=det_dir,
            _cls_info_path=cls_dir, _det_type=det_type, _labels_path=lab_dir, test_mode=test_mode, _fp_ratio=0.5,
            _dead_trk_ratio=0.5, nsweeps=10, _root_path="", _num_point_features=5, virtual=False,
            pipeline=lambda res, info: ({"metadata": res["metadata"]}, None))
        random.seed(seed)
        mod.NuScenesDataset.get_sensor_data(self, 1)
        info = self._nusc_infos[1]
        out = {"prev_det_boxes": info["prev_det_boxes"], "det_boxes": info["det_boxes"],
               "num_prev_det_boxes": info["num_prev_det_boxes"], "num_det_boxes": info["num_det_boxes"],
               "n_prev_cls": len(info["prev_cls_det_boxes"]), "n_cls": len(info["cls_det_boxes"]),
               "prev_scores": np.array([c["detection_score"] for c in info["prev_cls_det_boxes"]]),
               "scores": np.array([c["detection_score"] for c in info["cls_det_boxes"]])}
        if not test_mode:
            out["gt"] = info["gt"]
        return out


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        mod = ref_dataset_loader.load()
    cases = [dict(seed=1, n_prev=12, n_cur=15, max_objects=20, det_type=None, test_mode=False, first=False),
             dict(seed=2, n_prev=40, n_cur=35, max_objects=20, det_type=["car", "truck"], test_mode=False, first=False),
             dict(seed=3, n_prev=9, n_cur=11, max_objects=20, det_type=["car"], test_mode=False, first=True),
             dict(seed=4, n_prev=30, n_cur=0, max_objects=16, det_type=None, test_mode=True, first=False),
             dict(seed=5, n_prev=0, n_cur=25, max_objects=16, det_type=None, test_mode=True, first=False),
             dict(seed=6, n_prev=60, n_cur=70, max_objects=50, det_type=["pedestrian", "car"], test_mode=False, first=False)]
    for c in cases:
        out = run_case(mod, **c)
        path = os.path.join(ROOT, "tests", "golden", "formats_seed%d.npz" % c["seed"])
        np.savez_compressed(path, case=json.dumps(c), **out)
        print(path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
