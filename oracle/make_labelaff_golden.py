"""TEST INFRASTRUCTURE — fixtures for ``shasta_b200.formats.label_affinity`` written by the UNMODIFIED reference script
``/root/reference/preprocessing/make_gt_shasta.py`` (its ``main``, lines 42-168).

The script is not importable as is: it parses ``sys.argv`` at import time and needs the nuScenes devkit, pyquaternion,
``detection_nms`` and ``gt_association`` (its own preprocessing helpers, which in turn need the devkit). They are
replaced by stub modules whose only job is to hand the script a synthetic scene: per frame a list of detections, a
list of ground-truth boxes with instance ids, and the detection<->GT association (``associate`` stub). What is pinned
is the script's own logic - which previous detection is matched / dead / FN and which detection is newborn.

    python -m oracle.make_labelaff_golden        (authoring container only; writes tests/golden/labelaff_seed*.json)
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("SHASTA_REF_ROOT", "/root/reference")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synthetic_scene(seed, frames=9, max_obj=9):
    """Per frame: ``gt_ids`` (instance ids of the GT boxes), ``det_gt`` (per detection: index of the GT it is a true
    positive of, or -1 = false positive), i.e. tp_ind_pairs = {det: gt}; GTs nobody detects are the FNs."""
    rng = np.random.default_rng(seed)
    alive = list(range(4))
    next_id = 4
    out = []
    for _ in range(frames):
        alive = [i for i in alive if rng.random() > 0.2]
        while len(alive) < max_obj and rng.random() < 0.6:
            alive.append(next_id)
            next_id += 1
        gt_ids = [int(i) for i in rng.permutation(alive)]
        det_gt = [g for g in range(len(gt_ids)) if rng.random() < 0.75]      # detected GTs
        det_gt += [-1] * int(rng.integers(0, 3))                             # false positives
        det_gt = [int(x) for x in rng.permutation(det_gt)]
        out.append({"gt_ids": gt_ids, "det_gt": det_gt})
    return out


def run_reference(scene):
    """Executes the reference script's main() on the synthetic scene; returns per frame (matched, newborn)."""
    tokens = ["tok%02d" % i for i in range(len(scene))]
    frames_by_id = {}

    def associate(frame_gt, frame_gt_types, frame_dets, frame_types, threshold=None):
        f = scene[frame_dets.frame]
        tp = {d: g for d, g in enumerate(f["det_gt"]) if g >= 0}
        fp = [d for d, g in enumerate(f["det_gt"]) if g < 0]
        fn = [g for g in range(len(f["gt_ids"])) if g not in tp.values()]
        return (None,) * 7 + (tp, fp, fn)

    class Dets(list):     # a list that remembers which frame it belongs to (the associate stub needs it)
        frame = -1

    def load_dets(path, data_folder, scene_name):
        dets = []
        for i, f in enumerate(scene):
            d = Dets(range(len(f["det_gt"])))
            d.frame = i
            dets.append(d)
        return dets, [["car"] * len(f["det_gt"]) for f in scene]

    class Nusc:
        scene = [{"name": "scene-0001", "first_sample_token": tokens[0]}]

        def get(self, table, token):
            i = tokens.index(token)
            return {"prev": tokens[i - 1] if i > 0 else "", "next": tokens[i + 1] if i + 1 < len(tokens) else ""}

    stubs = {}
    for name in ("nuscenes", "nuscenes.utils", "nuscenes.utils.data_classes", "nuscenes.nuscenes", "nuscenes.utils.splits",
                 "pyquaternion", "detection_nms", "gt_association", "gt_association.associate", "tqdm"):
        stubs[name] = types.ModuleType(name)
    stubs["nuscenes.utils.data_classes"].Box = object
    stubs["pyquaternion"].Quaternion = object
    stubs["detection_nms"].load_dets = load_dets
    stubs["detection_nms"].nu_array2mot_bbox = lambda x: x
    stubs["gt_association.associate"].associate = associate
    stubs["nuscenes.nuscenes"].NuScenes = object
    stubs["nuscenes.utils"].splits = stubs["nuscenes.utils.splits"]
    stubs["nuscenes"].utils = stubs["nuscenes.utils"]

    class _Bar:
        def __init__(self, *a, **k):
            pass

        def update(self, n):
            pass

        def close(self):
            pass
    stubs["tqdm"].tqdm = _Bar
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    argv, cwd = sys.argv, os.getcwd()
    try:
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, "data", "nusc_preprocssed"))     # (sic: the path the script opens)
            with open(os.path.join(tmp, "data", "nusc_preprocssed", "train_frame_info.json"), "w") as fh:
                json.dump({t: {} for t in tokens}, fh)
            os.chdir(tmp)
            sys.argv = ["make_gt_shasta.py"]
            src = open(os.path.join(REF, "preprocessing", "make_gt_shasta.py")).read()
            ns = {"__name__": "make_gt_shasta_under_test"}
            exec(compile(src, "make_gt_shasta.py", "exec"), ns)            # the unmodified file; __main__ block skipped
            ns["load_gt_bboxes"] = lambda gt_folder, data_folder, seg: (
                [list(range(len(f["gt_ids"]))) for f in scene], [list(f["gt_ids"]) for f in scene],
                [["car"] * len(f["gt_ids"]) for f in scene])
            ns["main"](Nusc(), ["scene-0001"], "cp", 2.0, os.path.join(tmp, "out"), "gt", "det", "gt_shasta")
            res = []
            for t in tokens:
                z = np.load(os.path.join(tmp, "out", "gt_shasta", "cp", "individual_frames", t + ".npz"), allow_pickle=True)
                m = z["matched"]
                res.append({"matched": None if m.dtype == object or m.shape == () else m.tolist(),
                            "newborn": z["newborn"].tolist()})
            return res
    finally:
        os.chdir(cwd)
        sys.argv = argv
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def main():
    for seed in (1, 2, 3):
        scene = synthetic_scene(seed)
        out = run_reference(scene)
        path = os.path.join(GOLDEN_DIR, "labelaff_seed%d.json" % seed)
        with open(path, "w") as fh:
            json.dump({"seed": seed, "scene": scene, "outputs": out}, fh)
        print(path, "frames", len(scene), "matched shapes",
              [None if o["matched"] is None else np.array(o["matched"]).shape for o in out])


if __name__ == "__main__":
    main()
