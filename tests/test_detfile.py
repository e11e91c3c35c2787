"""The binary detection batch file (shasta_b200/detfile.py) against the JSON path it replaces:
``formats.frame_pair_example`` on the per-frame detection lists (itself pinned to the reference's dataset code by
tests/test_formats.py). Packed arrays, counts and keep indices must be identical bit for bit, incl. the class filter,
empty frames, scene starts and the seeded sub-sampling of over-full frames."""
import copy
import random

import numpy as np
import pytest

from shasta_b200 import detfile, formats

CLASSES = ["car", "truck", "pedestrian"]


def _frames(seed, n_scenes=3, n_frames=5, max_dets=14):
    rng = np.random.default_rng(seed)
    frames = []
    for s in range(n_scenes):
        for f in range(n_frames):
            n = 0 if (s == 1 and f == 2) else int(rng.integers(1, max_dets))
            dets, cls = [], []
            # the det_path boxes are in the SENSOR frame of their sweep, the cls_info entries in the GLOBAL frame
            # (preprocessing/filter_track_types.py): a per-frame rigid transform (ego pose) separates the two
            ego_yaw = float(rng.uniform(-np.pi, np.pi))
            ego_t = rng.uniform(-500, 500, 3)
            c, sn = np.cos(ego_yaw), np.sin(ego_yaw)
            R = np.array([[c, -sn, 0.0], [sn, c, 0.0], [0.0, 0.0, 1.0]])
            for _ in range(n):
                q = rng.normal(size=4)
                if rng.random() < 0.5:
                    q /= np.linalg.norm(q)          # both normalised and raw quaternions
                box = (list(rng.uniform(-50, 50, 3)) + list(rng.uniform(0.5, 5, 3)) + list(q) + list(rng.normal(0, 3, 2)))
                dets.append([float(x) for x in box])
                gq = rng.normal(size=4)
                info = {"sample_token": "s%d_f%d" % (s, f),
                        "translation": [float(x) for x in R @ np.array(box[0:3]) + ego_t], "size": dets[-1][3:6],
                        "rotation": [float(x) for x in gq / np.linalg.norm(gq)],
                        "velocity": [float(x) for x in (R @ np.array(box[10:12] + [0.0]))[:2]],
                        "detection_name": CLASSES[int(rng.integers(0, 3))], "detection_score": float(rng.uniform(0.05, 1))}
                if rng.random() < 0.6:
                    info["attribute_name"] = ["vehicle.moving", "vehicle.parked", "pedestrian.standing"][int(rng.integers(0, 3))]
                if rng.random() < 0.3:                # keys beyond the nuScenes detection schema survive the round trip
                    info["num_pts"] = int(rng.integers(0, 500))
                cls.append(info)
            ts = 1_600_000_000_000_000 + (s * 100 + f) * 500_000 + int(rng.integers(0, 2000))
            frames.append({"token": "s%d_f%d" % (s, f), "prev_token": "" if f == 0 else "s%d_f%d" % (s, f - 1),
                           "timestamp": ts, "prev_timestamp": ts if f == 0 else frames[-1]["timestamp"],
                           "dets": dets, "cls": cls})
    return frames


@pytest.mark.parametrize("det_type,M", [(None, 20), (["car"], 20), (["car", "pedestrian"], 4), (["bus"], 6), (None, 3)])
def test_batch_equals_json_path(tmp_path, det_type, M):
    frames = _frames(11)
    path = str(tmp_path / "dets.shdb")
    detfile.write_detection_file(path, frames)
    df = detfile.DetectionFile(path)
    assert df.n_frames == len(frames) and df.tokens == [f["token"] for f in frames]
    by_token = {f["token"]: f for f in frames}
    order = list(range(len(frames)))
    batch = df.frame_pair_batch(order, M, det_type, rng=random.Random(5))
    ref_rng = random.Random(5)
    sampled = 0
    for j, i in enumerate(order):
        f = frames[i]
        prev = by_token.get(f["prev_token"]) if f["prev_token"] else None
        dt = 1e-6 * f["timestamp"] - 1e-6 * f["prev_timestamp"]
        ex = formats.frame_pair_example(None if prev is None else copy.deepcopy(prev["dets"]),
                                        None if prev is None else copy.deepcopy(prev["cls"]), copy.deepcopy(f["dets"]),
                                        copy.deepcopy(f["cls"]), M, dt, det_type, rng=ref_rng)
        assert np.array_equal(batch["det_boxes"][j].view(np.uint32), ex["det_boxes"][0].view(np.uint32)), f["token"]
        assert np.array_equal(batch["prev_det_boxes"][j].view(np.uint32), ex["prev_det_boxes"][0].view(np.uint32))
        assert int(batch["n_det"][j]) == ex["num_det_boxes"] and int(batch["n_prev"][j]) == ex["num_prev_det_boxes"]
        assert batch["keep"][j] == ex["keep"] and batch["prev_keep"][j] == ex["prev_keep"]
        got = df.cls_info(batch["rows"][j], f["token"])
        assert got == ex["cls_det_boxes"]
        sampled += len(f["dets"]) > M
    if M <= 4:
        assert sampled > 0, "the case was meant to exercise the seeded sub-sampling"


def test_round_trip_and_errors(tmp_path):
    frames = _frames(3, n_scenes=2, n_frames=3)
    path = str(tmp_path / "a.shdb")
    detfile.write_detection_file(path, frames)
    df = detfile.DetectionFile(path)
    assert df.frames() == frames
    assert df.frame_index("s1_f2") == 5 and int(df.prev_index[3]) == -1 and int(df.prev_index[4]) == 3
    bad = tmp_path / "bad.shdb"
    bad.write_bytes(b"NOTSHDB0" + b"\0" * 64)
    with pytest.raises(ValueError):
        detfile.DetectionFile(str(bad))
    data = open(path, "rb").read()
    cut = tmp_path / "cut.shdb"
    cut.write_bytes(data[:len(data) // 2])
    with pytest.raises(ValueError):
        detfile.DetectionFile(str(cut))


def test_empty_file(tmp_path):
    path = str(tmp_path / "e.shdb")
    detfile.write_detection_file(path, [])
    df = detfile.DetectionFile(path)
    assert df.n_frames == 0 and df.rows.shape == (0, 14)
    b = df.frame_pair_batch([], 5)
    assert b["det_boxes"].shape == (0, 5, 11)


def test_provider_packing_is_independent_of_batching(tmp_path):
    """Over-full frames are sub-sampled with a per-frame-pair generator: the choice must not depend on which other
    frame pairs share the batch (ranks and batch sizes differ between runs of the same job)."""
    from shasta_b200 import multiclass
    frames = _frames(17, n_scenes=2, n_frames=4)
    path = str(tmp_path / "d.shdb")
    detfile.write_detection_file(path, frames)
    df = detfile.DetectionFile(path)
    prov = multiclass.DetectionFileProvider(df, {"all": None}, {"all": 3}, maps_for=None, device="cpu", seed=4)
    assert prov.scene_lengths == [4, 4] and prov.frame_of(1, 2) == 6
    whole = prov.pack("all", list(range(8)))
    assert (whole["n_det"] == 3).any()
    for i in range(8):
        one = prov.pack("all", [i])
        assert np.array_equal(one["det_boxes"][0], whole["det_boxes"][i])
        assert np.array_equal(one["prev_det_boxes"][0], whole["prev_det_boxes"][i])
        assert one["keep"][0] == whole["keep"][i] and one["prev_keep"][0] == whole["prev_keep"][i]
    other = multiclass.DetectionFileProvider(df, {"all": None}, {"all": 3}, maps_for=None, device="cpu", seed=5)
    assert any(other.pack("all", [i])["keep"][0] != whole["keep"][i] for i in range(8))


def test_scenes_rejects_out_of_order_frames(tmp_path):
    frames = _frames(2, n_scenes=1, n_frames=3)
    frames[1], frames[2] = frames[2], frames[1]
    path = str(tmp_path / "o.shdb")
    detfile.write_detection_file(path, frames)
    with pytest.raises(ValueError):
        detfile.DetectionFile(path).scenes()


def test_json_dirs_to_file_and_results_json(tmp_path):
    """The reference's on-disk layout (per-token detection + cls_info JSON, frame_info dict) -> one detection file;
    and the cp_<split>.json writer."""
    import json
    frames = _frames(23, n_scenes=2, n_frames=3)
    det_dir, cls_dir = tmp_path / "det", tmp_path / "cls"
    det_dir.mkdir(), cls_dir.mkdir()
    frame_info = {}
    for f in frames:
        (det_dir / (f["token"] + ".json")).write_text(json.dumps(f["dets"]))
        (cls_dir / (f["token"] + ".json")).write_text(json.dumps(f["cls"]))
        frame_info[f["token"]] = {"prev": f["prev_token"] or "outside_the_split", "timestamp": f["timestamp"],
                                  "prev_timestamp": f["prev_timestamp"]}
    got = detfile.frames_from_json_dirs(str(det_dir), str(cls_dir), frame_info, [f["token"] for f in frames])
    assert got == frames                       # JSON floats round-trip exactly; an unknown prev = scene start
    path = str(tmp_path / "all.shdb")
    detfile.write_detection_file(path, got)
    assert detfile.DetectionFile(path).frames() == frames

    results = {f["token"]: f["cls"] for f in frames}
    out = tmp_path / "cp_val.json"
    formats.write_results(str(out), results)
    back = json.loads(out.read_text())
    assert back["results"] == results
    assert back["meta"] == {"use_camera": False, "use_lidar": True, "use_radar": False, "use_map": False,
                            "use_external": False}


def test_cls_info_is_the_global_frame_entry_not_the_sensor_box(tmp_path):
    """the emitted detection dicts must carry the cls_info (GLOBAL frame) translation /
    rotation / velocity, not the det_path (SENSOR frame) box the head's input is packed from; further keys survive."""
    frames = _frames(21, n_scenes=1, n_frames=3)
    path = str(tmp_path / "g.shdb")
    detfile.write_detection_file(path, frames)
    df = detfile.DetectionFile(path)
    differs = extras = 0
    for i, f in enumerate(frames):
        b, n = int(df.row_begin[i]), int(df.row_count[i])
        got = df.cls_info(range(b, b + n), f["token"])
        assert got == f["cls"]
        for d, c in zip(f["dets"], got):
            differs += c["translation"] != d[0:3]
            extras += "num_pts" in c
            assert df.rows[b:b + n][:, 0:3].tolist() == [x[0:3] for x in f["dets"]]
    assert differs > 0 and extras > 0
