"""BASELINE.json configs[2] as a parity case: several class models (different max_obj, own weights) over a batch of
scenes, ragged batches, decode blocks gathered into scene order - every frame pair of every class against the CPU
oracle (O.forward + O.decode): the discrete decisions must be identical, the scores within 1e-4. The same workload
served from a resident ring with step graphs must reproduce the eager result bit for bit."""
import numpy as np
import pytest
import torch

from oracle import shasta_oracle as O
from shasta_b200 import multiclass, synthetic
from tests import gpu_util as G

pytestmark = pytest.mark.gpu

CLASSES = (("bus", 20), ("tiny", 6), ("bicycle", 50))
HW = 24


def _lanes(step_graphs):
    pc_start = (-HW * 0.3, -HW * 0.3)
    lanes, weights = [], {}
    for i, (name, M) in enumerate(CLASSES):
        weights[name] = synthetic.make_weights(M, seed=10 + i, peaky=600.0)
        lanes.append(multiclass.ClassLane(name, G.make_model(M, pc_start, weights[name]), step_graphs=step_graphs))
    return lanes, weights, pc_start


def _oracle_states(m1, m2, n_prev, n_det, M):
    d = O.decode(m1, m2, n_prev, n_det)
    prev_state = np.full(M, -1, np.int32)
    prev_state[:n_prev] = [1 if n in d["dead"] else 2 if n in d["fn"] else 0 for n in range(n_prev)]
    fn_score = np.zeros(M, np.float32)        # raw matched1[n,-2] of the FN rows (the kernel stores the raw value)
    for n in d["fn"]:
        fn_score[n] = float(m1[n, -2])
    det_state = np.full(M, -1, np.int32)
    det_state[:n_det] = 2
    det_score = np.zeros(M, np.float32)
    for k, nb in zip(d["keep_dets"], d["newborn"]):
        det_state[k] = 1 if nb else 0
        det_score[k] = float(m2[-1, k])         # raw matched2[-1,k]
    return prev_state, fn_score, det_state, det_score


def test_three_class_sequence_batch_matches_oracle():
    lengths = [3, 2, 4]
    lanes, weights, pc_start = _lanes(step_graphs=False)
    prov = multiclass.SyntheticProvider(CLASSES, lengths, HW, G.DEV, seed=5)
    res = multiclass.run_sequence_batch(lanes, lengths, prov, {"bus": 4, "tiny": 9, "bicycle": 2})
    decisions = 0
    for name, M in CLASSES:
        wt = O.weights_to_torch(weights[name])
        bx = prov.boxes[name]
        assert len(res[name]) == len(lengths)
        for s, n in enumerate(lengths):
            got = multiclass.decode_fields(res[name][s].cpu())
            assert tuple(res[name][s].shape) == (n, 6, M)
            for f in range(n):
                i = prov.pair_index(s, f)
                maps = prov.maps[name]
                m1, m2 = O.forward(wt, maps[i + s + 1:i + s + 2].cpu(), maps[i + s:i + s + 1].cpu(),
                                   bx["det_boxes"][i:i + 1].cpu().clone(), bx["prev_det_boxes"][i:i + 1].cpu(),
                                   pc_start=pc_start)
                n_prev, n_det = int(bx["n_prev"][i]), int(bx["n_det"][i])
                ps, fs, ds, sc = _oracle_states(m1[0], m2[0], n_prev, n_det, M)
                assert np.array_equal(got["prev_state"][f].numpy()[:n_prev], ps[:n_prev]), (name, s, f)
                assert np.array_equal(got["det_state"][f].numpy()[:n_det], ds[:n_det]), (name, s, f)
                fn_rows = ps[:n_prev] == 2
                assert np.allclose(got["fn_dead_prob"][f].numpy()[:n_prev][fn_rows], fs[:n_prev][fn_rows], atol=1e-4)
                kept = ds[:n_det] != 2
                assert np.allclose(got["det_fp_prob"][f].numpy()[:n_det][kept], sc[:n_det][kept], atol=1e-4)
                decisions += int((ps[:n_prev] > 0).sum() + (ds[:n_det] > 0).sum())
    assert decisions > 0, "weights not peaky enough: no dead / FN / newborn / FP decision was exercised"


def test_ring_with_step_graphs_equals_eager():
    lengths = [5, 3, 6]          # 14 frame pairs: batches of 4 -> 4, 4, 4, 2; ring of 2 -> slots reused, graphs replayed
    out = []
    for graphs in (False, True):
        lanes, _, _ = _lanes(step_graphs=graphs)
        prov = multiclass.SyntheticProvider(CLASSES, lengths, HW, G.DEV, seed=7, ring=2, batch_pairs=4)
        out.append(multiclass.run_sequence_batch(lanes, lengths, prov, 4))
        if graphs:
            assert all(len(l._graphs) == 3 for l in lanes)      # (slot 0, B=4), (slot 1, B=4), (slot 1, B=2)
    for name, _ in CLASSES:
        for a, b in zip(out[0][name], out[1][name]):
            assert torch.equal(a, b)
    # the ring serves the same batch again: rows 0-3 and 8-11 of the rank's stream are identical
    flat = torch.cat(out[1]["bus"], dim=0)
    assert torch.equal(flat[0:4], flat[8:12]) and not torch.equal(flat[0:4], flat[4:8])


def test_decode_out_argument_checks_shape():
    lanes, _, _ = _lanes(step_graphs=False)
    model = lanes[0].model
    m1 = torch.rand((2, 20, 22), device=G.DEV)
    m2 = torch.rand((2, 22, 20), device=G.DEV)
    with pytest.raises(ValueError):
        model.decode(m1, m2, [3, 4], [5, 6], out=torch.empty((6, 2, 21), dtype=torch.int32, device=G.DEV))
    blk = torch.empty((6, 2, 20), dtype=torch.int32, device=G.DEV)
    a = model.decode(m1, m2, [3, 4], [5, 6], out=blk)
    b = model.decode(m1, m2, [3, 4], [5, 6])
    assert all(torch.equal(a[k], b[k]) for k in b)
    assert a["prev_state"].data_ptr() == blk.data_ptr()


def test_detection_file_provider_matches_json_loop(tmp_path):
    """Binary detection file -> DetectionFileProvider -> batched, per-class lanes -> annotations, against the
    reference-shaped JSON loop (pipeline.run_class_sequence, batch size 1, itself checked against the CPU oracle in
    tests/test_gpu_pipeline.py): same surviving detections, flags and scores for every token of every class."""
    import copy

    from shasta_b200 import detfile, pipeline
    from tests.test_gpu_pipeline import _scene

    M, H, W = 20, 32, 32
    pc_start = (-W * 0.3, -H * 0.3)
    rng = np.random.default_rng(2)
    frames = []
    for s, nfr in enumerate((5, 3, 6)):
        for f in _scene(20 + s, nfr, M, extent=8.0):
            f["token"] = "s%d_%s" % (s, f["token"])
            f["prev_token"] = "" if f["prev_token"] == "" else "s%d_%s" % (s, f["prev_token"])
            f["timestamp"] += s * 10_000_000
            f["prev_timestamp"] += s * 10_000_000
            for c in f["cls"]:
                c["sample_token"] = f["token"]
                if rng.random() < 0.35:
                    c["detection_name"] = "truck"
            frames.append(f)
    path = str(tmp_path / "dets.shdb")
    detfile.write_detection_file(path, frames)
    df = detfile.DetectionFile(path)

    g = torch.Generator().manual_seed(9)
    cur_maps = torch.relu(torch.randn((len(frames), H, W, 64), generator=g)).to(G.DEV)
    first_prev = torch.relu(torch.randn((H, W, 64), generator=g)).to(G.DEV)

    def prev_map(i):
        p = int(df.prev_index[i])
        return first_prev if p < 0 else cur_maps[p]

    def maps_for(name, idx):
        return cur_maps[idx].contiguous(), torch.stack([prev_map(i) for i in idx])

    classes = {"car": ["car"], "truck": ["truck"]}
    lanes, models = [], {}
    for k, name in enumerate(classes):
        models[name] = G.make_model(M, pc_start, synthetic.make_weights(M, seed=30 + k, peaky=500.0))
        lanes.append(multiclass.ClassLane(name, models[name]))
    prov = multiclass.DetectionFileProvider(df, classes, {n: M for n in classes}, maps_for, G.DEV)
    assert prov.scene_lengths == [5, 3, 6]
    blocks = multiclass.run_sequence_batch(lanes, prov.scene_lengths, prov, 4)

    flags = 0
    for name, det_type in classes.items():
        got = prov.annotations(name, blocks[name])
        want = pipeline.run_class_sequence(
            models[name], copy.deepcopy(frames),
            lambda tok: (cur_maps[df.frame_index(tok)][None], prev_map(df.frame_index(tok))[None]), det_type=det_type)
        assert set(got) == set(want)
        for tok in want:
            assert len(got[tok]) == len(want[tok]), (name, tok)
            for x, y in zip(got[tok], want[tok]):
                assert x["translation"] == y["translation"] and x["detection_name"] == y["detection_name"]
                assert x.get("newborn") == y.get("newborn") and x.get("dead") == y.get("dead") and x.get("FN") == y.get("FN")
                assert abs(x["ref_detection_score"] - y["ref_detection_score"]) < 1e-6
                flags += bool(x.get("newborn")) + bool(x.get("dead")) + bool(x.get("FN"))
    assert flags > 0


def test_step_graphs_follow_weight_updates_and_data_writes():
    """a lane's captured step graph bakes in the packed-weight buffer. After load_state_dict (new packed
    buffer) and after an in-place ``param.data`` update + ``invalidate_packed()`` the lane must produce what an eager
    model with the same weights produces - not replay the stale capture."""
    M, H = 20, 32
    pc_start = (-H * 0.3, -H * 0.3)
    data = synthetic.make_frame_pairs(3, M, H, H, 77, pc_start=pc_start)
    w_a = synthetic.make_weights(M, seed=41, peaky=300.0)
    w_b = synthetic.make_weights(M, seed=42, peaky=300.0)
    model = G.make_model(M, pc_start, w_a)
    lane = multiclass.ClassLane("x", model, step_graphs=True)
    batch = {"bev": G.t(data["bev"]), "prev_bev": G.t(data["prev_bev"]), "det_boxes": G.t(data["det_boxes"]),
             "prev_det_boxes": G.t(data["prev_det_boxes"]),
             "n_prev": torch.from_numpy(data["n_prev"].astype(np.int32)).to(G.DEV),
             "n_det": torch.from_numpy(data["n_det"].astype(np.int32)).to(G.DEV)}

    def eager(weights):
        ref = G.make_model(M, pc_start, weights)
        with torch.no_grad():
            m1, m2 = ref.affinity(batch["bev"], batch["prev_bev"], batch["det_boxes"].clone(), batch["prev_det_boxes"])
            d = ref.decode(m1, m2, batch["n_prev"], batch["n_det"])
        return torch.stack([d[k].view(torch.int32) if k.endswith("prob") else d[k] for k in multiclass.DECODE_FIELDS])

    with torch.no_grad():
        first = lane.step(batch).clone()
        assert torch.equal(first, eager(w_a))
        assert torch.equal(lane.step(batch), first)                      # replay of the captured graph
        from shasta_b200 import load_matching_state_dict
        load_matching_state_dict(model, {k: torch.from_numpy(v) for k, v in w_b.items()})
        model.invalidate_packed()
        second = lane.step(batch).clone()
        assert torch.equal(second, eager(w_b)) and not torch.equal(second, first)
        # in-place update through .data does not bump the tensor version: invalidate_packed() is the documented hook
        model.aff[10].weight.data.mul_(0.5)
        model.invalidate_packed()
        w_c = {k: v.copy() for k, v in w_b.items()}
        w_c["aff.10.weight"] = w_c["aff.10.weight"] * np.float32(0.5)
        assert torch.equal(lane.step(batch), eager(w_c))
