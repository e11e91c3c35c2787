"""The per-class inference loop of tools/nusc_shasta/eval.py:110-181 on top of the CUDA head: frame-pair packing
(formats.py) -> ``Shasta.forward`` -> device decode -> per-token annotation lists (+ the ``dead`` post-pass), ready for
``PubTrackerMerged``. ``maps_for(token)`` supplies the two 64-channel channels-last BEV maps of a sample (current,
previous) - in the reference they come from the frozen trunk + shared_conv."""
import copy

import numpy as np
import torch

from . import formats


def run_class_sequence(model, frames, maps_for, det_type=None, device="cuda:0"):
    """``frames``: list of dicts ``{token, prev_token ('' at a scene start), timestamp, prev_timestamp (both in
    microseconds like the nuScenes frame info), dets, cls}`` with ``dets`` / ``cls`` in the per-frame detection JSON
    layout (formats.pack_detections). Detections of a frame are looked up by token, like the reference reads
    ``<token>.json``. Returns ``{token: annos}`` (eval.py ``nusc_annos['results']``)."""
    by_token = {f["token"]: f for f in frames}
    results, dead_tracker = {}, {}
    M = model.max_obj
    for f in frames:
        token, prev_token = f["token"], f["prev_token"]
        dead_tracker.setdefault(token, {"dead_idx": [], "keep_idx": []})
        time_diff = 1e-6 * f["timestamp"] - 1e-6 * f["prev_timestamp"]
        prev = by_token.get(prev_token) if prev_token != "" else None
        ex = formats.frame_pair_example(None if prev is None else copy.deepcopy(prev["dets"]),
                                        None if prev is None else copy.deepcopy(prev["cls"]),
                                        copy.deepcopy(f["dets"]), copy.deepcopy(f["cls"]), M, time_diff, det_type)
        bev, prev_bev = maps_for(token)
        example = {"det_boxes": torch.from_numpy(ex["det_boxes"]).to(device),
                   "prev_det_boxes": torch.from_numpy(ex["prev_det_boxes"]).to(device),
                   "bev_feature": bev, "prev_bev_feature": prev_bev}
        with torch.no_grad():
            m1, m2, _ = model(example, train_mode=False)
            n_prev, n_det = len(ex["prev_cls_det_boxes"]), len(ex["cls_det_boxes"])
            dec = {k: v[0].cpu().numpy() for k, v in model.decode(m1, m2, [n_prev], [n_det]).items()}
        time_lag = float(ex["prev_det_boxes"][0, 0, 9])
        annos, dead_idx, keep = formats.annos_from_decode(ex["prev_cls_det_boxes"], ex["cls_det_boxes"],
                                                          dec["prev_state"], dec["fn_dead_prob"], dec["det_state"],
                                                          dec["det_fp_prob"], token, time_lag)
        if n_prev > 0:
            dead_tracker.setdefault(prev_token, {"dead_idx": [], "keep_idx": []})["dead_idx"].extend(dead_idx)
        if n_det > 0:
            dead_tracker[token]["keep_idx"] = keep
        results[token] = annos
    return formats.mark_dead(results, dead_tracker)
