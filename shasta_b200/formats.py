"""Data formats either side of the affinity path — host-side, numpy only.

* ``pack_detections``: per-frame detection JSON -> the zero-padded ``(max_objects, 11)`` box array
  ``[x, y, z, w, l, h, yaw, vx, vy, dt, score]`` the head reads (det3d/datasets/nuscenes/nuscenes.py:207-293).
* ``build_gt_affinity``: the ``(M+2, M+2)`` training target from the per-frame label file (``matched`` (N, K+2),
  ``newborn`` (K)) incl. the dead-track / false-positive sub-sampling (nuscenes.py:297-349).
* ``label_affinity``: those label matrices from the detection<->ground-truth associations of two consecutive frames
  (preprocessing/make_gt_shasta.py:88-157).
* ``frame_pair_example``: both frames of one sample -> the ``example`` keys ``Shasta.forward`` consumes.

Random sub-sampling draws from a ``random.Random``-compatible object in the reference's order (``sample`` for the
truncation to ``max_objects``, ``shuffle`` for the dead-track and FP subsets), so a seeded run reproduces the
reference's choice exactly.
"""
import random as _random

import numpy as np


def quaternion_yaw(q):
    """nuscenes.py:35-50 with pyquaternion's ``rotation_matrix`` written out: yaw of the box x axis after rotating by
    the (normalised) quaternion ``q = (w, x, y, z)``; returns ``np.array([yaw])``."""
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    if n > 0:
        q = q / n
    w, x, y, z = q
    v0 = 1.0 - 2.0 * (y * y + z * z)      # first column of the rotation matrix
    v1 = 2.0 * (x * y + z * w)
    return np.array([np.arctan2(v1, v0)])


def pack_detections(det_boxes, cls_info, max_objects, time_diff, det_type=None, rng=None):
    """nuscenes.py:213-246 (previous frame) / :255-293 (current frame).

    det_boxes: list of 12-float lists ``translation(3) size(3) rotation(4: w,x,y,z) velocity(2)``; cls_info: list of
    dicts with ``detection_name`` and ``detection_score``. Returns ``(boxes (max_objects, 11) float64, keep indices
    into the input lists, kept cls_info dicts, number of real rows)``. More than ``max_objects`` detections are
    sub-sampled with ``rng.sample`` and kept in their original order."""
    rng = rng if rng is not None else _random
    boxes = np.zeros((max_objects, 11))
    keep = [i for i in range(max_objects)]
    kept_cls = []
    if len(det_boxes) == 0:
        return boxes, keep, kept_cls, 0
    keep = []
    rows = []
    for i, (b, info) in enumerate(zip(det_boxes, cls_info)):
        if det_type is not None and info['detection_name'] not in det_type:
            continue
        score = info['detection_score']
        translation, size, rotation, velocity = np.array(b[:3]), np.array(b[3:6]), np.array(b[6:10]), np.array(b[10:12])
        rows.append(np.concatenate((translation, size, quaternion_yaw(rotation), velocity, np.array([time_diff]),
                                    np.array([score]))))
        kept_cls.append(info)
        keep.append(i)
    n = 0
    if len(rows) > 0:
        if len(rows) > max_objects:
            idx = rng.sample(range(len(rows)), max_objects)
            idx.sort()
            rows = [rows[i] for i in idx]
            kept_cls = [kept_cls[i] for i in idx]
            keep = [keep[i] for i in idx]
        n = len(rows)
        rows = np.array(rows)
        boxes[:rows.shape[0], :] = rows
    return boxes, keep, kept_cls, n


def build_gt_affinity(matched, newborn, prev_keep, keep, max_objects, has_prev, fp_ratio, dead_trk_ratio, rng=None):
    """nuscenes.py:297-349. ``matched`` (N_prev, K+2) / ``newborn`` (K) come from the frame's label file and refer
    to ALL detections of the two frames; ``prev_keep`` / ``keep`` are the indices ``pack_detections`` kept.
    Returns ``(gt (M+2, M+2) float64, num_prev_det_boxes or None, num_det_boxes)`` — the two counts are what the
    reference writes back into ``info`` (the previous count only when there is a previous frame)."""
    rng = rng if rng is not None else _random
    M = max_objects
    gt = np.zeros((M + 2, M + 2))
    num_prev = None
    if has_prev:
        gt[:len(prev_keep), :] = 0
        temp = matched[prev_keep]
        temp = temp[:, keep]
        gt[:len(prev_keep), :len(keep)] = temp
        gt[:len(prev_keep), -2] = matched[prev_keep, -2]                      # dead tracks
        gt[:len(prev_keep), -1] = 1 - gt[:len(prev_keep), :].sum(axis=1)      # FNs
        dead_trk = gt[:len(prev_keep), -2]
        fn = gt[:len(prev_keep), -1]
        prev_tp = gt[:len(prev_keep), :-2].sum(axis=1) + fn
        prev_tp_idx = list(np.nonzero(prev_tp == 1)[0])
        dead_trk_idx = list(np.nonzero(dead_trk == 1)[0])
        rng.shuffle(dead_trk_idx)
        num_keep_dead_trk = int(dead_trk_ratio * prev_tp.sum())
        temp_prev_keep = dead_trk_idx[:num_keep_dead_trk] + prev_tp_idx
        temp_prev_keep.sort()
        num_prev = len(temp_prev_keep)
        gt[:len(temp_prev_keep), :] = gt[temp_prev_keep, :]
        gt[len(temp_prev_keep):-2, :] = np.zeros((M - len(temp_prev_keep), M + 2))
    gt[-2, :len(keep)] = newborn[keep]                                        # newborns
    fp = 1 - gt[:, :len(keep)].sum(axis=0)                                    # FPs
    gt[-1, :len(keep)] = fp
    tp = gt[:-1, :len(keep)].sum(axis=0)
    tp_idx = list(np.nonzero(tp == 1)[0])
    fp_idx = list(np.nonzero(fp == 1)[0])
    rng.shuffle(fp_idx)
    num_keep_fp = int(fp_ratio * tp.sum())
    temp_keep = fp_idx[:num_keep_fp] + tp_idx
    temp_keep.sort()
    gt[:, :len(temp_keep)] = gt[:, temp_keep]
    gt[:, len(temp_keep):-2] = np.zeros((M + 2, M - len(temp_keep)))
    return gt, num_prev, len(temp_keep)


def label_affinity(tp_ind_pairs, frame_gt_ids, fn_inds, num_dets, prev=None):
    """Ground-truth affinity of one frame from the detection<->GT association, as preprocessing/make_gt_shasta.py:88-157
    defines it (pinned against that script's own output: oracle/make_labelaff_golden.py, tests/golden/labelaff_*.json).

    ``tp_ind_pairs``: detection index -> GT index of the true positives, ``frame_gt_ids``: instance id per GT index,
    ``fn_inds``: GT indices nobody detected, ``num_dets`` = K. ``prev = (prev_tp_ind_pairs, prev_gt_ids, N)`` of the
    previous frame, or None at the start of a scene. Returns ``(matched (N, K+2) or None, newborn (K,))``:
    a previous true positive is matched to the current true positive of the same instance, else it is an FN track
    (column K+1) when that instance is still annotated but undetected, else dead (column K); previous false positives
    are dead; a current true positive whose instance had no previous true positive is newborn."""
    K = num_dets
    newborn = np.zeros((K,))
    if prev is None:
        newborn[[k for k in tp_ind_pairs if k < K]] = 1
        return None, newborn
    prev_tp_ind_pairs, prev_gt_ids, N = prev
    matched = np.zeros((N, K + 2))
    # instance id -> the FIRST previous detection that was a true positive of it (the script resolves ids with
    # list.index, i.e. first occurrence in association order)
    prev_det_of = {}
    prev_tp = [(det, prev_gt_ids[g]) for det, g in prev_tp_ind_pairs.items()]
    for det, inst in prev_tp:
        prev_det_of.setdefault(inst, det)
    continued = set()
    for det, g in tp_ind_pairs.items():
        inst = frame_gt_ids[g]
        if inst in prev_det_of:
            matched[prev_det_of[inst], det] = 1
            continued.add(inst)
        else:
            newborn[det] = 1
    # still annotated, but undetected in this frame: the track goes on as a false negative
    first_gt_of = {}
    for g, inst in enumerate(frame_gt_ids):
        first_gt_of.setdefault(inst, g)
    undetected = set(fn_inds)
    for det, inst in prev_tp:
        if inst not in continued and first_gt_of.get(inst, -1) in undetected:
            matched[det, K + 1] = 1
    matched[:, K] = 1 - matched.sum(axis=1)      # everything else died (false positives included)
    return matched, newborn


def frame_pair_example(prev_dets, prev_cls, cur_dets, cur_cls, max_objects, time_diff, det_type=None, rng=None):
    """Both frames of one sample -> float32 ``(1, M, 11)`` arrays under the keys ``Shasta.forward`` reads, plus the
    bookkeeping the decode needs (kept cls_info lists, real counts). ``prev_dets is None`` = first frame of a scene."""
    if prev_dets is None:
        pb, pk, pc, pn = np.zeros((max_objects, 11)), list(range(max_objects)), [], 0
    else:
        pb, pk, pc, pn = pack_detections(prev_dets, prev_cls, max_objects, time_diff, det_type, rng)
    cb, ck, cc, cn = pack_detections(cur_dets, cur_cls, max_objects, time_diff, det_type, rng)
    return {"prev_det_boxes": pb.astype(np.float32)[None], "det_boxes": cb.astype(np.float32)[None],
            "prev_cls_det_boxes": pc, "cls_det_boxes": cc, "num_prev_det_boxes": pn, "num_det_boxes": cn,
            "prev_keep": pk, "keep": ck}


def annos_from_decode(prev_cls, cur_cls, prev_state, fn_dead_prob, det_state, det_fp_prob, token, time_lag):
    """The dict-level half of the eval loop (tools/nusc_shasta/eval.py:126-181) on the arrays ``shasta_decode_f32``
    returns for ONE frame pair (prev_state: 0 keep / 1 dead / 2 FN; det_state: 0 keep / 1 newborn / 2 dropped FP).
    Mutates the cls_info dicts like the reference (FN boxes are propagated by ``velocity * time_lag``, ``newborn`` /
    ``ref_detection_score`` are attached) and returns ``(annos for this token, dead previous indices, kept current
    indices)`` — the last two feed the reference's ``dead_tracker`` post-pass (eval.py:175-181).
    ``fn_dead_prob[n]`` = matched1[n, -2] of an FN row, ``det_fp_prob[k]`` = matched2[-1, k] of a kept detection (float32,
    as the decode kernel stores them): ``ref_detection_score = 1 - value`` is formed here in Python double, exactly like
    the reference's ``1 - tensor.item()`` (eval.py:148,169), so the emitted JSON is bit-identical."""
    annos, fn_annos, dead_idx, keep_dets = [], [], [], []
    for n in range(len(prev_cls)):
        if prev_state[n] == 1:
            dead_idx.append(n)
        elif prev_state[n] == 2:
            box = prev_cls[n]
            box["translation"][:2] = [t + time_lag * v for t, v in zip(box["translation"][:2], box["velocity"])]
            box["FN"] = True
            box["token"] = token
            box["ref_detection_score"] = 1 - float(fn_dead_prob[n])
            fn_annos.append(box)
    for k in range(len(cur_cls)):
        if det_state[k] == 2:
            continue
        if det_state[k] == 1:
            cur_cls[k]["newborn"] = True
        cur_cls[k]["ref_detection_score"] = 1 - float(det_fp_prob[k])
        keep_dets.append(k)
        annos.append(cur_cls[k])
    annos.extend(fn_annos)
    return annos, dead_idx, keep_dets


def mark_dead(results, dead_tracker):
    """eval.py:175-181: after all frames, flag the kept detections of a token that the NEXT frame declared dead."""
    for token in results.keys():
        dead_idx = dead_tracker[token]['dead_idx']
        keep_idx = dead_tracker[token]['keep_idx']
        for i in dead_idx:
            if i in keep_idx:
                results[token][keep_idx.index(i)]['dead'] = True
    return results


def write_results(path, results):
    """tools/nusc_shasta/eval.py:184-193: the ``cp_<split>.json`` the downstream tracker reads - ``results`` is
    ``{token: annos}`` (``annos_from_decode`` + ``mark_dead``), ``meta`` the fixed lidar-only modality block."""
    import json
    nusc_annos = {"results": results,
                  "meta": {"use_camera": False, "use_lidar": True, "use_radar": False, "use_map": False,
                           "use_external": False}}
    with open(path, "w") as f:
        json.dump(nusc_annos, f)
    return nusc_annos
