"""``PubTrackerMerged`` — the greedy centre-distance ID tracker downstream of the affinity head
(tools/nusc_shasta/pub_tracker_merged.py:55-225): same constructor / ``reset`` / ``step_centertrack(results,
time_lag)`` interface and the same outputs (``tracking_id``, ``age``, ``active``, ``ref_detection_score``; it consumes the
``newborn`` / ``dead`` flags written by the decode, eval.py:126-181), but its own organisation: the live tracks and
the frame's detections are struct-of-arrays tables, the per-class life-cycle rules are array expressions over them, and
for every class present in the frame the distance matrix, the validity mask and the greedy assignment
(track_utils.py:3-14) run as ONE launch of ``shasta_greedy_assign_f32`` (one thread block per class). There is no CPU
fallback for that step; ``hungarian=True`` is not implemented (the reference's eval.py:252 and pub_test.py use
``hungarian=False``). Parity: tests/test_tracker.py against fixtures written by the unmodified reference tracker.
"""
import ctypes

import numpy as np
import torch

from . import _cabi

NUSCENES_TRACKING_NAMES = ['bicycle', 'bus', 'car', 'motorcycle', 'pedestrian', 'trailer', 'truck']

# pub_tracker_merged.py:24-32
NUSCENE_CLS_VELOCITY_ERROR = {'car': 2, 'truck': 2, 'bus': 4, 'trailer': 2, 'pedestrian': 0.75, 'motorcycle': 2,
                              'bicycle': 1.5}
# pub_tracker_merged.py:34-42
TRK_REF = {
    'bicycle': {'alpha': 0.5, 'beta': 0.4, 'ref': True},
    'bus': {'alpha': 0.5, 'beta': 0.7, 'ref': True},
    'car': {'alpha': 0.5, 'beta': 0.5, 'ref': True},
    'motorcycle': {'alpha': 0.5, 'beta': 0.5, 'ref': True},
    'pedestrian': {'alpha': 0.5, 'beta': 0.5, 'ref': True},
    'trailer': {'alpha': 0.5, 'beta': 0.4, 'ref': True},
    'truck': {'alpha': 0.5, 'beta': 0.5, 'ref': True},
}


def greedy_assign_batch(dets, tracks, max_diff, det_cat, track_cat, device="cuda:0"):
    """Lists (one entry per problem) of dets (N,2) f32, tracks (M,2) f32, max_diff (N,) f32, det_cat (N,) i32,
    track_cat (M,) i32 -> lists of match (N,) int32 (-1 = unmatched), det_near (N,) bool, track_near (M,) bool."""
    P = len(dets)
    if P == 0:
        return [], [], []
    nmax = max(1, max(len(d) for d in dets))
    mmax = max(1, max(len(t) for t in tracks))
    hd = np.zeros((P, nmax, 2), np.float32)
    ht = np.zeros((P, mmax, 2), np.float32)
    hm = np.zeros((P, nmax), np.float32)
    hdc = np.zeros((P, nmax), np.int32)
    htc = np.zeros((P, mmax), np.int32)
    nd = np.zeros(P, np.int32)
    nt = np.zeros(P, np.int32)
    for p in range(P):
        n, m = len(dets[p]), len(tracks[p])
        nd[p], nt[p] = n, m
        if n:
            hd[p, :n], hm[p, :n], hdc[p, :n] = dets[p], max_diff[p], det_cat[p]
        if m:
            ht[p, :m], htc[p, :m] = tracks[p], track_cat[p]
    dev = torch.device(device)
    g = [torch.from_numpy(a).to(dev) for a in (hd, ht, hm, hdc, htc, nd, nt)]
    match = torch.empty((P, nmax), dtype=torch.int32, device=dev)
    dnear = torch.empty((P, nmax), dtype=torch.int32, device=dev)
    tnear = torch.empty((P, mmax), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.lib().shasta_greedy_assign_f32(
            *[t.data_ptr() for t in g], P, nmax, mmax, match.data_ptr(), dnear.data_ptr(), tnear.data_ptr(),
            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _cabi.check(rc, "shasta_greedy_assign_f32")
    match, dnear, tnear = match.cpu().numpy(), dnear.cpu().numpy(), tnear.cpu().numpy()
    return ([match[p, :nd[p]] for p in range(P)], [dnear[p, :nd[p]] != 0 for p in range(P)],
            [tnear[p, :nt[p]] != 0 for p in range(P)])


class _Tracks:
    """Struct-of-arrays state of the live tracks (all classes): one row per track, columns instead of per-track dicts."""

    __slots__ = ("cls", "ct", "tracking", "score", "age", "active", "tid", "dead", "payload")

    def __init__(self, n=0):
        self.cls = np.zeros(n, np.int64)
        self.ct = np.zeros((n, 2))              # last centre (advanced by -tracking while a track coasts)
        self.tracking = np.zeros((n, 2))        # velocity * -time_lag of the detection the track was last fed with
        self.score = np.zeros(n)                # ref_detection_score
        self.age = np.zeros(n, np.int64)
        self.active = np.zeros(n, np.int64)
        self.tid = np.zeros(n, np.int64)
        self.dead = np.zeros(n, bool)           # the decode of the NEXT frame declared the detection dead
        self.payload = []                       # the detection dict every emitted record is built from

    def __len__(self):
        return len(self.tid)

    @staticmethod
    def concat(parts):
        out = _Tracks()
        if parts:
            for k in ("cls", "ct", "tracking", "score", "age", "active", "tid", "dead"):
                setattr(out, k, np.concatenate([getattr(p_, k) for p_ in parts]))
            out.payload = [d for p_ in parts for d in p_.payload]
        return out

    def take(self, idx):
        out = _Tracks()
        for k in ("cls", "ct", "tracking", "score", "age", "active", "tid", "dead"):
            setattr(out, k, getattr(self, k)[idx])
        out.payload = [self.payload[int(i)] for i in idx]
        return out


_ALPHA = np.array([TRK_REF[n]["alpha"] for n in NUSCENES_TRACKING_NAMES])
_BETA = np.array([TRK_REF[n]["beta"] for n in NUSCENES_TRACKING_NAMES])
_USE_REF = np.array([TRK_REF[n]["ref"] for n in NUSCENES_TRACKING_NAMES])
_MAX_DIFF = np.array([NUSCENE_CLS_VELOCITY_ERROR[n] for n in NUSCENES_TRACKING_NAMES], np.float32)
_CLS_INDEX = {n: i for i, n in enumerate(NUSCENES_TRACKING_NAMES)}


class PubTrackerMerged(object):
    """Same interface as the reference class (constructor, ``reset``, ``step_centertrack(results, time_lag)`` returning
    the list of track dicts with ``tracking_id`` / ``age`` / ``active`` / ``ref_detection_score`` / ``ct``), different
    inside: the frame's detections and the live tracks are columns of arrays, every class is handled with array
    expressions (score blending, ageing, id assignment, coasting), the assignment itself is one GPU launch for all
    classes, and dicts are only touched once per emitted record at the end."""

    def __init__(self, hungarian=False, max_age=0, device="cuda:0"):
        if hungarian:
            raise NotImplementedError("shasta_b200.PubTrackerMerged implements the greedy assignment "
                                      "(hungarian=False, the reference's setting in eval.py / pub_test.py)")
        self.hungarian = hungarian
        self.max_age = max_age
        self.device = device
        self.NUSCENE_CLS_VELOCITY_ERROR = NUSCENE_CLS_VELOCITY_ERROR
        self.reset()

    def reset(self):
        self.id_count = 0
        self._state = _Tracks()
        self.tracks = []

    def step_centertrack(self, results, time_lag):
        """pub_tracker_merged.py:72-225. ``results``: list of detection dicts of one frame (all classes)."""
        if len(results) == 0:
            self._state, self.tracks = _Tracks(), []
            return []
        # ---- the frame as columns (pub_tracker_merged.py:80-92); detections of other classes are not tracked
        cls_all = np.array([_CLS_INDEX.get(d["detection_name"], -1) for d in results])
        sel = np.flatnonzero(cls_all >= 0)
        dets = _Tracks(len(sel))
        dets.cls = cls_all[sel]
        dets.payload = [results[int(i)] for i in sel]
        if len(sel):
            dets.ct = np.array([d["translation"][:2] for d in dets.payload], dtype=np.float64).reshape(-1, 2)
            dets.tracking = np.array([d["velocity"][:2] for d in dets.payload], dtype=np.float64).reshape(-1, 2) * -1 * time_lag
        det_conf = np.array([d["detection_score"] for d in dets.payload], dtype=np.float64)
        tp_prob = np.array([d.get("ref_detection_score", 0.0) for d in dets.payload], dtype=np.float64)
        newborn = np.array(["newborn" in d for d in dets.payload], dtype=bool)
        dets.dead = np.array(["dead" in d for d in dets.payload], dtype=bool)
        old = self._state
        # float32 operands of the assignment, exactly as the reference forms them (:104-111)
        det_pos = (dets.ct + dets.tracking.astype(np.float32)).astype(np.float32)
        trk_pos = old.ct.astype(np.float32)

        # ---- one launch: every class that has both detections and tracks (:113-137)
        classes = [c for c in range(len(NUSCENES_TRACKING_NAMES)) if (dets.cls == c).any()]
        di = {c: np.flatnonzero(dets.cls == c) for c in classes}
        ti = {c: np.flatnonzero(old.cls == c) for c in classes}
        todo = [c for c in classes if len(ti[c]) > 0]
        match, dnear, tnear = greedy_assign_batch(
            [det_pos[di[c]] for c in todo], [trk_pos[ti[c]] for c in todo], [_MAX_DIFF[dets.cls[di[c]]] for c in todo],
            [dets.cls[di[c]].astype(np.int32) for c in todo], [old.cls[ti[c]].astype(np.int32) for c in todo], self.device)
        solved = {c: (match[k], dnear[k], tnear[k]) for k, c in enumerate(todo)}

        # ---- bookkeeping per class, on columns (:139-222). Output order per class: matched detections, new tracks,
        # coasting tracks - the order the reference appends them in
        parts = []
        for c in classes:
            d_idx, t_idx = di[c], ti[c]
            beta, alpha, use_ref = _BETA[c], _ALPHA[c], _USE_REF[c]
            if len(t_idx) > 0:
                m, det_near, track_near = solved[c]
            else:
                m, det_near, track_near = -np.ones(len(d_idx), np.int64), np.zeros(len(d_idx), bool), np.zeros(0, bool)
            hit = m >= 0
            # matched detections keep the id, blend the score, restart the age
            md, mt = d_idx[hit], t_idx[m[hit]]
            a = dets.take(md)
            a.tid = old.tid[mt]
            a.score = ((tp_prob[md] > alpha) * beta * det_conf[md] + (1 - beta) * old.score[mt]) if use_ref else det_conf[md]
            a.age = np.ones(len(md), np.int64)
            a.active = old.active[mt] + 1
            # unmatched detections start a track unless the head did not call them newborn and a track is in range
            um = ~hit
            if len(t_idx) > 0:
                um &= newborn[d_idx] | ~det_near
            ud = d_idx[um]
            b = dets.take(ud)
            b.tid = self.id_count + 1 + np.arange(len(ud), dtype=np.int64)
            self.id_count += len(ud)
            b.score = beta * det_conf[ud] if use_ref else det_conf[ud]
            b.age = np.ones(len(ud), np.int64)
            b.active = np.ones(len(ud), np.int64)
            # unmatched tracks coast while young enough, unless declared dead with a detection in range
            taken = np.zeros(len(t_idx), bool)
            taken[m[hit]] = True
            coast = ~taken & ~(old.dead[t_idx] & track_near) & (old.age[t_idx] < self.max_age)
            ut = t_idx[coast]
            e = old.take(ut)
            e.age = e.age + 1
            e.active = np.zeros(len(ut), np.int64)
            if use_ref:
                e.score = (1 - beta) * e.score
            e.ct = e.ct + e.tracking * -1          # move forward
            parts += [a, b, e]
        new = _Tracks.concat(parts)
        # ---- records: the reference's per-track dict view, written once per emitted track
        ret = []
        for k, d in enumerate(new.payload):
            d["ct"] = new.ct[k]
            d["tracking"] = new.tracking[k]
            d["label_preds"] = int(new.cls[k])
            d["tracking_id"] = int(new.tid[k])
            d["ref_detection_score"] = float(new.score[k])
            d["age"] = int(new.age[k])
            d["active"] = int(new.active[k])
            ret.append(d)
        self._state, self.tracks = new, ret
        return ret
