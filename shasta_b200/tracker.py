"""``PubTrackerMerged`` — the greedy centre-distance ID tracker downstream of the affinity head
(tools/nusc_shasta/pub_tracker_merged.py:55-225), same constructor / ``reset`` / ``step_centertrack(results,
time_lag)`` interface and the same per-detection dict bookkeeping (``tracking_id``, ``age``, ``active``,
``ref_detection_score``, the ``newborn`` / ``dead`` flags written by the decode, eval.py:126-181).

What moved to the GPU is the part that is arithmetic: for every class present in the frame the distance matrix, the
validity mask and the greedy assignment (track_utils.py:3-14) run as ONE launch of ``shasta_greedy_assign_f32`` (one
thread block per class). There is no CPU fallback for that step; ``hungarian=True`` is not implemented (the
reference's eval.py:252 and pub_test.py use ``hungarian=False``).
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


class PubTrackerMerged(object):
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
        self.tracks = []

    def step_centertrack(self, results, time_lag):
        """pub_tracker_merged.py:72-225. ``results``: list of detection dicts of one frame (all classes)."""
        if len(results) == 0:
            self.tracks = []
            return []
        # ---- per class: operands of the assignment (pub_tracker_merged.py:80-120)
        groups = []
        for nusc_name in NUSCENES_TRACKING_NAMES:
            temp = []
            for det in results:
                if det['detection_name'] != nusc_name:
                    continue
                det['ct'] = np.array(det['translation'][:2])
                det['tracking'] = np.array(det['velocity'][:2]) * -1 * time_lag
                det['label_preds'] = NUSCENES_TRACKING_NAMES.index(det['detection_name'])
                temp.append(det)
            if len(temp) == 0:
                continue
            curr_tracks = [track for track in self.tracks if track['detection_name'] == nusc_name]
            dets = np.array([det['ct'] + det['tracking'].astype(np.float32) for det in temp], np.float32)
            item_cat = np.array([item['label_preds'] for item in temp], np.int32)
            track_cat = np.array([track['label_preds'] for track in curr_tracks], np.int32)
            max_diff = np.array([self.NUSCENE_CLS_VELOCITY_ERROR[box['detection_name']] for box in temp], np.float32)
            tracks = np.array([pre_det['ct'] for pre_det in curr_tracks], np.float32).reshape(-1, 2)
            groups.append((nusc_name, temp, curr_tracks, dets, tracks, max_diff, item_cat, track_cat))
        # ---- one launch for the classes that have tracks to match against
        todo = [g for g in groups if len(g[4]) > 0]
        match, dnear, tnear = greedy_assign_batch([g[3] for g in todo], [g[4] for g in todo], [g[5] for g in todo],
                                                  [g[6] for g in todo], [g[7] for g in todo], self.device)
        solved = {g[0]: (match[i], dnear[i], tnear[i]) for i, g in enumerate(todo)}
        # ---- bookkeeping (pub_tracker_merged.py:139-222)
        ret = []
        for nusc_name, curr_results, curr_tracks, dets, tracks, max_diff, item_cat, track_cat in groups:
            if len(tracks) > 0:
                m, det_near, track_near = solved[nusc_name]
                matches = [(i, int(j)) for i, j in enumerate(m) if j >= 0]
            else:
                assert len(curr_tracks) == 0
                matches, det_near, track_near = [], None, None
            matched_d = set(i for i, _ in matches)
            matched_t = set(j for _, j in matches)
            unmatched_dets = [d for d in range(dets.shape[0]) if d not in matched_d]
            unmatched_tracks = [d for d in range(tracks.shape[0]) if d not in matched_t]
            for i, j in matches:
                track = curr_results[i]
                track['tracking_id'] = curr_tracks[j]['tracking_id']
                if TRK_REF[track['detection_name']]['ref']:
                    alpha, beta = TRK_REF[track['detection_name']]['alpha'], TRK_REF[track['detection_name']]['beta']
                    prev_track_conf = curr_tracks[j]['ref_detection_score']
                    tp_prob = track['ref_detection_score']
                    det_conf = track['detection_score']
                    track['ref_detection_score'] = (tp_prob > alpha) * beta * det_conf + (1 - beta) * prev_track_conf
                else:
                    track['ref_detection_score'] = track['detection_score']
                track['age'] = 1
                track['active'] = curr_tracks[j]['active'] + 1
                ret.append(track)
            for i in unmatched_dets:
                track = curr_results[i]
                if len(tracks) > 0:
                    if 'newborn' not in track.keys() and det_near[i]:
                        continue
                self.id_count += 1
                track['tracking_id'] = self.id_count
                if TRK_REF[track['detection_name']]['ref']:
                    beta = TRK_REF[track['detection_name']]['beta']
                    track['ref_detection_score'] = beta * track['detection_score']
                else:
                    track['ref_detection_score'] = track['detection_score']
                track['age'] = 1
                track['active'] = 1
                ret.append(track)
            for i in unmatched_tracks:
                track = curr_tracks[i]
                if 'dead' in track.keys() and track_near[i]:
                    continue
                if track['age'] < self.max_age:
                    track['age'] += 1
                    track['active'] = 0
                    ct = track['ct']
                    if TRK_REF[track['detection_name']]['ref']:
                        beta = TRK_REF[track['detection_name']]['beta']
                        track['ref_detection_score'] = (1 - beta) * track['ref_detection_score']
                    if 'tracking' in track:
                        offset = track['tracking'] * -1  # move forward
                        track['ct'] = ct + offset
                    ret.append(track)
        self.tracks = ret
        return ret
