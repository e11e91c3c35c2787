"""TEST INFRASTRUCTURE — CPU restatement of the greedy ID tracker (tools/nusc_shasta/pub_tracker_merged.py:72-225,
track_utils.py:3-14) in plain numpy with the distance matrix materialised exactly like the reference. Only tests may
import it. Pinned against the UNMODIFIED reference tracker run in the authoring container: oracle/make_tracker_golden.py
wrote tests/golden/tracker_*.json, tests/test_tracker.py replays them."""
import copy

import numpy as np

NAMES = ['bicycle', 'bus', 'car', 'motorcycle', 'pedestrian', 'trailer', 'truck']
VEL_ERR = {'car': 2, 'truck': 2, 'bus': 4, 'trailer': 2, 'pedestrian': 0.75, 'motorcycle': 2, 'bicycle': 1.5}
BETA = {'bicycle': 0.4, 'bus': 0.7, 'car': 0.5, 'motorcycle': 0.5, 'pedestrian': 0.5, 'trailer': 0.4, 'truck': 0.5}
ALPHA = 0.5


def greedy_assignment(dist):
    """track_utils.py:3-14."""
    matched = []
    if dist.shape[1] == 0 or dist.shape[0] == 0:
        return np.array(matched, np.int32).reshape(-1, 2)
    for i in range(dist.shape[0]):
        j = dist[i].argmin()
        if dist[i][j] < 1e16:
            dist[:, j] = 1e18
            matched.append([i, j])
    return np.array(matched, np.int32).reshape(-1, 2)


def masked_distance(dets, tracks, max_diff, item_cat, track_cat):
    """pub_tracker_merged.py:123-131."""
    N, M = dets.shape[0], tracks.shape[0]
    dist = (((tracks.reshape(1, -1, 2) - dets.reshape(-1, 1, 2)) ** 2).sum(axis=2))
    dist = np.sqrt(dist)
    invalid = ((dist > max_diff.reshape(N, 1)) + (item_cat.reshape(N, 1) != track_cat.reshape(1, M))) > 0
    return dist + invalid * 1e18


class Tracker:
    def __init__(self, max_age=0):
        self.max_age = max_age
        self.reset()

    def reset(self):
        self.id_count = 0
        self.tracks = []

    def step(self, results, time_lag):
        if len(results) == 0:
            self.tracks = []
            return []
        ret = []
        for name in NAMES:
            cur = []
            for det in results:
                if det['detection_name'] != name:
                    continue
                det['ct'] = np.array(det['translation'][:2])
                det['tracking'] = np.array(det['velocity'][:2]) * -1 * time_lag
                det['label_preds'] = NAMES.index(name)
                cur.append(det)
            if not cur:
                continue
            trk = [t for t in self.tracks if t['detection_name'] == name]
            dets = np.array([d['ct'] + d['tracking'].astype(np.float32) for d in cur], np.float32)
            item_cat = np.array([d['label_preds'] for d in cur], np.int32)
            track_cat = np.array([t['label_preds'] for t in trk], np.int32)
            max_diff = np.array([VEL_ERR[name] for _ in cur], np.float32)
            tracks = np.array([t['ct'] for t in trk], np.float32)
            if len(tracks) > 0:
                dist = masked_distance(dets, tracks, max_diff, item_cat, track_cat)
                matches = greedy_assignment(copy.deepcopy(dist))
            else:
                dist = None
                matches = np.array([], np.int32).reshape(-1, 2)
            un_d = [d for d in range(dets.shape[0]) if d not in matches[:, 0]]
            un_t = [d for d in range(tracks.shape[0]) if d not in matches[:, 1]]
            beta = BETA[name]
            for m in matches:
                t = cur[m[0]]
                t['tracking_id'] = trk[m[1]]['tracking_id']
                t['ref_detection_score'] = ((t['ref_detection_score'] > ALPHA) * beta * t['detection_score']
                                            + (1 - beta) * trk[m[1]]['ref_detection_score'])
                t['age'] = 1
                t['active'] = trk[m[1]]['active'] + 1
                ret.append(t)
            for i in un_d:
                t = cur[i]
                if len(tracks) > 0 and 'newborn' not in t and (dist[i, :] <= VEL_ERR[name]).sum():
                    continue
                self.id_count += 1
                t['tracking_id'] = self.id_count
                t['ref_detection_score'] = beta * t['detection_score']
                t['age'] = 1
                t['active'] = 1
                ret.append(t)
            for i in un_t:
                t = trk[i]
                if 'dead' in t and (dist[:, i] <= VEL_ERR[name]).sum():
                    continue
                if t['age'] < self.max_age:
                    t['age'] += 1
                    t['active'] = 0
                    t['ref_detection_score'] = (1 - beta) * t['ref_detection_score']
                    if 'tracking' in t:
                        t['ct'] = t['ct'] + t['tracking'] * -1
                    ret.append(t)
        self.tracks = ret
        return ret


def synthetic_sequence(seed, frames=12, per_class=(3, 14), names=('car', 'pedestrian', 'bus', 'truck'), flags=True):
    """A scene's worth of per-frame detection lists in the layout eval.py writes (translation, velocity,
    detection_name, detection_score, ref_detection_score, optional newborn / dead flags): objects move with their
    velocity plus noise, some disappear, some appear, some detections are duplicated nearby."""
    rng = np.random.default_rng(seed)
    objs = []
    for name in names:
        for _ in range(int(rng.integers(*per_class))):
            objs.append({'name': name, 'p': rng.uniform(-40, 40, 2), 'v': rng.normal(0, 3, 2)})
    out = []
    for f in range(frames):
        dets = []
        for o in objs:
            if rng.random() < 0.1:
                continue
            p = o['p'] + rng.normal(0, 0.25, 2)
            d = {'translation': [float(p[0]), float(p[1]), float(rng.normal(0, 1))],
                 'velocity': [float(o['v'][0]), float(o['v'][1])], 'detection_name': o['name'],
                 'detection_score': float(rng.uniform(0.1, 1)), 'ref_detection_score': float(rng.uniform(0, 1)),
                 'sample_token': 'f%d' % f}
            if flags and rng.random() < 0.15:
                d['newborn'] = True
            if flags and rng.random() < 0.1:
                d['dead'] = True
            dets.append(d)
            if rng.random() < 0.08:  # a near-duplicate detection
                q = copy.deepcopy(d)
                q['translation'][0] += float(rng.normal(0, 0.3))
                q.pop('newborn', None)
                dets.append(q)
        for o in objs:
            o['p'] = o['p'] + o['v'] * 0.5
        if rng.random() < 0.3:
            objs.append({'name': names[int(rng.integers(len(names)))], 'p': rng.uniform(-40, 40, 2), 'v': rng.normal(0, 3, 2)})
        if f == 7:
            dets = [] if seed % 2 else dets   # an empty frame resets the track list
        order = rng.permutation(len(dets))
        out.append([dets[i] for i in order])
    return out


def summarize(ret):
    """Order-preserving, JSON-able view of a step's output."""
    return [{'name': t['detection_name'], 'id': int(t['tracking_id']), 'age': int(t['age']), 'active': int(t['active']),
             'score': float(t['ref_detection_score']), 'x': float(t['translation'][0]), 'ctx': float(t['ct'][0]),
             'cty': float(t['ct'][1])} for t in ret]
