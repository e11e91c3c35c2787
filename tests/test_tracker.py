"""Greedy ID tracker: the numpy oracle against the reference's own outputs (committed fixtures written
by oracle/make_tracker_golden.py from the unmodified pub_tracker_merged.py), and the product tracker - host
bookkeeping + shasta_greedy_assign_f32 on the GPU - against both."""
import copy
import glob
import json
import os

import numpy as np
import pytest

from oracle import tracker_oracle as TO

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tracker_*.json")))


def _load(path):
    with open(path) as f:
        return json.load(f)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_fixture(path):
    g = _load(path)
    trk = TO.Tracker(max_age=g["max_age"])
    for frame, want in zip(copy.deepcopy(TO.synthetic_sequence(g["seed"])), g["outputs"]):
        got = TO.summarize(trk.step(frame, g["time_lag"]))
        assert got == want


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


def test_oracle_greedy_assignment_edge_cases():
    assert TO.greedy_assignment(np.zeros((0, 3))).shape == (0, 2)
    assert TO.greedy_assignment(np.zeros((3, 0))).shape == (0, 2)
    d = np.array([[1.0, 1.0, 5.0], [1.0, 1e18, 1e18], [1e18, 1e18, 1e18]])
    assert TO.greedy_assignment(d.copy()).tolist() == [[0, 0]]   # tie -> first column; row 1 loses its only option


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_device_tracker_matches_reference_fixture(path):
    from shasta_b200.tracker import PubTrackerMerged
    g = _load(path)
    trk = PubTrackerMerged(hungarian=False, max_age=g["max_age"])
    for frame, want in zip(copy.deepcopy(TO.synthetic_sequence(g["seed"])), g["outputs"]):
        got = TO.summarize(trk.step_centertrack(frame, g["time_lag"]))
        assert got == want


@pytest.mark.gpu
def test_device_tracker_matches_oracle_on_long_crowded_sequences():
    from shasta_b200.tracker import PubTrackerMerged
    for seed in (11, 12):
        seq = TO.synthetic_sequence(seed, frames=20, per_class=(40, 90), names=TO.NAMES)
        a, b = TO.Tracker(max_age=3), PubTrackerMerged(max_age=3)
        for f1, f2 in zip(copy.deepcopy(seq), copy.deepcopy(seq)):
            assert TO.summarize(b.step_centertrack(f2, 0.5)) == TO.summarize(a.step(f1, 0.5))


@pytest.mark.gpu
def test_greedy_assign_kernel_against_numpy():
    """Random padded batches incl. empty problems, category mismatches, exact ties and far-away points."""
    from shasta_b200.tracker import greedy_assign_batch
    rng = np.random.default_rng(0)
    dets, tracks, md, dc, tc = [], [], [], [], []
    for p in range(40):
        n, m = int(rng.integers(0, 300)), int(rng.integers(1, 300))
        if p == 3:
            n = 0
        d = rng.uniform(-20, 20, (n, 2)).astype(np.float32)
        t = rng.uniform(-20, 20, (m, 2)).astype(np.float32)
        if n > 4 and m > 4:
            t[:4] = d[:4]            # zero distances
            t[5 % m] = t[4]          # duplicated track -> tie
        dets.append(d), tracks.append(t)
        md.append(rng.choice([0.75, 1.5, 2.0, 4.0], n).astype(np.float32))
        dc.append(rng.integers(0, 2, n).astype(np.int32)), tc.append(rng.integers(0, 2, m).astype(np.int32))
    match, dnear, tnear = greedy_assign_batch(dets, tracks, md, dc, tc)
    for p in range(40):
        n, m = len(dets[p]), len(tracks[p])
        want = -np.ones(n, np.int32)
        if n:
            dist = TO.masked_distance(dets[p], tracks[p], md[p], dc[p], tc[p])
            for i, j in TO.greedy_assignment(dist.copy()):
                want[i] = j
            assert np.array_equal(dnear[p], (dist < 1e16).any(axis=1))
            assert np.array_equal(tnear[p], (dist < 1e16).any(axis=0))
        else:
            assert not tnear[p].any()
        assert np.array_equal(match[p], want), p


@pytest.mark.gpu
def test_hungarian_is_refused():
    from shasta_b200.tracker import PubTrackerMerged
    with pytest.raises(NotImplementedError):
        PubTrackerMerged(hungarian=True)
